"""Repeats the same cold translation solve and checks that iteration counts are identical
(development tool: the solve must be deterministic)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dpgo_b200 as D
nx, ny, nz = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "100,100,100").split(","))
nodes = int(sys.argv[2]) if len(sys.argv) > 2 else 64
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
g, _, X0 = D.grid3d(nx, ny, nz)
drv = D.DPGOStar(g, nodes, D.Options(loss="trivial"))
assert drv.initialize(X0) == 0 and drv.update() == 0 and drv.iterate() == 0
seen = collections.Counter()
for i in range(n):
    drv.reset_counters()
    ms = drv.profile_pass("g00_solve", 1)      # 3 warm-up + 1 timed identical solves
    c = drv.counters()
    seen[(c.solve_iters, c.reserved[0])] += 1
    print(i, c.solve_iters, c.reserved[0], "%.3f ms" % ms, flush=True)
print("distinct (node-iterations, pose-iterations) over %d repeats of 4 solves: %s" % (n, dict(seen)))
print("last: %.3f ms" % ms)

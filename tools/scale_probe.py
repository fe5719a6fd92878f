"""Scale probe: per-iteration time, counters and per-kernel device times at a given grid size."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import dpgo_b200 as D

nx, ny, nz, nodes = (int(v) for v in sys.argv[1:5])
alg = sys.argv[5] if len(sys.argv) > 5 else "star"
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 10
loss = sys.argv[7] if len(sys.argv) > 7 else "trivial"
t0 = time.time()
g, Xgt, X0 = D.grid3d(nx, ny, nz)
print("graph: N=%d E=%d gen %.1fs" % (g.num_poses, g.num_edges, time.time() - t0)); sys.stdout.flush()
t0 = time.time()
cls = D.DPGOStar if alg == "star" else D.DPGOHash
drv = cls(g, nodes, D.Options(loss=loss))
print("setup %.1fs sizes=%s" % (time.time() - t0, drv.sizes())); sys.stdout.flush()
t0 = time.time()
assert drv.initialize(X0) == 0
print("initialize %.3fs" % (time.time() - t0))
assert drv.update() == 0
drv.synchronize()
for it in range(iters):
    drv.reset_counters()
    f, gn = drv.objective()
    t0 = time.time()
    rc = drv.iterate(); assert rc == 0, D.load().mmpgo_last_error()
    drv.synchronize(); t1 = time.time()
    drv.communicate(); drv.update(); drv.synchronize(); t2 = time.time()
    c = drv.counters()
    ref = sum(drv.node_scalars(a).refined for a in range(nodes))
    print("it %2d 2F=%.10g |g|=%.4g iterate %.2fms update %.2fms launches=%d intra=%d inter=%d prox=%d solves=%d solve_it=%d tcg=%d tnt=%d vec=%d refined=%d" % (
        it, 2 * f, 2 * gn, 1e3 * (t1 - t0), 1e3 * (t2 - t1), c.launches, c.intra_passes, c.inter_passes, c.prox_passes,
        c.solve_calls, c.solve_iters, c.tcg_iterations, c.tnt_iterations, c.vector_passes, ref))
    sys.stdout.flush()
s = drv.sizes()
E_intra = s["bsr_entries"] // 2
N = s["own_poses"]
for k in drv.KERNEL_KINDS:
    ms = drv.profile_pass(k, 20)
    print("kernel %-15s %.4f ms" % (k, ms))
ms = drv.profile_pass("k2_eval", 20)
alg_bytes = 120 * E_intra + 192 * N
print("k2_eval: algorithmic %.1f MB -> %.1f GB/s" % (alg_bytes / 1e6, alg_bytes / ms / 1e6))

"""Times the persistent translation solve (profile kind 8) on the C4-shaped graph under the
MMPGO_TS_* experiment knobs (development tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dpgo_b200 as D
nx, ny, nz = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "100,100,100").split(","))
nodes = int(sys.argv[2]) if len(sys.argv) > 2 else 64
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
grids = [int(v) for v in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0, 148]
modes = [int(v) for v in sys.argv[5].split(",")] if len(sys.argv) > 5 else [0]
g, _, X0 = D.grid3d(nx, ny, nz)
drv = D.DPGOStar(g, nodes, D.Options(loss="trivial"))
assert drv.initialize(X0) == 0 and drv.update() == 0 and drv.iterate() == 0
for iters in (20,):
    for grid in grids:
      for mode in modes:
        os.environ["MMPGO_TS_ITERS"] = str(iters); os.environ["MMPGO_TS_MODE"] = str(mode); os.environ["MMPGO_TS_GRID"] = str(grid)
        ms = drv.profile_pass("g00_solve", reps)
        print("iters %d grid %d mode %d: %.3f ms/solve  %.1f us/iter" % (iters, grid, mode, ms, 1e3 * ms / iters), flush=True)

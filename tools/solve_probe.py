"""Times / profiles the translation solve alone on the benchmark graph:
    python tools/solve_probe.py [reps] [solver] [nx,ny,nz] [nodes]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import os
import dpgo_b200 as D
if os.environ.get("MMPGO_LIB"):          # A/B builds of the library (development)
    import dpgo_b200.lib as _L
    _L.LIB_PATH = os.environ["MMPGO_LIB"]

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
solver = sys.argv[2] if len(sys.argv) > 2 else "direct"
dims = tuple(int(v) for v in sys.argv[3].split(",")) if len(sys.argv) > 3 else (100, 100, 100)
nodes = int(sys.argv[4]) if len(sys.argv) > 4 else 64
g, _, X0 = D.grid3d(*dims)
t0 = time.time()
drv = D.DPGOStar(g, nodes, D.Options(dense_solve_max_n=0, translation_solver=solver))
print("set_graph %.1f s" % (time.time() - t0), drv.solver_info())
assert drv.initialize(X0) == 0 and drv.update() == 0 and drv.iterate() == 0
if solver == "direct" and "nodry" not in sys.argv:
    print("g00_solve dry (stages + barriers only) ms:", drv.profile_pass("g00_solve_dry", reps))
    for us, wj, cj in drv.solver_stage_times()[:3]:
        print("dry stage %8.1f us  warp jobs %7d" % (us, wj))
print("g00_solve ms:", drv.profile_pass("g00_solve", reps))
if solver == "direct":
    print("g00_solve with per-supernode dependencies ms:", drv.profile_pass("g00_solve_deps", reps))
    print("g00_solve with a grid barrier per level ms:", drv.profile_pass("g00_solve_levels", reps))
for us, wj, cj in drv.solver_stage_times():
    print("stage %8.1f us  warp jobs %7d  cta jobs %5d" % (us, wj, cj))

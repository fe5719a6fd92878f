"""Times / profiles the K2 passes (and K1, K3) alone on the benchmark graph:
    python tools/k2_probe.py [reps] [kinds,comma,separated] [nx,ny,nz] [nodes]"""
import sys, time
sys.path.insert(0, ".")
import os
import dpgo_b200 as D
if os.environ.get("MMPGO_LIB"):          # A/B builds of the library (development)
    import dpgo_b200.lib as _L
    _L.LIB_PATH = os.environ["MMPGO_LIB"]

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
kinds = sys.argv[2].split(",") if len(sys.argv) > 2 else ["k2_eval", "k2_grad", "k2_hv", "k2_g01", "k1_inter", "k3_prox", "edge_objective"]
dims = tuple(int(v) for v in sys.argv[3].split(",")) if len(sys.argv) > 3 else (100, 100, 100)
nodes = int(sys.argv[4]) if len(sys.argv) > 4 else 64
if os.environ.get("PROBE_SPHERE"):       # PROBE_SPHERE=robots,poses_per_robot: the configs[4] shape, Welsch, AMM-PGO#
    r, ppr = (int(v) for v in os.environ["PROBE_SPHERE"].split(","))
    g, _, X0 = D.sphere_rings(r, ppr)
    nodes = r
    t0 = time.time()
    drv = D.DPGOHash(g, nodes, D.Options(loss="welsch"))
else:
    g, _, X0 = D.grid3d(*dims)
    t0 = time.time()
    drv = D.DPGOStar(g, nodes, D.Options())
print("set_graph %.1f s" % (time.time() - t0))
assert drv.initialize(X0) == 0 and drv.update() == 0 and drv.iterate() == 0
for k in kinds:
    print("%-16s %.4f ms" % (k, drv.profile_pass(k, reps)))

"""Copy-ring translation solve on the 1M-pose graph: tile-dealing chunk size vs cold-solve and step time (development tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dpgo_b200 as D
g, _, X0 = D.grid3d(100, 100, 100)
drv = D.DPGOStar(g, 64, D.Options(loss="trivial"))
assert drv.initialize(X0) == 0 and drv.update() == 0
for _ in range(3):
    assert drv.iterate() == 0; drv.communicate(); drv.update()
for k in ("k2_eval", "k2_grad", "k2_hv", "k2_g01", "k1_inter", "k3_prox"):
    print("kernel %s %.4f ms" % (k, drv.profile_pass(k, 20)), flush=True)
for chunk in (sys.argv[1:] or ["8", "4", "2", "1", "16"]):
    os.environ["MMPGO_TS_CHUNK"] = chunk
    cold = drv.profile_pass("g00_solve", 5)
    drv.synchronize(); t0 = time.time()
    for _ in range(6):
        assert drv.iterate() == 0; drv.communicate(); drv.update()
    drv.synchronize()
    print("chunk %s: cold solve %.3f ms, step %.3f ms" % (chunk, cold, (time.time() - t0) / 6 * 1e3), flush=True)

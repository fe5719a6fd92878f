"""Copy-ring translation solve with and without the hand-off of its tail to k_tsolve_lite (development tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dpgo_b200 as D
dims = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "100,100,100").split(","))
nodes = int(sys.argv[2]) if len(sys.argv) > 2 else 64
g, _, X0 = D.grid3d(*dims)
os.environ["MMPGO_TS_KERNEL"] = "ring"
drv = D.DPGOStar(g, nodes, D.Options(loss="trivial"))
assert drv.initialize(X0) == 0 and drv.update() == 0
for _ in range(3):
    assert drv.iterate() == 0; drv.communicate(); drv.update()
for ho in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["0", "2", "4", "8", "16"]):
    os.environ["MMPGO_TS_HANDOFF"] = ho
    cold = drv.profile_pass("g00_solve", 5)
    drv.synchronize(); t0 = time.time()
    for _ in range(6):
        assert drv.iterate() == 0, D.load().mmpgo_last_error(); drv.communicate(); drv.update()
    drv.synchronize()
    print("handoff at %s live nodes: cold solve %.3f ms, step %.3f ms, 2F=%.12g" % (ho, cold, (time.time() - t0) / 6 * 1e3, 2 * drv.objective()[0]), flush=True)

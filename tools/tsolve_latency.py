import os, sys
sys.path.insert(0, "/root/repo")
import dpgo_b200 as D
for dims, nodes in (((16,16,8),1), ((16,16,8),2), ((32,32,16),4), ((50,50,50),8)):
    g, _, X0 = D.grid3d(*dims)
    drv = D.DPGOStar(g, nodes, D.Options(loss="trivial", dense_solve_max_n=0))
    assert drv.initialize(X0) == 0 and drv.update() == 0 and drv.iterate() == 0
    os.environ["MMPGO_TS_ITERS"] = "100"
    ms = drv.profile_pass("g00_solve", 5)
    print(dims, nodes, "nodes: %.3f ms per 100-iteration solve -> %.1f us/iteration" % (ms, 10 * ms), flush=True)

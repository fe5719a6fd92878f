"""Ring vs lite translation-solve kernels on 1 GPU at the per-GPU shard sizes of the 8/4/2-GPU runs
of the 1M-pose workload (development tool): fixed-iteration solve latency, then whole AMM-PGO* steps."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dpgo_b200 as D

cases = [((100, 100, 13), 8), ((100, 100, 25), 16), ((100, 100, 50), 32)]
if len(sys.argv) > 1:
    cases = cases[: int(sys.argv[1])]
for dims, nodes in cases:
    g, _, X0 = D.grid3d(*dims)
    for kernel in ("ring", "lite"):
        os.environ["MMPGO_TS_KERNEL"] = kernel
        os.environ.pop("MMPGO_TS_ITERS", None)
        drv = D.DPGOStar(g, nodes, D.Options(loss="trivial"))
        assert drv.initialize(X0) == 0 and drv.update() == 0
        for _ in range(3):
            assert drv.iterate() == 0, D.load().mmpgo_last_error()
            drv.communicate(); drv.update()
        drv.synchronize()
        t0 = time.time()
        for _ in range(10):
            assert drv.iterate() == 0
            drv.communicate(); drv.update()
        drv.synchronize()
        step_ms = 100 * (time.time() - t0)
        cold = drv.profile_pass("g00_solve", 5)
        os.environ["MMPGO_TS_ITERS"] = "100"
        fixed = drv.profile_pass("g00_solve", 5)
        print("%s nodes=%d poses=%d kernel=%s: step %.3f ms, cold solve %.3f ms, 100-iteration solve %.3f ms = %.1f us/iteration, 2F=%.12g"
              % (dims, nodes, g.num_poses, kernel, step_ms, cold, fixed, 10 * fixed, 2 * drv.objective()[0]), flush=True)
        del drv

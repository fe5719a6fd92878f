"""First-contact GPU probe: prints oracle-vs-CUDA diagnostics for a ladder of small cases."""
import sys, time, traceback
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import dpgo_b200 as D
import parity

def case(name, g, nn, X0, iters, **kw):
    t0 = time.time()
    try:
        out = parity.run_both(g, nn, X0, iters, **kw)
    except Exception:
        print("CASE", name, "FAILED"); traceback.print_exc(); return
    err = parity.rel_trace_error(out)
    et, eR = parity.pose_error(out, g.d)
    nodeerr = np.abs(out["fobj_ref"] - out["fobj_gpu"]) / np.abs(out["fobj_ref"])
    print("CASE %-28s iters=%d max_rel_F=%.3e first5=%s node_max=%.3e pose_t=%.2e pose_R=%.2e refined_eq=%s F0=%.8g Fend=%.8g  (%.1fs)" % (
        name, iters, err.max(), np.array2string(err[:5], precision=2), nodeerr.max(), et, eR,
        bool((out["refined_ref"] == out["refined_gpu"]).all()), out["fobj_ref"][0].sum(), out["fobj_ref"][-1].sum(), time.time() - t0))
    c = out["drv"].counters()
    print("     counters: launches=%d intra=%d inter=%d prox=%d solves=%d solve_iters=%d tcg=%d tnt=%d" % (
        c.launches, c.intra_passes, c.inter_passes, c.prox_passes, c.solve_calls, c.solve_iters, c.tcg_iterations, c.tnt_iterations))
    sys.stdout.flush()

g, Xgt, X0 = D.grid3d(6, 6, 6, seed=1)
case("grid6 trivial it0", g, 4, X0, 0)
case("grid6 trivial mm noTNT", g, 4, X0, 3, scheme="MM", max_iterations=0)
case("grid6 trivial amm noTNT", g, 4, X0, 5, max_iterations=0)
case("grid6 trivial mm", g, 4, X0, 3, scheme="MM")
case("grid6 trivial amm", g, 4, X0, 10)
case("grid6 trivial amm jacobi", g, 4, X0, 10, preconditioner="Jacobi")
case("grid6 trivial amm noprec", g, 4, X0, 10, preconditioner="None")
case("grid6 huber amm noTNT", g, 4, X0, 5, loss="huber", max_iterations=0)
case("grid6 huber amm", g, 4, X0, 10, loss="huber")
case("grid6 gm amm", g, 4, X0, 10, loss="gm")
case("grid6 welsch amm", g, 4, X0, 10, loss="welsch")
case("grid6 trivial star", g, 4, X0, 10, algorithm="star")
case("grid6 huber star", g, 4, X0, 10, algorithm="star", loss="huber")
case("grid6 trivial amm pcg", g, 4, X0, 10, dense_solve_max_n=0)
g2, _, X2 = D.city2d(12, 12, seed=2)
case("city12 trivial amm", g2, 4, X2, 10)
case("city12 gm amm", g2, 4, X2, 10, loss="gm")
g3, _, X3 = D.sphere_rings(6, 40, seed=3)
case("sphere6x40 welsch amm", g3, 6, X3, 10, loss="welsch")
case("grid6 trivial amm 50", g, 4, X0, 50)

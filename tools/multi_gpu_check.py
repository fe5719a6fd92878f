"""torchrun script: the sharded run over W ranks (one process per GPU, NCCL) reproduces the single-handle run,
with the NCCL transport driven from C++ (mmpgo_nccl_init) and with the torch.distributed callbacks.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py"""
import faulthandler, os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, ".")
import dpgo_b200 as D
from dpgo_b200 import multi

faulthandler.dump_traceback_later(int(os.environ.get("MMPGO_CHECK_WATCHDOG", "240")), exit=True)   # a hang becomes a stack trace, not a lost GPU hour
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ok_all = True
CASES = (("hash", "trivial", "grid", (8, 8, 8), 8, 12, 2048, True), ("hash", "huber", "grid", (8, 8, 8), 8, 12, 2048, False),
         ("star", "trivial", "grid", (8, 8, 8), 8, 12, 2048, True), ("star", "welsch", "grid", (8, 8, 8), 8, 12, 0, True),
         ("star", "trivial", "grid", (30, 30, 16), 16, 6, 0, True), ("star", "trivial", "grid", (30, 30, 16), 16, 6, 0, False),
         # BASELINE.json configs[4] in small: multi-robot sphere, Welsch, decentralised AMM-PGO#
         ("hash", "welsch", "sphere", (16, 600), 16, 10, 0, True))
for (alg, loss, kind, dims, nodes, iters, dense, native) in CASES:
    faulthandler.dump_traceback_later(int(os.environ.get("MMPGO_CHECK_WATCHDOG", "240")), exit=True)
    if rank == 0:
        print("case", alg, loss, kind, dims, "native" if native else "callbacks", flush=True)
    g, _, X0 = D.grid3d(*dims, seed=4) if kind == "grid" else D.sphere_rings(dims[0], dims[1], seed=4)
    drv = multi.make_driver(g, nodes, D.Options(loss=loss, device=lr, dense_solve_max_n=dense), alg, rank, world, native_nccl=native)
    assert drv.initialize(X0) == 0 and drv.update() == 0
    tr = [drv.global_objective()[0]]
    for _ in range(iters):
        D.lib.check(drv.iterate()); D.lib.check(drv.communicate()); D.lib.check(drv.update())
        tr.append(drv.global_objective()[0])
    Xl = drv.X()
    Xs = torch.from_numpy(np.ascontiguousarray(Xl)).cuda()
    dist.all_reduce(Xs)          # ranks own disjoint rows; the rest are zero
    if rank == 0:
        ref, tr_ref = D.run_dist_pgo(g, nodes, X0, iters, D.Options(loss=loss, device=lr, dense_solve_max_n=dense), alg)
        tr_ref = np.array([t[0] / 2 for t in tr_ref])
        err = np.abs(np.array(tr) - tr_ref) / np.abs(tr_ref)
        perr = np.abs(Xs.cpu().numpy() - ref.X()).max()
        sc, rc = drv.halo_counts()
        # AMM-PGO# has no scalar collective: the same bits; AMM-PGO* sums its global scalars per rank first
        ok = (err.max() < 1e-13 and perr == 0.0) if alg == "hash" else (err.max() < 1e-9 and perr < 1e-7)
        ok_all &= ok
        print("%s %-8s %s %s nodes=%d world=%d transport=%s: max rel F err %.2e pose err %.2e send=%s %s" % (
            alg, loss, kind, dims, nodes, world, drv.transport, err.max(), perr, sc.tolist(), "OK" if ok else "FAIL"))
    dist.barrier()
    drv.close()
if rank == 0:
    print("MULTI_GPU_CHECK", "PASS" if ok_all else "FAIL")
dist.destroy_process_group()

"""torchrun script: the sharded run over W ranks reproduces the single-handle run.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py"""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, ".")
import dpgo_b200 as D
from dpgo_b200 import multi

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ok_all = True
for (alg, loss, dims, nodes, iters, dense) in (("hash", "trivial", (8, 8, 8), 8, 12, 2048), ("hash", "huber", (8, 8, 8), 8, 12, 2048),
                                               ("star", "trivial", (8, 8, 8), 8, 12, 2048), ("star", "welsch", (8, 8, 8), 8, 12, 0),
                                               ("star", "trivial", (30, 30, 16), 16, 6, 0)):
    g, _, X0 = D.grid3d(*dims, seed=4)
    drv = multi.make_driver(g, nodes, D.Options(loss=loss, device=lr, dense_solve_max_n=dense), alg, rank, world)
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
        ok = err.max() < 1e-9 and perr < 1e-7
        ok_all &= ok
        print("%s %-8s %s nodes=%d world=%d: max rel F err %.2e pose err %.2e exchanges=%d allreduces=%d send=%s %s" % (
            alg, loss, dims, nodes, world, err.max(), perr, drv.exchanges, drv.allreduces, sc.tolist(), "OK" if ok else "FAIL"))
    dist.barrier()
if rank == 0:
    print("MULTI_GPU_CHECK", "PASS" if ok_all else "FAIL")
dist.destroy_process_group()

"""N > 1 host logic on CPU: world_size-2 (and 3) gloo processes run the exchange plan of the
sharded driver (mmpgo_plan_halo) and push pose ids through the same all_to_all_single /
all_reduce calls the GPU path uses."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import dpgo_b200 as D
from dpgo_b200 import multi


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nodes, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g, _, _ = D.grid3d(6, 6, 6, seed=1)
        sc, rc, sg, rg = multi.plan_halo(g, nodes, world, rank)
        # the wire carries one payload per boundary pose: use the pose id itself
        send = torch.from_numpy(sg.astype(np.float64))
        recv = torch.empty(int(rc.sum()), dtype=torch.float64)
        dist.all_to_all_single(recv, send, [int(x) for x in rc], [int(x) for x in sc])
        ok = np.array_equal(recv.numpy().astype(np.int64), rg)
        # a pose is received by exactly the ranks it is sent to
        tot = torch.tensor([float(sc.sum()), float(rc.sum())])
        dist.all_reduce(tot)
        ok = ok and tot[0].item() == tot[1].item() and sc[rank] == 0 and rc[rank] == 0
        # every received pose belongs to the peer it came from
        rnb = multi.rank_node_begin(nodes, world)
        off = 0
        for peer in range(world):
            ids = rg[off:off + rc[peer]]
            off += rc[peer]
            from oracle import g2o as og2o
            node, _ = og2o.partition_index(g.num_poses, nodes, ids)
            ok = ok and bool(np.all((node >= rnb[peer]) & (node < rnb[peer + 1])))
        # two-array exchange of AMM-PGO* (per-peer chunks [a | b]): pack with the library's index
        # maps, ONE all_to_all with doubled counts, unpack into the halo rows of both arrays
        n_own = 1000 + rank                        # any row offset of the halo
        sa, sb, ra, rb, hrow = multi.plan_halo_pair(sc, rc, n_own)
        ns, nr = int(sc.sum()), int(rc.sum())
        buf = np.zeros(2 * ns)
        buf[sa] = sg                               # array a carries the pose id, array b its negative
        buf[sb] = -sg.astype(np.float64) - 0.5
        recv2 = torch.empty(2 * nr, dtype=torch.float64)
        dist.all_to_all_single(recv2, torch.from_numpy(buf), [2 * int(x) for x in rc], [2 * int(x) for x in sc])
        xa, xb = np.zeros(n_own + nr), np.zeros(n_own + nr)
        xa[hrow] = recv2.numpy()[ra]
        xb[hrow] = recv2.numpy()[rb]
        ok = ok and np.array_equal(xa[n_own:], rg.astype(np.float64))
        ok = ok and np.array_equal(xb[n_own:], -rg.astype(np.float64) - 0.5)
        ok = ok and np.array_equal(hrow, n_own + np.arange(nr))
        q.put((rank, bool(ok), int(sc.sum())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nodes", [(2, 4), (3, 6)])
def test_halo_plan_over_gloo(world, nodes):
    ctx = mp.get_context("spawn")
    res = None
    for attempt in range(2):          # a second rendezvous on another port if the first one was lost (port taken
        q = ctx.Queue()               # between _free_port() and the bind, or a loaded machine)
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, nodes, q)) for r in range(world)]
        for p in procs:
            p.start()
        try:
            res = [q.get(timeout=300) for _ in procs]
        except Exception:
            res = None
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.terminate()
        if res is not None:
            break
    assert res is not None, "no result from the gloo ranks"
    assert all(ok for _, ok, _ in res), res
    assert sum(n for _, _, n in res) > 0


def test_plan_matches_reference_sent_recv_semantics():
    # sent_/recv_ of generate_data_info (DPGO_utils.cpp:426-435) restricted to rank boundaries
    from oracle import data_matrix as dm, g2o as og2o
    from parity import to_measurements
    g, _, _ = D.grid3d(5, 5, 4, seed=3)
    nodes, world = 4, 2
    per_node, g_index, part = og2o.partition(g.num_poses, nodes, to_measurements(g))
    rnb = multi.rank_node_begin(nodes, world)
    for rank in range(world):
        sc, rc, sg, rg = multi.plan_halo(g, nodes, world, rank)
        want_send, want_recv = set(), set()
        for a in range(rnb[rank], rnb[rank + 1]):
            info = dm.generate_data_info(a, per_node[a])
            for b, poses in info.sent.items():
                if not (rnb[rank] <= b < rnb[rank + 1]):
                    want_send |= {g_index[a][p] for p in poses}
            for b, poses in info.recv.items():
                if not (rnb[rank] <= b < rnb[rank + 1]):
                    want_recv |= {g_index[b][p] for p in poses}
        assert set(sg.tolist()) == want_send
        assert set(rg.tolist()) == want_recv

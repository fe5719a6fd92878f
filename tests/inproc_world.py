"""Several sharded handles in ONE process on ONE GPU (test infrastructure).

Each rank of a world of W handles runs in its own thread; the transport callbacks of
mmpgo_set_sharding (include/mmpgo.h) are served by device-to-device copies between the handles'
send and receive buffers, with thread barriers where NCCL would rendezvous.  The exchange plan,
the packing kernels, the halo rows and the all-reduce call sites are exactly those of the
multi-GPU run, so a single-GPU box can check that a sharded run reproduces the single-handle run.
"""
from __future__ import annotations

import threading

import numpy as np

import dpgo_b200 as D
from dpgo_b200 import lib as L
from dpgo_b200 import multi


class _Rank:
    def __init__(self, world, rank, graph, num_nodes, options, algorithm):
        import torch
        self.torch = torch
        self.world, self.rank = world, rank
        rnb = multi.rank_node_begin(num_nodes, world.W)
        cls = D.DPGOStar if algorithm == "star" else D.DPGOHash
        self.drv = cls(graph, num_nodes, options, int(rnb[rank]), int(rnb[rank + 1]))
        self._ex = L.EXCHANGE_FN(self._exchange)
        self._ar = L.ALLREDUCE_FN(self._allreduce)
        L.check(self.drv.lib.mmpgo_set_sharding(self.drv._h, rank, world.W, L.iptr(rnb), self._ex, self._ar, None))
        self.exchanges = 0
        self.allreduces = 0

    def _view(self, ptr, n):
        dev = self.torch.device("cuda", self.torch.cuda.current_device())
        if n == 0:
            return self.torch.empty(0, dtype=self.torch.float64, device=dev)
        return self.torch.as_tensor(multi._DevArray(ptr, n), device=dev)

    def _exchange(self, user, send_ptr, send_counts, recv_ptr, recv_counts):
        try:
            w = self.world
            W = w.W
            sc = [int(send_counts[q]) for q in range(W)]
            rc = [int(recv_counts[q]) for q in range(W)]
            self.drv.synchronize()                      # the send buffer is complete
            w.send[self.rank] = (send_ptr, sc)
            w.barrier.wait()
            recv = self._view(recv_ptr, sum(rc))
            ro = 0
            for q in range(W):
                sp, scq = w.send[q]
                so = sum(scq[: self.rank])
                n = scq[self.rank]
                assert n == rc[q], "exchange plan mismatch"
                if n:
                    recv[ro:ro + n].copy_(self._view(sp, sum(scq))[so:so + n])
                ro += n
            self.torch.cuda.synchronize()
            w.barrier.wait()                            # nobody reuses a send buffer before it was read
            self.exchanges += 1
            return 0
        except Exception as e:                          # never let an exception cross the C boundary
            print("in-process exchange failed:", repr(e))
            self.world.barrier.abort()
            return 1

    def _allreduce(self, user, vals, n):
        try:
            w = self.world
            host = np.ctypeslib.as_array(vals, shape=(n,))
            w.red[self.rank] = host.copy()
            w.barrier.wait()
            tot = np.zeros(n)
            for q in range(w.W):                        # rank order on every rank: same bits everywhere
                tot += w.red[q]
            w.barrier.wait()
            host[:] = tot
            self.allreduces += 1
            return 0
        except Exception as e:
            print("in-process allreduce failed:", repr(e))
            self.world.barrier.abort()
            return 1


class InProcWorld:
    """W sharded handles of one graph, driven in lock step from W threads."""

    def __init__(self, graph, num_nodes, W, algorithm="hash", **opts):
        self.W = W
        self.barrier = threading.Barrier(W)
        self.send = [None] * W
        self.red = [None] * W
        self.ranks = [_Rank(self, r, graph, num_nodes, D.Options(**opts), algorithm) for r in range(W)]
        self.N, self.d = graph.num_poses, graph.d

    def _all(self, fn):
        out, err = [None] * self.W, []

        def run(r):
            try:
                out[r] = fn(self.ranks[r])
            except Exception as e:                      # noqa: BLE001
                err.append(e)
                self.barrier.abort()
        ts = [threading.Thread(target=run, args=(r,)) for r in range(self.W)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        if err:
            raise err[0]
        return out

    def run(self, X0, iters):
        """initialize / update, then `iters` x (iterate, communicate, update).  Returns
        (trace of F summed over ranks, assembled global X, per-node fobj of the last update)."""
        trace = []

        def start(rk):
            assert rk.drv.initialize(X0) == 0
            L.check(rk.drv.update())
            return rk.drv.objective()[0]

        def step(rk):
            L.check(rk.drv.iterate())
            L.check(rk.drv.communicate())
            L.check(rk.drv.update())
            return rk.drv.objective()[0]
        trace.append(sum(self._all(start)))
        for _ in range(iters):
            trace.append(sum(self._all(step)))
        X = np.zeros(((self.d + 1) * self.N, self.d), order="F")
        for rk in self.ranks:
            X += rk.drv.X()
        return np.array(trace), X

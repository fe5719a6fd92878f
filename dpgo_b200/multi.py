"""Multi-GPU sharding of the robot nodes: one process per GPU, torch.distributed for
the plumbing (NCCL on GPUs; the same code runs over gloo for CPU-side tests of the
exchange plan).

Rank r owns the contiguous nodes [r*P/W, (r+1)*P/W) (robot nodes are contiguous pose-id
ranges, C++/DPGO/src/DPGO_utils.cpp:147-158).  Per iteration only the boundary poses of
inter-node loop closures cross NVLink (all_to_all_single of packed (d+1) x d blocks,
the wire format of DPGOHash::receive, C++/DPGO/src/DPGOHash.cpp:45-82), plus one small
all_reduce per global objective evaluation of AMM-PGO* (C++/DPGO/src/DPGOStar.cpp:147-171).
AMM-PGO# needs no scalar collective.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib as L
from .pgo import DPGOHash, DPGOStar, Options


def rank_node_begin(num_nodes, world):
    return np.array([r * num_nodes // world for r in range(world + 1)], dtype=np.int32)


def plan_halo(graph, num_nodes, world, rank):
    """Host-only exchange plan (mmpgo_plan_halo): (send_counts, recv_counts, send_gids,
    recv_gids) for `rank`."""
    lib = L.load()
    rnb = rank_node_begin(num_nodes, world)
    sc = np.zeros(world, dtype=np.int64)
    rc = np.zeros(world, dtype=np.int64)
    lp = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))
    L.check(lib.mmpgo_plan_halo(graph.num_poses, num_nodes, graph.num_edges, L.iptr(graph.i),
                                L.iptr(graph.j), world, L.iptr(rnb), rank, lp(sc), lp(rc),
                                None, 0, None, 0))
    sg = np.zeros(max(int(sc.sum()), 1), dtype=np.int64)
    rg = np.zeros(max(int(rc.sum()), 1), dtype=np.int64)
    L.check(lib.mmpgo_plan_halo(graph.num_poses, num_nodes, graph.num_edges, L.iptr(graph.i),
                                L.iptr(graph.j), world, L.iptr(rnb), rank, lp(sc), lp(rc),
                                lp(sg), len(sg), lp(rg), len(rg)))
    return sc, rc, sg[: int(sc.sum())], rg[: int(rc.sum())]


def plan_halo_pair(send_counts, recv_counts, n_own):
    """Host-only index maps of the two-array exchange (mmpgo_plan_halo_pair): (send_a, send_b,
    recv_a, recv_b, halo_row), in poses."""
    lib = L.load()
    sc = np.ascontiguousarray(send_counts, dtype=np.int64)
    rc = np.ascontiguousarray(recv_counts, dtype=np.int64)
    ns, nr = max(int(sc.sum()), 1), max(int(rc.sum()), 1)
    out = [np.zeros(ns, dtype=np.int32), np.zeros(ns, dtype=np.int32), np.zeros(nr, dtype=np.int32),
           np.zeros(nr, dtype=np.int32), np.zeros(nr, dtype=np.int32)]
    lp = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))
    L.check(lib.mmpgo_plan_halo_pair(len(sc), lp(sc), lp(rc), int(n_own), *[L.iptr(a) for a in out]))
    return [a[: int(sc.sum())] for a in out[:2]] + [a[: int(rc.sum())] for a in out[2:]]


class _DevArray:
    """Zero-copy view of a raw device pointer for torch.as_tensor."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False),
                                         "version": 2}


class ShardedDriver:
    """A DPGOHash / DPGOStar over the nodes of this rank, with the NCCL transport bound."""

    def __init__(self, graph, num_nodes, options, algorithm, rank, world, native_nccl=True):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world = rank, world
        rnb = rank_node_begin(num_nodes, world)
        cls = DPGOStar if algorithm == "star" else DPGOHash
        self.drv = cls(graph, num_nodes, options, int(rnb[rank]), int(rnb[rank + 1]))
        self.exchanges = 0
        self.allreduces = 0
        self.exchange_bytes = 0
        # keep the callbacks alive for the lifetime of the handle
        self._ex = L.EXCHANGE_FN(self._exchange)
        self._ar = L.ALLREDUCE_FN(self._allreduce)
        L.check(self.drv.lib.mmpgo_set_sharding(self.drv._h, rank, world, L.iptr(rnb), self._ex, self._ar,
                                                None))
        self._scratch = torch.zeros(16, dtype=torch.float64, device="cuda") if world > 1 else None
        # collectives are issued with the library's stream current: ProcessGroupNCCL orders its
        # own stream after / before it with events, so no host synchronisation is needed
        self._ext = torch.cuda.ExternalStream(self.drv.stream()) if world > 1 else None
        self._views = {}
        self._ard = L.ALLREDUCE_DEV_FN(self._allreduce_dev)
        self.transport = "callbacks"
        if world > 1:
            L.check(self.drv.lib.mmpgo_set_device_allreduce(self.drv._h, self._ard))
            if dist.get_backend() == "nccl" and native_nccl:
                # the library drives NCCL itself (grouped send / recv + all-reduce on its own stream, from C++);
                # torch.distributed only carries the 128-byte communicator id to the ranks
                idb = (C.c_ubyte * 128)()
                if rank == 0:
                    L.check(self.drv.lib.mmpgo_nccl_unique_id(idb))
                t = torch.tensor(list(idb), dtype=torch.uint8, device="cuda")
                dist.broadcast(t, 0)
                idb = (C.c_ubyte * 128)(*t.cpu().tolist())
                # NCCL may print its version banner on stdout when a communicator is created: send it to stderr
                import os, sys
                sys.stdout.flush()
                saved = os.dup(1)
                os.dup2(2, 1)
                try:
                    rc = self.drv.lib.mmpgo_nccl_init(self.drv._h, idb)
                finally:
                    os.dup2(saved, 1)
                    os.close(saved)
                L.check(rc)
                self.transport = "nccl (C++)"

    def _view(self, ptr, n):
        """torch view of n device doubles at ptr (cached: the buffers of a handle are fixed)."""
        key = (ptr, n)
        t = self._views.get(key)
        if t is None:
            dev = self.torch.device("cuda", self.torch.cuda.current_device())
            t = self.torch.as_tensor(_DevArray(ptr, n), device=dev) if n else \
                self.torch.empty(0, dtype=self.torch.float64, device=dev)
            self._views[key] = t
        return t

    # ---- transport callbacks (called from inside libmmpgo, ordered on its stream) ----------
    def _exchange(self, user, send_ptr, send_counts, recv_ptr, recv_counts):
        try:
            torch, dist = self.torch, self.dist
            W = self.world
            sc = [int(send_counts[q]) for q in range(W)]
            rc = [int(recv_counts[q]) for q in range(W)]
            send, recv = self._view(send_ptr, sum(sc)), self._view(recv_ptr, sum(rc))
            with torch.cuda.stream(self._ext):
                dist.all_to_all_single(recv, send, rc, sc)
            self.exchanges += 1
            self.exchange_bytes += 8 * sum(sc)
            return 0
        except Exception as e:          # never let an exception cross the C boundary
            print("mmpgo exchange callback failed:", repr(e))
            return 1

    def _allreduce(self, user, vals, n):
        try:
            torch, dist = self.torch, self.dist
            host = np.ctypeslib.as_array(vals, shape=(n,))
            t = self._scratch[:n]
            with torch.cuda.stream(self._ext):
                t.copy_(torch.from_numpy(host))
                dist.all_reduce(t)
                host[:] = t.cpu().numpy()
            self.allreduces += 1
            return 0
        except Exception as e:
            print("mmpgo allreduce callback failed:", repr(e))
            return 1

    def _allreduce_dev(self, user, ptr, n):
        try:
            with self.torch.cuda.stream(self._ext):
                self.dist.all_reduce(self._view(ptr, n))
            self.allreduces += 1
            return 0
        except Exception as e:
            print("mmpgo device allreduce callback failed:", repr(e))
            return 1

    # ---- driver interface ------------------------------------------------------------
    def close(self):
        """Release the handle (and its NCCL communicator) now, at a point all ranks reach together, instead of
        whenever the garbage collector finds the callback cycle."""
        if self.world > 1:
            self.torch.cuda.synchronize()
            self.dist.barrier()
        self.drv.close()

    def __getattr__(self, name):
        return getattr(self.drv, name)

    def global_objective(self):
        """(F, |grad F|) summed over all ranks."""
        f, g = self.drv.objective()
        if self.world == 1:
            return f, g
        t = self.torch.tensor([f, g * g], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t)
        return float(t[0]), float(np.sqrt(t[1].item()))

    def solve_kernels_per_iter(self):
        return 3.0

    def stage_range(self):
        return self.drv.stage_range()

    def halo_counts(self):
        sc = np.zeros(self.world, dtype=np.int64)
        rc = np.zeros(self.world, dtype=np.int64)
        lp = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))
        L.check(self.drv.lib.mmpgo_halo_counts(self.drv._h, lp(sc), lp(rc)))
        return sc, rc


def make_driver(graph, num_nodes, options, algorithm="star", rank=0, world=1, native_nccl=True):
    return ShardedDriver(graph, num_nodes, options or Options(), algorithm, rank, world, native_nccl)


def e2e_loop(drv, X0, steps, E, d, N):
    """End-to-end arm: the loop of dist_pgo with HOST matrices (C++/examples/dist_pgo.cpp:497-530), steady state.
    Per iteration: iterate(); results().Xk of the local nodes into the caller's pinned global X (device -> host,
    :502-511); communicate(); update(); evaluate_f(X_host), the logging call dist_pgo makes on the host iterate
    (:523-524; host -> device: the rows of the local nodes and of their remote neighbours).  The solver state is
    not reset between steps.  At N > 1 every rank holds its own host matrix: its rows are current, the rows of
    remote neighbours are those of the start (the value of the logging call is not used, its traffic is)."""
    import time
    torch = drv.torch
    dist = drv.dist if drv.world > 1 else None
    pin = torch.empty((d, (d + 1) * N), dtype=torch.float64).pin_memory()
    Xh = pin.numpy().T
    Xh[:] = X0

    def sync():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
    assert drv.initialize(Xh) == 0 and drv.update() == 0
    for _ in range(3):                                   # reach the steady state of the Nesterov sequence
        L.check(drv.iterate()); L.check(drv.communicate()); L.check(drv.update())
    drv.X(out=Xh)
    if dist is not None:                                 # every rank starts from the assembled iterate
        t = torch.from_numpy(np.ascontiguousarray(Xh)).cuda()
        t[:] = 0
        own = drv.X()
        t += torch.from_numpy(np.ascontiguousarray(own)).cuda()
        dist.all_reduce(t)
        Xh[:] = t.cpu().numpy()
    drv.evaluate_f(Xh)
    sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        L.check(drv.iterate())
        drv.X(out=Xh)
        L.check(drv.communicate())
        L.check(drv.update())
        drv.evaluate_f(Xh)
    drv.synchronize()
    sync()
    dt = time.perf_counter() - t0
    lo, hi = drv.stage_range()
    up = np.array([float((hi - lo) * (d + 1) * d * 8), float(drv.sizes()["own_poses"] * (d + 1) * d * 8), dt])
    if dist is not None:
        t = torch.tensor(up, dtype=torch.float64, device="cuda")
        tm = t.clone()
        dist.all_reduce(t)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        up = np.array([float(t[0]), float(t[1]), float(tm[2])])
    return {"value": E * steps / up[2], "unit": "edge-updates/s", "h2d_bytes_per_step": int(up[0]),
            "d2h_bytes_per_step": int(up[1]), "steps": steps,
            "call_sequence": "iterate(); X(out=X_host); communicate(); update(); evaluate_f(X_host)  "
                             "[dist_pgo.cpp:497-530, steady state, solver state kept]",
            "host_buffers": "pinned", "bytes": "summed over ranks"}

"""Multi-process runner of the oracle's AMM-PGO* driver (TEST INFRASTRUCTURE / CPU baseline only).

The reference loops over the robot nodes in one process (`for alpha: ...->iterate()`,
C++/examples/dist_pgo.cpp:497-520) and gets its parallelism from OpenMP inside Eigen's
products and the SO(d) projection.  The numpy/scipy restatement holds the GIL in most of its
per-node work, so to time it "with all the host threads it can use" (bench.py, reference arm)
the per-node parts of DPGOStar::update / iterate / communicate (DPGOStar.cpp:126-390) run
in forked worker processes that each own a subset of the nodes; the global iterates X^k,
X^{k+1/2}, X^{k+1} live in shared memory; the master keeps the global objective evaluations and
the restart decisions.  The arithmetic per node is that of `oracle.dpgo.DPGOStar`, so the
iterates are bit-identical to the serial oracle (tests/test_oracle_parallel.py).
"""
from __future__ import annotations

import multiprocessing as mp

import numpy as np

from . import dpgo


def _shared_like(X):
    buf = mp.RawArray("d", int(X.size))
    A = np.frombuffer(buf, dtype=np.float64).reshape(X.shape)
    A[...] = X
    return A


class ParallelDPGOStar(dpgo.DPGOStar):
    """DPGOStar with the per-node work spread over `workers` forked processes."""

    def __init__(self, *args, workers=2, **kw):
        super().__init__(*args, **kw)
        self.workers = max(1, min(int(workers), self.num_nodes))
        self._pipes, self._procs = [], []
        self.timeout = 900.0

    # ---- master side ------------------------------------------------------------------
    def initialize(self, X):
        self.close()
        super().initialize(X)
        self.Xk, self.Xkh, self.Xkp = _shared_like(self.Xk), _shared_like(self.Xkh), _shared_like(self.Xkp)
        ctx = mp.get_context("fork")
        for w in range(self.workers):
            parent, child = ctx.Pipe()
            nodes = list(range(w, self.num_nodes, self.workers))
            pr = ctx.Process(target=self._serve, args=(child, nodes), daemon=True)
            pr.start()
            child.close()
            self._pipes.append(parent)
            self._procs.append(pr)
        return 0

    def _all(self, cmd):
        for c in self._pipes:
            c.send(cmd)
        out = {}
        for c, pr in zip(self._pipes, self._procs):
            waited = 0.0
            while not c.poll(1.0):                       # never block forever on a dead worker
                waited += 1.0
                if not pr.is_alive() or waited > self.timeout:
                    raise RuntimeError("oracle worker process died or timed out")
            r = c.recv()
            if isinstance(r, Exception):
                raise r
            out.update(r)
        return out

    def update(self):
        for a, (fobj, gn) in self._all("update").items():
            self.results[a].fobj_cur, self.results[a].gradFnorm = fobj, gn
        return 0

    def iterate(self):
        o = self.opts
        self.n_global_restarts = getattr(self, "n_global_restarts", 0)
        for a, refined in self._all("amm").items():
            self.results[a].last_refined = refined
        fobjh = self.gobj.evaluate_f(self.Xkh)
        if fobjh > self.F - o.psi * float(np.sum(np.square(self.Xkh - self.Xk))):
            self._all("pm")
            fobjh = self.gobj.evaluate_f(self.Xkh)
        fobj = self.gobj.evaluate_f(self.Xkp)
        if fobj > self.F - o.psi * float(np.sum(np.square(self.Xkp - self.Xk))):
            self.n_global_restarts += 1
            self._all("restart")
            fobj = self.gobj.evaluate_f(self.Xkp)
        if self.F - fobj < o.phi * (self.F - fobjh):
            self._all("safeguard")
            fobj = self.gobj.evaluate_f(self.Xkp)
        self._all("finish")
        self.Xk, self.Xkp = self.Xkp, self.Xk
        self.fobj = fobj
        self.F = self.F * (1 - o.eta[0]) + fobj * o.eta[0]
        return 0

    def communicate(self):
        self._all("communicate")
        return 0

    def close(self):
        for c in self._pipes:
            try:
                c.send("stop")
            except Exception:
                pass
        for pr in self._procs:
            pr.join(timeout=10)
        self._pipes, self._procs = [], []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- worker side (forked: owns the NodeState of `nodes`, shares the global iterates) ------
    def _serve(self, conn, nodes):
        per_node = {"pm": self._pm_pgo_n, "restart": self._restart_n, "safeguard": self._safeguard_n,
                    "communicate": self._communicate_n}
        while True:
            try:
                cmd = conn.recv()
            except EOFError:
                return
            if cmd == "stop":
                return
            try:
                out = {}
                if cmd == "update":
                    for a in nodes:
                        self._update_n(a)
                        out[a] = (self.results[a].fobj_cur, self.results[a].gradFnorm)
                elif cmd == "amm":
                    for a in nodes:
                        self._amm_pgo_n(a)
                        out[a] = self.results[a].last_refined
                elif cmd == "finish":
                    for a in nodes:
                        self._finish_n(a)
                    self.Xk, self.Xkp = self.Xkp, self.Xk       # same swap as the master
                else:
                    for a in nodes:
                        per_node[cmd](a)
                conn.send(out)
            except Exception as e:                                # surfaced in the master
                conn.send(e)

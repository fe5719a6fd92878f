"""ctypes driver of oracle/cpu_dpgo.cpp (TEST INFRASTRUCTURE / CPU BASELINE ONLY).

The per-node sparse matrices are assembled by oracle/data_matrix.py (the restatement of the reference's
builders, DPGO_utils.cpp:1398-2967), handed to the C++ restatement of the iteration as CSR arrays, and the
outer loop of dist_pgo (C++/examples/dist_pgo.cpp:446-531) is driven from here.  Setup (matrix assembly,
factorisation of G00) is outside every timed region, as in dist_pgo (:496-521 time iterate and update only).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import data_matrix as dm
from . import g2o

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libcpu_dpgo.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_long)
LOSS = {"trivial": 0, "huber": 1, "gm": 2, "welsch": 3}
PRECON = {"None": 0, "Jacobi": 1, "BlockJacobi": 2}


def available():
    return os.path.exists(LIB)


def build():
    """Needs /root/reference (links the reference's AVX2 projection kernels): authoring container only."""
    subprocess.check_call(["make", "-C", HERE, "_ref/libcpu_dpgo.so"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(LIB)
        l.cpu_dpgo_create.restype = C.c_void_p
        l.cpu_dpgo_create.argtypes = [C.c_int, C.c_int, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, _dp]
        l.cpu_dpgo_destroy.argtypes = [C.c_void_p]
        l.cpu_dpgo_set_matrix.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_int, _ip, _ip, _dp]
        l.cpu_dpgo_set_node.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_long, _lp, _lp, _dp, _ip, _ip, _dp, _ip]
        l.cpu_dpgo_set_edges.argtypes = [C.c_void_p, C.c_long, _lp, _lp, _dp, _dp, _dp, _dp, C.POINTER(C.c_ubyte)]
        l.cpu_dpgo_set_threads.argtypes = [C.c_void_p, C.c_int, C.c_int]
        for f in (l.cpu_dpgo_update, l.cpu_dpgo_iterate, l.cpu_dpgo_communicate, l.cpu_dpgo_star_restarts):
            f.argtypes = [C.c_void_p]
        l.cpu_dpgo_initialize.argtypes = [C.c_void_p, _dp]
        l.cpu_dpgo_get_X.argtypes = [C.c_void_p, _dp]
        l.cpu_dpgo_node_scalars.argtypes = [C.c_void_p, _dp]
        l.cpu_dpgo_evaluate_f.argtypes = [C.c_void_p, _dp]
        l.cpu_dpgo_evaluate_f.restype = C.c_double
        l.cpu_dpgo_weights.argtypes = [C.c_void_p, C.c_int, _dp]
        _lib = l
    return _lib


def _needed(algorithm, quadratic):
    names = ["G", "G01", "G10", "G11", "N"]
    if quadratic:
        names += ["S", "P0", "U"] + (["P", "Q"] if algorithm == "hash" else [])
    else:
        names += ["D", "B1", "V", "Q"]
    return names


def _node_payload(args):
    """Everything one node hands to the C++ side (runs in a worker process)."""
    a, meas_a, xi, quadratic, names = args
    info = dm.generate_data_info(a, meas_a)
    mats = dm.build_data_matrices(info, xi, quadratic)
    out = {"n": info.n, "m": info.m, "own_poses": info.own_poses, "nbr_keys": info.nbr_keys, "T": mats["T"]}
    for k in names + ["G00"]:
        M = sp.csr_matrix(mats[k])
        M.sort_indices()
        out[k] = (M.shape, M.indptr.astype(np.int32), M.indices.astype(np.int32), M.data.astype(np.float64))
    # fill-reducing ordering for the Cholesky factor of G00 (the reference: CHOLMOD's own AMD)
    lu = spla.splu(sp.csc_matrix(mats["G00"]), permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0,
                   options=dict(SymmetricMode=True))
    # SuperLU's perm_c maps an original column to its position; the factorisation wants position -> original
    out["perm"] = np.argsort(lu.perm_c).astype(np.int32)
    return a, out


class CpuDPGO:
    """DPGOHash (all nodes) / DPGOStar of the restated C++ reference.  Same driver methods as the reference:
    initialize / update / iterate / communicate."""

    def __init__(self, meas, num_poses, num_nodes, opts, algorithm="star", workers=None, threads=None, mode=0):
        self.l = lib()
        self.d, self.N, self.A = meas.d, num_poses, num_nodes
        self.algorithm = algorithm
        per_node, g_index, part = g2o.partition(num_poses, num_nodes, meas)
        quadratic = opts.loss == "trivial"
        o = np.array([opts.regularizer, opts.loss_reg, opts.accepted_delta, opts.eta[0], opts.eta[1], opts.psi, opts.phi,
                      opts.max_soft_restart_hits[0], opts.max_soft_restart_hits[1], opts.oscillation_cnt_period,
                      opts.max_oscillations, opts.grad_norm_tol, opts.preconditioned_grad_norm_tol,
                      opts.rel_func_decrease_tol, opts.stepsize_tol, opts.max_iterations, opts.max_iterations_accepted,
                      opts.max_tCG_iterations, opts.STPCG_kappa, opts.STPCG_theta], dtype=np.float64)
        self.h = self.l.cpu_dpgo_create(self.d, num_nodes, num_poses, 1 if algorithm == "star" else 0,
                                        1 if opts.scheme == "AMM" else 0, LOSS[opts.loss], PRECON[opts.preconditioner],
                                        o.ctypes.data_as(_dp))
        names = _needed(algorithm, quadratic)
        jobs = [(a, per_node[a], opts.regularizer, quadratic, names) for a in range(num_nodes)]
        workers = workers if workers is not None else min(os.cpu_count() or 1, num_nodes, 16)
        if workers > 1:
            import multiprocessing as mp
            with mp.get_context("fork").Pool(workers) as pool:
                results = list(pool.imap_unordered(_node_payload, jobs, chunksize=1))
        else:
            results = [_node_payload(j) for j in jobs]
        results.sort(key=lambda r: r[0])
        BIG = np.int64(1) << 40
        own_gid = []
        for a, out in results:
            own_gid.append(np.array([g_index[a][int(q)] for q in out["own_poses"]], dtype=np.int64))
        for a, out in results:
            for k in names:
                shape, ptr, idx, val = out[k]
                rc = self.l.cpu_dpgo_set_matrix(self.h, a, k.encode(), shape[0], shape[1], ptr.ctypes.data_as(_ip),
                                                idx.ctypes.data_as(_ip), val.ctypes.data_as(_dp))
                assert rc == 0, k
            keys = out["nbr_keys"]
            nn, npose = (keys >> 40).astype(np.int64), (keys & (BIG - 1)).astype(np.int64)
            nbr_gid = np.empty(len(keys), dtype=np.int64)
            for b in np.unique(nn):
                sel = nn == b
                loc = np.searchsorted(results[int(b)][1]["own_poses"], npose[sel])
                nbr_gid[sel] = own_gid[int(b)][loc]
            shape, ptr, idx, val = out["G00"]
            T = np.ascontiguousarray(out["T"], dtype=np.float64)
            og = np.ascontiguousarray(own_gid[a])
            rc = self.l.cpu_dpgo_set_node(self.h, a, out["n"][0], out["n"][1], out["m"][1], int(og[0]),
                                          og.ctypes.data_as(_lp), nbr_gid.ctypes.data_as(_lp), T.ctypes.data_as(_dp),
                                          ptr.ctypes.data_as(_ip), idx.ctypes.data_as(_ip), val.ctypes.data_as(_dp),
                                          out["perm"].ctypes.data_as(_ip))
            assert rc == 0, "G00 of node %d is not positive definite" % a
        inter = (part.i_node != part.j_node).astype(np.uint8)
        i, j = np.ascontiguousarray(meas.i_pose, dtype=np.int64), np.ascontiguousarray(meas.j_pose, dtype=np.int64)
        R, t = np.ascontiguousarray(meas.R, dtype=np.float64), np.ascontiguousarray(meas.t, dtype=np.float64)
        ka, ta = np.ascontiguousarray(meas.kappa, dtype=np.float64), np.ascontiguousarray(meas.tau, dtype=np.float64)
        self.l.cpu_dpgo_set_edges(self.h, len(i), i.ctypes.data_as(_lp), j.ctypes.data_as(_lp), R.ctypes.data_as(_dp),
                                  t.ctypes.data_as(_dp), ka.ctypes.data_as(_dp), ta.ctypes.data_as(_dp),
                                  inter.ctypes.data_as(C.POINTER(C.c_ubyte)))
        self.threads = threads or (os.cpu_count() or 1)
        self.l.cpu_dpgo_set_threads(self.h, mode, self.threads)

    def close(self):
        if getattr(self, "h", None):
            self.l.cpu_dpgo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def initialize(self, X):
        X = np.asfortranarray(X, dtype=np.float64)
        assert X.shape == ((self.d + 1) * self.N, self.d)
        return self.l.cpu_dpgo_initialize(self.h, X.ctypes.data_as(_dp))

    def update(self):
        return self.l.cpu_dpgo_update(self.h)

    def iterate(self):
        return self.l.cpu_dpgo_iterate(self.h)

    def communicate(self):
        return self.l.cpu_dpgo_communicate(self.h)

    def X(self):
        X = np.zeros(((self.d + 1) * self.N, self.d), order="F")
        self.l.cpu_dpgo_get_X(self.h, X.ctypes.data_as(_dp))
        return X

    def node_scalars(self):
        """(A, 5): fobj, gradFnorm, refined, tCG iterations, restarts."""
        out = np.zeros((self.A, 5))
        self.l.cpu_dpgo_node_scalars(self.h, out.ctypes.data_as(_dp))
        return out

    def evaluate_f(self, X):
        X = np.asfortranarray(X, dtype=np.float64)
        return self.l.cpu_dpgo_evaluate_f(self.h, X.ctypes.data_as(_dp))

    def weights(self, node, m1):
        w = np.zeros(max(m1, 1))
        n = self.l.cpu_dpgo_weights(self.h, node, w.ctypes.data_as(_dp))
        return w[:n]


def run(meas, num_poses, num_nodes, opts, X0, iters, algorithm="star", timing=None, **kw):
    """The loop of oracle.dist_pgo.run on the C++ restatement: returns dict(fobj_nodes, refined, X, seconds)."""
    drv = CpuDPGO(meas, num_poses, num_nodes, opts, algorithm, **kw)
    out = {"fobj_nodes": [], "refined": [], "step_seconds": []}
    drv.initialize(X0)
    t_acc = 0.0
    t0 = time.perf_counter()
    drv.update()
    t_acc += time.perf_counter() - t0
    out["fobj_nodes"].append(drv.node_scalars()[:, 0].copy())
    for _ in range(iters):
        t0 = time.perf_counter()
        drv.iterate()
        sc = drv.node_scalars()
        drv.communicate()
        drv.update()
        dt = time.perf_counter() - t0
        t_acc += dt
        out["step_seconds"].append(dt)
        out["refined"].append(sc[:, 2].astype(bool))
        out["fobj_nodes"].append(drv.node_scalars()[:, 0].copy())
    out["X"] = drv.X()
    out["seconds"] = t_acc
    out["tcg"] = drv.node_scalars()[:, 3]
    out["drv"] = drv
    if timing is not None:
        timing["seconds"] = t_acc
    return out

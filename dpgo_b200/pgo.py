"""Host-side mirror of the reference's per-node driver interface.

`DPGOHash` / `DPGOStar` keep the method names, argument meaning and int
return codes of C++/DPGO/include/DPGO/DPGOHash.h:13-107 and DPGOStar.h:13-93
(initialize / update / iterate / communicate / evaluate_f / results), but one
object drives ALL robot nodes of the graph that live on this GPU: the
reference's `for alpha in nodes: dpgo_hash[alpha]->iterate()` loops
(C++/examples/dist_pgo.cpp:497-520) collapse into one batched call through
the C ABI (include/mmpgo.h).  All arithmetic happens in libmmpgo.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib as L
from .graph import PoseGraph


class Options:
    """DPGO::Options with the values `dist_pgo` sets (dist_pgo.cpp:103-120)."""

    def __init__(self, **kw):
        self.c = L.Options()
        L.load().mmpgo_default_options(C.byref(self.c))
        self.set(**kw)

    def set(self, **kw):
        for k, v in kw.items():
            if k == "loss":
                self.c.loss = L.LOSS[v]
            elif k == "preconditioner":
                self.c.preconditioner = L.PRECON[v]
            elif k == "algorithm":
                self.c.algorithm = L.ALGORITHM[v]
            elif k == "scheme":
                self.c.scheme = L.SCHEME[v]
            elif k == "translation_solver":
                self.c.translation_solver = L.TSOLVER[v]
            elif k == "rescale":
                self.c.rescale = L.RESCALE[v]
            elif k in ("eta", "max_soft_restart_hits"):
                arr = getattr(self.c, k)
                arr[0], arr[1] = v
            else:
                if not hasattr(self.c, k):
                    raise AttributeError(k)
                setattr(self.c, k, v)
        return self


class _Driver:
    """Common part of DPGOHash / DPGOStar."""

    algorithm = "hash"

    def __init__(self, graph: PoseGraph, num_nodes: int, options: Options | None = None,
                 node_begin: int = 0, node_end: int | None = None):
        self.lib = L.load()
        self.graph = graph
        self.num_nodes = int(num_nodes)
        self.node_begin = int(node_begin)
        self.node_end = int(num_nodes if node_end is None else node_end)
        self.options = options or Options()
        self.options.set(algorithm=self.algorithm)
        self.d = graph.d
        self.N = graph.num_poses
        self._h = L._P()
        L.check(self.lib.mmpgo_create(C.byref(self.options.c), C.byref(self._h)))
        try:
            L.check(self.lib.mmpgo_set_graph(
                self._h, graph.d, graph.num_poses, self.num_nodes, self.node_begin,
                self.node_end, graph.num_edges, L.iptr(graph.i), L.iptr(graph.j),
                L.dptr(graph.R), L.dptr(graph.t), L.dptr(graph.kappa), L.dptr(graph.tau)))
        except Exception:
            self.lib.mmpgo_destroy(self._h)
            self._h = None
            raise

    def close(self):
        if getattr(self, "_h", None):
            self.lib.mmpgo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- the reference's driver methods ------------------------------------
    def initialize(self, X):
        """X: global iterate ((d+1)N x d), rows [t; R blocks] (dist_pgo.cpp:436-446)."""
        X = np.asfortranarray(X, dtype=np.float64)
        if X.shape != ((self.d + 1) * self.N, self.d):
            return -1
        return self.lib.mmpgo_initialize(self._h, L.dptr(X), X.shape[0])

    def update(self):
        return self.lib.mmpgo_update(self._h)

    def iterate(self):
        return self.lib.mmpgo_iterate(self._h)

    def communicate(self):
        return self.lib.mmpgo_communicate(self._h)

    def evaluate_f(self, X):
        """DPGOStar::evaluate_f (DPGOStar.cpp:713-761) over the edges owned here."""
        X = np.asfortranarray(X, dtype=np.float64)
        out = C.c_double()
        L.check(self.lib.mmpgo_evaluate_f(self._h, L.dptr(X), X.shape[0], C.byref(out)))
        return out.value

    def evaluate_grad(self, X):
        """DPGOStar::evaluate_grad (DPGOStar.cpp:763-829): Riemannian gradient of the global
        objective at X, in the layout of X (rows of the poses owned here; zero elsewhere)."""
        X = np.asfortranarray(X, dtype=np.float64)
        G = np.zeros_like(X, order="F")
        L.check(self.lib.mmpgo_evaluate_grad(self._h, L.dptr(X), X.shape[0], L.dptr(G), G.shape[0]))
        return G

    # -- results() ----------------------------------------------------------
    def X(self, out=None):
        """Global-layout copy of the current iterate (rows of local nodes).  `out`: an
        F-ordered ((d+1)N x d) float64 array to write into (e.g. pinned memory)."""
        X = np.zeros(((self.d + 1) * self.N, self.d), order="F") if out is None else out
        assert X.flags.f_contiguous and X.shape == ((self.d + 1) * self.N, self.d)
        L.check(self.lib.mmpgo_get_poses(self._h, L.dptr(X), X.shape[0]))
        return X

    def node_scalars(self, node):
        s = L.NodeScalars()
        L.check(self.lib.mmpgo_get_node_scalars(self._h, node, C.byref(s)))
        return s

    def weights(self, node):
        n = C.c_int64()
        L.check(self.lib.mmpgo_get_weights(self._h, node, None, 0, C.byref(n)))
        w = np.zeros(max(n.value, 1))
        L.check(self.lib.mmpgo_get_weights(self._h, node, L.dptr(w), len(w), C.byref(n)))
        return w[: n.value]

    def objective(self):
        """(F, |grad F|) of the current iterate from the per-node sums; dist_pgo
        prints 2F and 2|grad F| (dist_pgo.cpp:523-530)."""
        f, g2 = C.c_double(), C.c_double()
        L.check(self.lib.mmpgo_current_objective(self._h, C.byref(f), C.byref(g2)))
        return f.value, float(np.sqrt(g2.value))

    def counters(self):
        c = L.Counters()
        L.check(self.lib.mmpgo_get_counters(self._h, C.byref(c)))
        return c

    def reset_counters(self):
        L.check(self.lib.mmpgo_reset_counters(self._h))

    def sizes(self):
        s = (C.c_int64 * 8)()
        L.check(self.lib.mmpgo_graph_sizes(self._h, s))
        keys = ("own_poses", "halo_poses", "bsr_entries", "inter_half_edges",
                "owned_edges", "tiles", "local_nodes", "d")
        return dict(zip(keys, list(s)))

    def translation_solve(self, rhs):
        """t = -G00^{-1} rhs for every local node (the L_.solve of recover_translations,
        DPGOProblem.h:291); rhs: (own poses, d)."""
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        t = np.zeros_like(rhs)
        L.check(self.lib.mmpgo_translation_solve(self._h, L.dptr(rhs), L.dptr(t)))
        return t

    def preconditioner_info(self, node):
        """(lambda_max estimate of the node's G11, entries of the factor of all local G11 + reg I) of the
        RegularizedCholesky preconditioner (DPGOProblem.cpp:101-124)."""
        lam, nnz = C.c_double(), C.c_int64()
        L.check(self.lib.mmpgo_preconditioner_info(self._h, node, C.byref(lam), C.byref(nnz)))
        return lam.value, nnz.value

    def stage_range(self):
        """[lo, hi): the global pose ids whose rows initialize / evaluate_f / evaluate_grad copy to the device
        (the local nodes' own poses and their remote neighbours)."""
        lo, hi = C.c_int64(), C.c_int64()
        L.check(self.lib.mmpgo_stage_range(self._h, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def solver_info(self):
        s = (C.c_int64 * 8)()
        L.check(self.lib.mmpgo_solver_info(self._h, s))
        keys = ("solver", "factor_nnz", "factor_entries", "tree_height", "supernodes", "tasks_per_solve",
                "persistent_ctas", "dense_poses")
        out = dict(zip(keys, list(s)))
        out["solver"] = {v: k for k, v in L.TSOLVER.items()}[out["solver"]] if out["solver"] else "dense"
        return out

    def solver_stage_times(self):
        """[(microseconds, warp jobs, CTA jobs)] per stage of the last sparse direct solve."""
        n = C.c_int32()
        L.check(self.lib.mmpgo_solver_stage_times(self._h, None, None, None, 0, C.byref(n)))
        if n.value == 0:
            return []
        us, wj, cj = np.zeros(n.value), np.zeros(n.value, dtype=np.int32), np.zeros(n.value, dtype=np.int32)
        L.check(self.lib.mmpgo_solver_stage_times(self._h, L.dptr(us), L.iptr(wj), L.iptr(cj), n.value, C.byref(n)))
        return list(zip(us.tolist(), wj.tolist(), cj.tolist()))

    KERNEL_KINDS = {"k2_eval": 0, "k2_grad": 1, "k1_inter": 2, "k3_prox": 3, "edge_objective": 4,
                    "k2_hv": 6, "k2_g01": 7, "g00_solve": 8, "g00_solve_dry": 9, "g00_solve_levels": 10, "g00_solve_deps": 11}

    def profile_pass(self, kind, reps=20):
        """Average device milliseconds of one launch of a hot kernel (CUDA events on
        the library's stream)."""
        ms = C.c_float()
        L.check(self.lib.mmpgo_profile_pass(self._h, self.KERNEL_KINDS[kind], reps, C.byref(ms)))
        return ms.value

    def synchronize(self):
        L.check(self.lib.mmpgo_synchronize(self._h))

    def stream(self):
        return self.lib.mmpgo_stream(self._h)


class DPGOHash(_Driver):
    """AMM-PGO# (scheme AMM) / MM-PGO (scheme MM), decentralised restarts."""
    algorithm = "hash"


class DPGOStar(_Driver):
    """AMM-PGO*, restart decided on the global objective by the master node."""
    algorithm = "star"

    def star_objective(self):
        F, f, r = C.c_double(), C.c_double(), C.c_int32()
        L.check(self.lib.mmpgo_star_objective(self._h, C.byref(F), C.byref(f), C.byref(r)))
        return F.value, f.value, r.value


def project_to_SOdn(A, device=0):
    """project_to_SO3n / project_to_SO2n (C++/DPGO/include/DPGO/DPGO_utils.h:515-565) on the
    device: the polar projection of every row-major d x d block of A (n, d, d)."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    U = np.empty_like(A)
    L.check(L.load().mmpgo_project_to_sodn(A.shape[1], A.shape[0], L.dptr(A), L.dptr(U), device))
    return U


def run_dist_pgo(graph, num_nodes, X0, iters, options=None, algorithm="hash", log=True):
    """The outer loop of `dist_pgo` (dist_pgo.cpp:446-531) on one GPU.
    Returns (driver, trace) with trace[k] = (2F, 2|grad F|) before iteration k."""
    cls = DPGOStar if algorithm == "star" else DPGOHash
    drv = cls(graph, num_nodes, options)
    rc = drv.initialize(X0)
    if rc:
        raise L.MmpgoError(rc, L.load().mmpgo_last_error().decode())
    L.check(drv.update())
    trace = []
    for it in range(iters):
        if log:
            f, g = drv.objective()
            trace.append((2 * f, 2 * g))
        L.check(drv.iterate())
        L.check(drv.communicate())
        L.check(drv.update())
    f, g = drv.objective()
    trace.append((2 * f, 2 * g))
    return drv, trace

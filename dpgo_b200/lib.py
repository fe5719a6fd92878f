"""ctypes binding of libmmpgo.so (the C ABI declared in include/mmpgo.h).

The library is the product: there is no CPU fallback.  Loading fails loudly
when the shared object is missing, and every call fails loudly when no CUDA
device is present (mmpgo_create returns MMPGO_ERR_CUDA).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmmpgo.so")


class MmpgoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("mmpgo error %d: %s" % (code, msg))
        self.code = code


class Options(C.Structure):
    """struct mmpgo_options (include/mmpgo.h)."""
    _fields_ = [
        ("algorithm", C.c_int32), ("scheme", C.c_int32), ("loss", C.c_int32),
        ("preconditioner", C.c_int32),
        ("regularizer", C.c_double), ("loss_reg", C.c_double),
        ("accepted_delta", C.c_double), ("eta", C.c_double * 2),
        ("psi", C.c_double), ("phi", C.c_double),
        ("max_soft_restart_hits", C.c_int32 * 2),
        ("oscillation_cnt_period", C.c_int32), ("max_oscillations", C.c_int32),
        ("grad_norm_tol", C.c_double), ("preconditioned_grad_norm_tol", C.c_double),
        ("rel_func_decrease_tol", C.c_double), ("stepsize_tol", C.c_double),
        ("max_iterations", C.c_int32), ("max_iterations_accepted", C.c_int32),
        ("max_tCG_iterations", C.c_int32),
        ("STPCG_kappa", C.c_double), ("STPCG_theta", C.c_double),
        ("dense_solve_max_n", C.c_int32),
        ("translation_solve_tol", C.c_double),
        ("translation_solve_max_iters", C.c_int32), ("device", C.c_int32),
        ("translation_solver", C.c_int32), ("rescale", C.c_int32), ("max_rescale_count", C.c_int32),
        ("reg_Cholesky_precon_max_condition_number", C.c_double), ("reserved", C.c_int32 * 2),
    ]


class NodeScalars(C.Structure):
    """struct mmpgo_node_scalars."""
    _fields_ = [
        ("fobj", C.c_double), ("f", C.c_double), ("Gk", C.c_double),
        ("gradFnorm", C.c_double), ("Fk", C.c_double * 2), ("s", C.c_double),
        ("s_next", C.c_double), ("gamma", C.c_double),
        ("iters", C.c_int32), ("soft_restart_hits", C.c_int32 * 2),
        ("num_oscillations", C.c_int32), ("refined", C.c_int32),
        ("restarts", C.c_int32), ("tcg_iterations", C.c_int32),
        ("tnt_iterations", C.c_int32), ("n0", C.c_int32), ("n1", C.c_int32),
        ("m0", C.c_int32), ("m1", C.c_int32),
        ("translation_solve_iters", C.c_int32), ("reserved", C.c_int32),
    ]


class Counters(C.Structure):
    """struct mmpgo_counters."""
    _fields_ = [
        ("launches", C.c_int64), ("intra_passes", C.c_int64),
        ("inter_passes", C.c_int64), ("prox_passes", C.c_int64),
        ("solve_calls", C.c_int64), ("solve_iters", C.c_int64),
        ("tcg_iterations", C.c_int64), ("tnt_iterations", C.c_int64),
        ("vector_passes", C.c_int64), ("reserved", C.c_int64 * 7),
    ]


LOSS = {"trivial": 0, "none": 0, "huber": 1, "gm": 2, "geman-mcclure": 2, "welsch": 3}
PRECON = {"None": 0, "Jacobi": 1, "BlockJacobi": 2, "RegularizedCholesky": 3}
ALGORITHM = {"hash": 0, "star": 1}
SCHEME = {"MM": 0, "AMM": 1}
RESCALE = {"Static": 0, "Dynamic": 1}
TSOLVER = {"auto": 0, "pcg": 1, "pcg_ring": 2, "pcg_lite": 3, "direct": 4}

# every symbol include/mmpgo.h declares, with its signature
_P = C.c_void_p
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)
EXCHANGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, _lp, C.c_void_p, _lp)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_int32)
ALLREDUCE_DEV_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int32)
SIGNATURES = {
    "mmpgo_set_sharding": (C.c_int, [_P, C.c_int32, C.c_int32, _ip, EXCHANGE_FN, ALLREDUCE_FN, C.c_void_p]),
    "mmpgo_set_device_allreduce": (C.c_int, [_P, ALLREDUCE_DEV_FN]),
    "mmpgo_halo_counts": (C.c_int, [_P, _lp, _lp]),
    "mmpgo_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "mmpgo_nccl_init": (C.c_int, [_P, C.c_void_p]),
    "mmpgo_plan_halo": (C.c_int, [C.c_int64, C.c_int32, C.c_int64, _ip, _ip, C.c_int32, _ip, C.c_int32,
                                  _lp, _lp, _lp, C.c_int64, _lp, C.c_int64]),
    "mmpgo_plan_halo_pair": (C.c_int, [C.c_int32, _lp, _lp, C.c_int64, _ip, _ip, _ip, _ip, _ip]),
    "mmpgo_version": (C.c_char_p, []),
    "mmpgo_last_error": (C.c_char_p, []),
    "mmpgo_default_options": (None, [C.POINTER(Options)]),
    "mmpgo_create": (C.c_int, [C.POINTER(Options), C.POINTER(_P)]),
    "mmpgo_destroy": (C.c_int, [_P]),
    "mmpgo_set_graph": (C.c_int, [_P, C.c_int32, C.c_int64, C.c_int32, C.c_int32,
                                  C.c_int32, C.c_int64, _ip, _ip, _dp, _dp, _dp, _dp]),
    "mmpgo_initialize": (C.c_int, [_P, _dp, C.c_int64]),
    "mmpgo_update": (C.c_int, [_P]),
    "mmpgo_iterate": (C.c_int, [_P]),
    "mmpgo_communicate": (C.c_int, [_P]),
    "mmpgo_get_poses": (C.c_int, [_P, _dp, C.c_int64]),
    "mmpgo_get_node_scalars": (C.c_int, [_P, C.c_int32, C.POINTER(NodeScalars)]),
    "mmpgo_get_weights": (C.c_int, [_P, C.c_int32, _dp, C.c_int64, C.POINTER(C.c_int64)]),
    "mmpgo_evaluate_f": (C.c_int, [_P, _dp, C.c_int64, _dp]),
    "mmpgo_evaluate_grad": (C.c_int, [_P, _dp, C.c_int64, _dp, C.c_int64]),
    "mmpgo_current_objective": (C.c_int, [_P, _dp, _dp]),
    "mmpgo_translation_solve": (C.c_int, [_P, _dp, _dp]),
    "mmpgo_star_objective": (C.c_int, [_P, _dp, _dp, _ip]),
    "mmpgo_graph_sizes": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "mmpgo_solver_info": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "mmpgo_preconditioner_info": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "mmpgo_stage_range": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "mmpgo_solver_stage_times": (C.c_int, [_P, _dp, _ip, _ip, C.c_int32, _ip]),
    "mmpgo_profile_pass": (C.c_int, [_P, C.c_int32, C.c_int32, C.POINTER(C.c_float)]),
    "mmpgo_get_counters": (C.c_int, [_P, C.POINTER(Counters)]),
    "mmpgo_reset_counters": (C.c_int, [_P]),
    "mmpgo_synchronize": (C.c_int, [_P]),
    "mmpgo_stream": (C.c_void_p, [_P]),
    "mmpgo_project_to_sodn": (C.c_int, [C.c_int32, C.c_int64, _dp, _dp, C.c_int32]),
    "mmpgo_mf_host_solve": (C.c_int, [C.c_int32, _ip, _ip, _dp, C.c_int32, C.c_int32, C.c_int32, _dp, _dp, _lp]),
}

_lib = None


def load():
    """dlopen libmmpgo.so and bind every declared entry point."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s is missing: build it with `make` (or __graft_entry__.build()); "
            "dpgo_b200 has no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise MmpgoError(rc, load().mmpgo_last_error().decode())


def dptr(a):
    assert a.dtype == np.float64
    return a.ctypes.data_as(_dp)


def iptr(a):
    assert a.dtype == np.int32
    return a.ctypes.data_as(_ip)

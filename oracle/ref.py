"""ctypes bindings of oracle/_ref (TEST INFRASTRUCTURE ONLY): the pieces of the reference that
compile from their own sources (oracle/Makefile) -- the AVX2 SO(2)/SO(3) projections and the
header-only TNT / STPCG.  Used to pin the restatement in oracle/sod.py and oracle/solver.py (and
the CUDA projection) against real reference code."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_dp = C.POINTER(C.c_double)


def build():
    """(Re)build oracle/_ref; needs /root/reference, i.e. only works in the authoring container."""
    subprocess.check_call(["make", "-C", HERE])


def available():
    return all(os.path.exists(os.path.join(REF_DIR, n)) for n in
               ("libref_sod.so", "libref_sod_nofma.so", "libref_tnt.so"))


def _sod(nofma):
    lib = C.CDLL(os.path.join(REF_DIR, "libref_sod_nofma.so" if nofma else "libref_sod.so"))
    for f in (lib.ref_project_to_SO3n, lib.ref_project_to_SO2n):
        f.argtypes = [_dp, _dp, C.c_long]
        f.restype = C.c_int
    return lib


def project(A, nofma=False):
    """A: (n, d, d) row-major blocks, n >= 4 -> the reference's projection of every block."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    n, d = A.shape[0], A.shape[1]
    U = np.empty_like(A)
    lib = _sod(nofma)
    fn = lib.ref_project_to_SO3n if d == 3 else lib.ref_project_to_SO2n
    rc = fn(A.ctypes.data_as(_dp), U.ctypes.data_as(_dp), n)
    if rc:
        raise ValueError("reference projection needs n >= 4")
    return U


class _Params(C.Structure):
    _fields_ = [(k, C.c_double) for k in
                ("gradient_tolerance", "preconditioned_gradient_tolerance", "relative_decrease_tolerance",
                 "stepsize_tolerance", "Delta0", "eta1", "eta2", "alpha1", "alpha2", "Delta_tolerance",
                 "kappa_fgr", "theta")] + \
               [(k, C.c_long) for k in ("max_iterations", "max_iterations_accepted", "max_TPCG_iterations")]


class _Out(C.Structure):
    _fields_ = [("f", C.c_double), ("gradfx_norm", C.c_double), ("status", C.c_long), ("iterations", C.c_long),
                ("n_inner", C.c_long), ("inner_iterations", C.c_long * 64), ("gain_ratios", C.c_double * 64)]


_F = C.CFUNCTYPE(C.c_double, _dp)
_QM = C.CFUNCTYPE(None, _dp, _dp)
_OP = C.CFUNCTYPE(None, _dp, _dp, _dp)
_MET = C.CFUNCTYPE(C.c_double, _dp, _dp, _dp)


def tnt(f, QM, metric, retract, x0, precon, params):
    """Same calling convention as oracle.solver.tnt, executed by the reference's TNT.h / STPCG.
    Points and tangent vectors are numpy matrices; their shapes are taken from x0 and QM(x0)."""
    lib = C.CDLL(os.path.join(REF_DIR, "libref_tnt.so"))
    xs = x0.shape
    g0, _ = QM(x0)
    ts = g0.shape
    nx, nt = int(np.prod(xs)), int(np.prod(ts))
    state = {}

    def X(p):
        return np.ctypeslib.as_array(p, shape=(nx,)).reshape(xs)

    def T(p):
        return np.ctypeslib.as_array(p, shape=(nt,)).reshape(ts)

    def c_f(px):
        return float(f(X(px).copy()))

    def c_qm(px, pg):
        x = X(px).copy()
        grad, hess = QM(x)
        state["hess"] = hess
        T(pg)[...] = grad

    def c_hess(px, pv, po):
        T(po)[...] = state["hess"](X(px).copy(), T(pv).copy())

    def c_met(px, pa, pb):
        return float(metric(X(px).copy(), T(pa).copy(), T(pb).copy()))

    def c_ret(px, pv, po):
        X(po)[...] = retract(X(px).copy(), T(pv).copy())

    def c_pre(px, pv, po):
        T(po)[...] = precon(X(px).copy(), T(pv).copy())

    P = _Params()
    for k, _t in _Params._fields_:
        setattr(P, k, getattr(params, k))
    out = _Out()
    xo = np.zeros(nx)
    cb = [_F(c_f), _QM(c_qm), _OP(c_hess), _MET(c_met), _OP(c_ret), _OP(c_pre) if precon is not None else C.cast(None, _OP)]
    lib.ref_tnt.restype = C.c_int
    x0c = np.ascontiguousarray(x0, dtype=np.float64).ravel()
    rc = lib.ref_tnt(C.c_long(nx), C.c_long(nt), x0c.ctypes.data_as(_dp), *cb, C.byref(P),
                     xo.ctypes.data_as(_dp), C.byref(out))
    if rc:
        raise RuntimeError("reference TNT threw")

    class R:
        pass
    r = R()
    r.x = xo.reshape(xs)
    r.f = out.f
    r.gradfx_norm = out.gradfx_norm
    r.status_code = out.status
    r.inner_iterations = [int(out.inner_iterations[i]) for i in range(min(out.n_inner, 64))]
    r.gain_ratios = [float(out.gain_ratios[i]) for i in range(min(out.n_inner, 64))]
    return r

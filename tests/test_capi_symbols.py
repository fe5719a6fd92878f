"""The C-ABI library loads and exports every symbol include/mmpgo.h declares
(no compute: this runs without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from dpgo_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mmpgo.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mmpgo_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    lib = L.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
        assert n in L.SIGNATURES, "no ctypes signature for %s" % n
    assert set(L.SIGNATURES) == set(names)


def test_struct_sizes_match_header_layout():
    # mmpgo_options: 4 int32 + 7 doubles + ... must be 8-byte aligned and stable
    assert ctypes.sizeof(L.Options) % 8 == 0
    assert ctypes.sizeof(L.NodeScalars) == 9 * 8 + 14 * 4
    assert ctypes.sizeof(L.Counters) == 16 * 8


def test_default_options_are_dist_pgo_values():
    o = L.Options()
    L.load().mmpgo_default_options(ctypes.byref(o))
    assert o.regularizer == 1e-11 and o.loss_reg == 0.25            # dist_pgo.cpp:107,120
    assert (o.eta[0], o.eta[1]) == (5e-4, 2.5e-2)                  # :111-112
    assert (o.max_soft_restart_hits[0], o.max_soft_restart_hits[1]) == (10, 25)
    assert o.max_iterations == 10 and o.max_iterations_accepted == 1
    assert o.STPCG_kappa == 0.05 and o.STPCG_theta == 0.9
    assert o.grad_norm_tol == 1e-3 and o.preconditioned_grad_norm_tol == 1e-4


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    o = L.Options()
    lib = L.load()
    lib.mmpgo_default_options(ctypes.byref(o))
    h = ctypes.c_void_p()
    rc = lib.mmpgo_create(ctypes.byref(o), ctypes.byref(h))
    assert rc == -2
    assert b"no CPU fallback" in lib.mmpgo_last_error()


def test_null_handle_is_an_error_not_a_crash():
    lib = L.load()
    assert lib.mmpgo_update(None) == -1
    assert lib.mmpgo_iterate(None) == -1
    assert lib.mmpgo_destroy(None) == 0

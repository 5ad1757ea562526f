"""The C restatement of the arithmetic contract (oracle/p1tet_oracle.c) must
reproduce the reference's element-local data and load vector bitwise."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from cases import load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "oracle", "p1tet_oracle.c")
SO = os.path.join(ROOT, "oracle", "_build", "libp1tet_oracle.so")


def _lib():
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", SO,
                               SRC, "-lm"])
    return C.CDLL(SO)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("name", ["tet_p1_tensor6", "tet_p1_ball2", "tet_p1_refined3",
                                  "tet_p1_morphed5", "tet_p1_tensor_nonuniform"])
def test_c_oracle_bitwise(name):
    lib = _lib()
    g = load(name)
    p = np.ascontiguousarray(g["p"]); t = np.ascontiguousarray(g["t"])
    phi = np.ascontiguousarray(g["phi"]); dphi = np.ascontiguousarray(g["dphi"])
    W = np.ascontiguousarray(g["W"])
    nel, nqp, N = t.shape[1], W.shape[0], int(g["N"])
    local = np.empty(16 * nel)
    rc = lib.p1tet_laplace_local(_ptr(p), C.c_int64(p.shape[1]), _ptr(t), C.c_int64(nel),
                                 _ptr(dphi), _ptr(W), C.c_int(nqp), _ptr(local))
    assert rc == 0
    assert np.array_equal(local, g["laplace_local"])
    vec = np.empty(N)
    rc = lib.p1tet_unit_load(_ptr(p), C.c_int64(p.shape[1]), _ptr(t), C.c_int64(nel), _ptr(phi),
                             _ptr(W), C.c_int(nqp), C.c_int64(N), _ptr(vec))
    assert rc == 0
    assert np.array_equal(vec, g["unit_load_vec"])

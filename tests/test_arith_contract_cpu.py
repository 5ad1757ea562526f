"""The arithmetic contract of the kernels, checked on the CPU against numpy:
the shipped header csrc/skb_common.cuh is compiled by g++ (tests/host_arith.py) and its results
must be bit-identical to the reference's numpy expressions - the same bar the GPU parity tests
apply to whole kernels."""
import ctypes as C

import numpy as np
import pytest

import host_arith


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


def test_exact_div_is_ieee_division():
    """exact_div (reciprocal + two Markstein corrections) == RN(a / b), on random operands over
    the whole safe exponent range and on operands built to sit next to rounding boundaries."""
    lib = host_arith.lib()
    rng = np.random.default_rng(11)
    n = 400000
    a = rng.standard_normal(n) * np.exp2(rng.integers(-250, 120, n).astype(np.float64))
    b = rng.standard_normal(n) * np.exp2(rng.integers(-300, 150, n).astype(np.float64))
    # adversarial: quotients of small integers / near-ties (q * b rounded back), mantissa edges
    k = 100000
    bi = rng.integers(1, 2 ** 26, k).astype(np.float64)
    qi = rng.integers(1, 2 ** 27, k).astype(np.float64) + 0.5
    a = np.r_[a, qi * bi, np.nextafter(qi * bi, np.inf), 1.0 + rng.random(k) * 2.0 ** -30,
              np.full(k, 1.0)]
    b = np.r_[b, bi, bi, 3.0 * np.ones(k), 1.0 + rng.random(k) * 2.0 ** -26]
    b[b == 0] = 1.0
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    assert all(lib.host_exp_in_safe_range(float(v)) for v in (a[:50].tolist() + b[:50].tolist()))
    q = np.empty_like(a)
    lib.host_exact_div(_ptr(a), _ptr(b), _ptr(q), C.c_int64(a.size))
    assert np.array_equal(q, a / b)


def test_exponent_guard():
    lib = host_arith.lib()
    for v, ok in ((0.0, 1), (-0.0, 1), (1.0, 1), (2.0 ** -400, 1), (2.0 ** 400, 1),
                  (2.0 ** -401, 0), (2.0 ** 402, 0), (5e-324, 0), (np.inf, 0), (np.nan, 0)):
        assert lib.host_exp_in_safe_range(float(v)) == ok, v


@pytest.mark.parametrize("scale", [1.0, 2.0 ** -70, 2.0 ** 45, 2.0 ** -140, 2.0 ** 140])
def test_affine_geometry_matches_the_reference_expressions(scale):
    """A, b, det, inv of affine_load / affine_invert<3> == mapping_affine.py:55-131 evaluated by
    numpy (oracle.affine_geometry), bit for bit - also for tiny / huge coordinates, where
    divide9 takes its plain-division branch."""
    from types import SimpleNamespace
    from oracle import skfem_oracle as O
    lib = host_arith.lib()
    rng = np.random.default_rng(5)
    g = np.sort(np.r_[0., rng.uniform(0.05, 0.95, 6), 1.])
    m = O.mesh_tet_tensor(g, g ** 2, np.sqrt(g))
    p = m.p.copy()
    p[0] = p[0] + 0.03 * np.sin(7 * p[1])
    p = np.ascontiguousarray(p * scale)
    t = np.ascontiguousarray(m.t.astype(np.int32))
    nel = t.shape[1]
    A, b, inv, det = np.empty((nel, 3, 3)), np.empty((nel, 3)), np.empty((nel, 3, 3)), np.empty(nel)
    lib.host_affine3(_ptr(p), C.c_int64(p.shape[1]), _ptr(t), C.c_int64(nel),
                     _ptr(A), _ptr(b), _ptr(inv), _ptr(det))
    ref = O.affine_geometry(SimpleNamespace(p=p, t=t))
    assert np.array_equal(A, ref.A.transpose(2, 0, 1)) and np.array_equal(b, ref.b.T)
    assert np.array_equal(det, ref.detA)
    assert np.array_equal(inv, ref.invA.transpose(2, 0, 1))


def test_affine_geometry_2d():
    from types import SimpleNamespace
    from oracle import skfem_oracle as O
    lib = host_arith.lib()
    m = O.refine_tri(O.mesh_tri_default(), 3)
    p = np.ascontiguousarray(m.p + 0.02 * np.sin(5 * m.p[::-1]))
    t = np.ascontiguousarray(m.t.astype(np.int32))
    nel = t.shape[1]
    inv, det = np.empty((nel, 2, 2)), np.empty(nel)
    lib.host_affine2(_ptr(p), C.c_int64(p.shape[1]), _ptr(t), C.c_int64(nel), _ptr(inv), _ptr(det))
    ref = O.affine_geometry(SimpleNamespace(p=p, t=t))
    assert np.array_equal(det, ref.detA) and np.array_equal(inv, ref.invA.transpose(2, 0, 1))


def test_streaming_pairwise_sum_is_numpys():
    """pw_sum(n, f) == np.sum over a contiguous axis (numpy's pairwise order) for every length
    the quadrature rules produce and around the 8 / 128 block boundaries; seq_sum == the plain
    left-to-right sum numpy uses for Fortran-ordered products."""
    lib = host_arith.lib()
    rng = np.random.default_rng(2)
    for n in list(range(1, 41)) + [63, 64, 65, 127, 128, 129, 136, 255, 256, 257, 343, 512, 1000]:
        v = np.ascontiguousarray(rng.standard_normal(n) * np.exp2(rng.integers(-20, 20, n)))
        got = lib.host_pw_sum(_ptr(v), C.c_int(n))
        assert got == np.sum(v), n
        seq = 0.0
        for x in v:
            seq = seq + x
        assert lib.host_seq_sum(_ptr(v), C.c_int(n)) == seq, n

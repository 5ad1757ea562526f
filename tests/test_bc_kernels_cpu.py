"""Boundary-condition / SpMV kernels (csrc/skb_bc.cu) run on the CPU from the shipped source
(tests/host_bc.py) against scipy, bit for bit: ``condense`` as scipy's ``A[I][:, I]`` and
``b[I] - A[I][:, D] @ x[D]`` (skfem/utils.py:462-603), ``enforce`` as the oracle's restatement of
skfem/utils.py:327-400, SpMV as scipy's ``csr_matvec``."""
import ctypes as C

import numpy as np
import pytest

import host_bc
from cases import load, mesh_of


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


def _system(name):
    from oracle import skfem_oracle as O
    g = load(name)
    refdom, ename = ("tri", "tri_p1") if name.startswith("c1") else ("tet", "tet_p2")
    b = O.cell_basis(mesh_of(g, refdom), O.element(ename))
    A = O.assemble_bilinear(O.laplace, b)
    A.sort_indices()
    f = O.assemble_linear(O.unit_load, b)
    N = A.shape[0]
    rng = np.random.default_rng(4)
    D = np.sort(rng.choice(N, size=N // 5, replace=False)).astype(np.int32)
    x = rng.standard_normal(N)
    return A, f, D, x


@pytest.mark.parametrize("name", ["c1_tri_p1_refined4", "tet_p2_morphed3"])
def test_condense_and_spmv_kernels_match_scipy(name):
    lib = host_bc.lib()
    A, f, D, x = _system(name)
    N = A.shape[0]
    indptr, indices, data = (np.ascontiguousarray(v) for v in (A.indptr, A.indices, A.data))
    # SpMV
    y = np.full(N, np.nan)
    lib.host_spmv(_p(indptr), _p(indices), _p(data), _p(x), _p(y), C.c_int64(N))
    assert np.array_equal(y, A @ x)
    # condense: rows / columns I, right-hand side b[I] - A[I][:, D] x[D]
    I = np.setdiff1d(np.arange(N), D).astype(np.int32)
    colmap = np.full(N, -1, dtype=np.int32)
    colmap[I] = np.arange(len(I), dtype=np.int32)
    counts = np.zeros(len(I), dtype=np.int32)
    lib.host_condense_count(_p(indptr), _p(indices), _p(I), C.c_int64(len(I)), _p(colmap),
                            _p(counts))
    new_indptr = np.r_[0, np.cumsum(counts)].astype(np.int32)
    nnz = int(new_indptr[-1])
    new_indices, new_data = np.full(nnz, -1, dtype=np.int32), np.full(nnz, np.nan)
    bout = np.full(len(I), np.nan)
    lib.host_condense_fill(_p(indptr), _p(indices), _p(data), _p(I), C.c_int64(len(I)),
                           _p(colmap), _p(new_indptr), _p(new_indices), _p(new_data), _p(x),
                           _p(f), _p(bout))
    ref = A[I][:, I].tocsr()
    ref.sort_indices()
    assert np.array_equal(new_indptr, ref.indptr) and np.array_equal(new_indices, ref.indices)
    assert np.array_equal(new_data, ref.data)
    assert np.array_equal(bout, f[I] - A[I][:, D] @ x[D])


@pytest.mark.parametrize("name", ["c1_tri_p1_refined4", "tet_p2_morphed3"])
def test_enforce_kernel_matches_the_reference_result(name):
    from oracle import skfem_oracle as O
    lib = host_bc.lib()
    A, f, D, x = _system(name)
    indptr, indices = np.ascontiguousarray(A.indptr), np.ascontiguousarray(A.indices)
    data = np.ascontiguousarray(A.data.copy())
    missing = np.zeros(1, dtype=np.int32)
    lib.host_enforce(_p(indptr), _p(indices), _p(data), _p(D), C.c_int64(len(D)),
                     C.c_double(1.0), _p(missing))
    assert missing[0] == 0                      # every constrained row has a stored diagonal
    ours = type(A)((data, indices, indptr), shape=A.shape).toarray()
    ref = O.enforce(A, D=D)
    ref = ref[0] if isinstance(ref, tuple) else ref
    dense = ref.toarray()
    # the kernel handles the rows D; the symmetric column elimination is a second pass on the
    # transpose in skfem_b200/utils.py - compare the rows
    assert np.array_equal(ours[D], dense[D])

"""Test infrastructure: the sparsity-plan kernels (csrc/skb_plan.cu) on the host.

``make_keys`` / ``head_flags`` / ``counts`` / ``finalize`` / ``csr_reduce`` / ``vec_reduce`` are
scalar grid-stride kernels; the launchers chain them with a CUB radix sort (stable, by key) and
a CUB inclusive scan.  Here the kernels are compiled by g++ from the shipped source (same shim
as tests/host_local.py) and chained by ``symbolic`` / ``reduce`` below with numpy's stable
argsort and cumsum standing in for the two CUB calls - the replacement for
``COOData._assemble_scipy_csr`` (coo_data.py:27-36) end to end on the CPU.  Not product code.
"""
import ctypes as C
import os
import subprocess

import numpy as np

import host_local

ROOT = host_local.ROOT
SRC = os.path.join(host_local.CSRC, "skb_plan.cu")
LIB = os.path.join(host_local.OUT_DIR, "libplan_host.so")

DRIVERS = r"""
using namespace skb;
extern "C" void host_make_keys(const int32_t *dofs_v, const int32_t *dofs_u, int nbv, int64_t nel,
                               int64_t ncoo, uint64_t ncols, uint64_t sentinel, const double *local,
                               int drop_zeros, uint64_t *keys, uint32_t *vals) {
  make_keys_kernel(dofs_v, dofs_u, nbv, nel, ncoo, ncols, sentinel, local, drop_zeros, keys, vals);
}
extern "C" void host_head_flags(const uint64_t *keys, int64_t ncoo, uint64_t sentinel,
                                uint32_t *flag) {
  head_flags_kernel(keys, ncoo, sentinel, flag);
}
extern "C" void host_counts(const uint64_t *keys, const uint32_t *slot, int64_t ncoo,
                            uint64_t sentinel, unsigned long long *counts) {
  skb_tid.x = 0; counts_kernel(keys, slot, ncoo, sentinel, counts);   // thread 0: nnz
  skb_tid.x = 1; counts_kernel(keys, slot, ncoo, sentinel, counts);   // thread 1: nkeep
  skb_tid.x = 0;
}
extern "C" void host_finalize(const uint64_t *keys, const uint32_t *vals, const uint32_t *slot,
                              int64_t nkeep, int64_t nnz, int64_t nrows, uint64_t ncols,
                              int32_t *indptr, int32_t *indices, uint32_t *segptr, uint32_t *perm) {
  finalize_kernel(keys, vals, slot, nkeep, nnz, nrows, ncols, indptr, indices, segptr, perm);
}
extern "C" void host_csr_reduce(const double *local, const uint32_t *perm, const uint32_t *segptr,
                                int64_t nnz, double *data) {
  csr_reduce_kernel(local, perm, segptr, nnz, data);
}
extern "C" void host_csr_reduce_em(const double *local_em, uint32_t nel, uint32_t nbu, uint32_t nbv,
                                   const uint32_t *perm, const uint32_t *segptr, int64_t nnz,
                                   double *data) {
  csr_reduce_em_kernel(local_em, nel, nbu, nbv, perm, segptr, nnz, data);
}
extern "C" void host_vec_reduce(const double *local, const uint32_t *perm, const uint32_t *segptr,
                                const int32_t *indptr, int64_t nrows, double *vec) {
  vec_reduce_kernel(local, perm, segptr, indptr, nrows, vec);
}
"""


def _host_source():
    body = open(SRC).read().split('#include "skb_common.cuh"', 1)[1]
    body = body[:body.index("static inline int nblocks")]
    prelude = (host_local.PRELUDE % {"hdr": host_local.HDR}).replace(
        "#define threadIdx skb_zero", "static skb_idx3 skb_tid = {0, 0, 0};\n#define threadIdx skb_tid")
    open_ns = body.count("namespace skb {") - body.count("}  // namespace skb")
    return prelude + body + "}\n" * open_ns + DRIVERS


def build():
    os.makedirs(host_local.OUT_DIR, exist_ok=True)
    if os.path.exists(LIB) and os.path.getmtime(LIB) >= max(
            os.path.getmtime(SRC), os.path.getmtime(host_local.HDR), os.path.getmtime(__file__),
            os.path.getmtime(host_local.__file__)):
        return LIB
    cpp = os.path.join(host_local.OUT_DIR, "plan_host.cpp")
    with open(cpp, "w") as f:
        f.write(_host_source())
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC",
                    "-Wno-unknown-pragmas", "-I", cuda_inc, "-o", LIB, cpp],
                   check=True, capture_output=True)
    return LIB


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


def symbolic(local, dofs_v, dofs_u, nel, nrows, ncols, drop_zeros=True):
    """skb_plan_symbolic + skb_plan_finalize: -> dict(indptr, indices, segptr, perm, nnz)."""
    lib = C.CDLL(build())
    dofs_v = np.ascontiguousarray(dofs_v, dtype=np.int32)
    nbv = dofs_v.shape[0]
    nbu = 1 if dofs_u is None else dofs_u.shape[0]
    if dofs_u is not None:
        dofs_u = np.ascontiguousarray(dofs_u, dtype=np.int32)
    ncoo = nbv * nbu * nel
    sentinel = nrows * ncols
    keys, vals = np.empty(ncoo, dtype=np.uint64), np.empty(ncoo, dtype=np.uint32)
    lib.host_make_keys(_p(dofs_v), _p(dofs_u), C.c_int(nbv), C.c_int64(nel), C.c_int64(ncoo),
                       C.c_uint64(ncols), C.c_uint64(sentinel), _p(local), C.c_int(int(drop_zeros)),
                       _p(keys), _p(vals))
    order = np.argsort(keys, kind="stable")          # cub::DeviceRadixSort::SortPairs (stable)
    keys, vals = np.ascontiguousarray(keys[order]), np.ascontiguousarray(vals[order])
    flag = np.empty(ncoo, dtype=np.uint32)
    lib.host_head_flags(_p(keys), C.c_int64(ncoo), C.c_uint64(sentinel), _p(flag))
    slot = np.cumsum(flag, dtype=np.uint32)          # cub::DeviceScan::InclusiveSum
    counts = np.zeros(2, dtype=np.uint64)
    lib.host_counts(_p(keys), _p(slot), C.c_int64(ncoo), C.c_uint64(sentinel), _p(counts))
    nnz, nkeep = int(counts[0]), int(counts[1])
    indptr = np.full(nrows + 1, -1, dtype=np.int32)
    indices = np.full(nnz, -1, dtype=np.int32)
    segptr = np.zeros(nnz + 1, dtype=np.uint32)
    perm = np.zeros(max(nkeep, 1), dtype=np.uint32)
    lib.host_finalize(_p(keys), _p(vals), _p(slot), C.c_int64(nkeep), C.c_int64(nnz),
                      C.c_int64(nrows), C.c_uint64(ncols), _p(indptr), _p(indices), _p(segptr),
                      _p(perm))
    return dict(indptr=indptr, indices=indices, segptr=segptr, perm=perm, nnz=nnz, nkeep=nkeep)


def csr_reduce(local, plan):
    lib = C.CDLL(build())
    data = np.full(plan["nnz"], np.nan)
    lib.host_csr_reduce(_p(local), _p(plan["perm"]), _p(plan["segptr"]), C.c_int64(plan["nnz"]),
                        _p(data))
    return data


def csr_reduce_em(local_em, nel, nbu, nbv, plan):
    """skb_csr_reduce_em: the plan's COO indices against element-major local data."""
    lib = C.CDLL(build())
    data = np.full(plan["nnz"], np.nan)
    lib.host_csr_reduce_em(_p(local_em), C.c_uint32(nel), C.c_uint32(nbu), C.c_uint32(nbv),
                           _p(plan["perm"]), _p(plan["segptr"]), C.c_int64(plan["nnz"]), _p(data))
    return data


def vec_reduce(local, plan, nrows):
    lib = C.CDLL(build())
    vec = np.full(nrows, np.nan)
    lib.host_vec_reduce(_p(local), _p(plan["perm"]), _p(plan["segptr"]), _p(plan["indptr"]),
                        C.c_int64(nrows), _p(vec))
    return vec

"""Test infrastructure: the boundary-condition / SpMV kernels (csrc/skb_bc.cu) on the host.

Scalar grid-stride kernels (one thread per row); compiled by g++ from the shipped source with
the shim of tests/host_local.py plus a plain ``atomicExch``.  Not product code.
"""
import ctypes as C
import os
import subprocess

import host_local

SRC = os.path.join(host_local.CSRC, "skb_bc.cu")
LIB = os.path.join(host_local.OUT_DIR, "libbc_host.so")

DRIVERS = r"""
using namespace skb;
extern "C" void host_enforce(const int32_t *indptr, const int32_t *indices, double *data,
                             const int32_t *D, int64_t nD, double diag, int *missing) {
  enforce_kernel(indptr, indices, data, D, nD, diag, missing);
}
extern "C" void host_condense_count(const int32_t *indptr, const int32_t *indices, const int32_t *I,
                                    int64_t nI, const int32_t *colmap, int32_t *counts) {
  condense_count_kernel(indptr, indices, I, nI, colmap, counts);
}
extern "C" void host_condense_fill(const int32_t *indptr, const int32_t *indices, const double *data,
                                   const int32_t *I, int64_t nI, const int32_t *colmap,
                                   const int32_t *new_indptr, int32_t *new_indices,
                                   double *new_data, const double *x, const double *b,
                                   double *bout) {
  condense_fill_kernel(indptr, indices, data, I, nI, colmap, new_indptr, new_indices, new_data, x,
                       b, bout);
}
extern "C" void host_spmv(const int32_t *indptr, const int32_t *indices, const double *data,
                          const double *x, double *y, int64_t nrows) {
  spmv_kernel(indptr, indices, data, x, y, nrows);
}
"""


def _host_source():
    body = open(SRC).read().split('#include "skb_common.cuh"', 1)[1]
    body = body[:body.index("static inline int bc_blocks")]
    prelude = (host_local.PRELUDE % {"hdr": host_local.HDR}).replace(
        "static inline void __syncthreads() {}",
        "static inline void __syncthreads() {}\n"
        "static inline int atomicExch(int *p, int v) { int o = *p; *p = v; return o; }")
    open_ns = body.count("namespace skb {") - body.count("}  // namespace skb")
    return prelude + body + "}\n" * open_ns + DRIVERS


def build():
    os.makedirs(host_local.OUT_DIR, exist_ok=True)
    if os.path.exists(LIB) and os.path.getmtime(LIB) >= max(
            os.path.getmtime(SRC), os.path.getmtime(host_local.HDR), os.path.getmtime(__file__),
            os.path.getmtime(host_local.__file__)):
        return LIB
    cpp = os.path.join(host_local.OUT_DIR, "bc_host.cpp")
    with open(cpp, "w") as f:
        f.write(_host_source())
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC",
                    "-Wno-unknown-pragmas", "-I", cuda_inc, "-o", LIB, cpp],
                   check=True, capture_output=True)
    return LIB


def lib():
    return C.CDLL(build())

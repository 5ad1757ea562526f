"""Test infrastructure: the block-cooperative hexahedral kernels on the host.

``local_hex_kernel`` (csrc/skb_local.cu: ElementHex1 / the scalar Hex2 path, reference operation
order) and ``local_hex_sf_kernel`` (csrc/skb_hex_sf.cu: ElementHex2 by sum factorisation) use a
whole thread block per element batch with ``__syncthreads()`` between their stages.  Here the
shipped sources are compiled with g++ and every CUDA thread of a block becomes a host thread:
``threadIdx`` / ``blockIdx`` are thread-local, ``__syncthreads()`` is a ``std::barrier`` over the
block, ``__shared__`` variables are function statics and the dynamic shared memory a global
buffer (blocks run one after the other), kernel parameters are passed by value.  CUDA
intrinsics are spelled in standard C++ as in tests/host_local.py, ``-ffp-contract=off`` stands
for nvcc's ``-fmad=false``.  tests/test_hex_kernels_cpu.py compares the results with the
reference's element-local data (tests/golden).  Not product code.
"""
import ctypes as C
import os
import subprocess

import host_local

ROOT = host_local.ROOT
CSRC = host_local.CSRC
SRC_LOCAL = os.path.join(CSRC, "skb_local.cu")
SRC_SF = os.path.join(CSRC, "skb_hex_sf.cu")
HDR = host_local.HDR
OUT_DIR = host_local.OUT_DIR
LIB = os.path.join(OUT_DIR, "libblock_host.so")

PRELUDE = r"""
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>
static inline double __drcp_rn(double x) { return 1.0 / x; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline int __double2hiint(double v) { int64_t b; std::memcpy(&b, &v, 8); return (int)(b >> 32); }
static inline int __double2loint(double v) { int64_t b; std::memcpy(&b, &v, 8); return (int)(b & 0xffffffff); }
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline int atomicExch(int *p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
static std::barrier<> *skb_block_barrier = nullptr;
static inline void __syncthreads() { skb_block_barrier->arrive_and_wait(); }
#include "%(hdr)s"
struct skb_idx3 { unsigned x, y, z; };
static thread_local skb_idx3 skb_tid = {0, 0, 0}, skb_bid = {0, 0, 0};
static skb_idx3 skb_bdim = {1, 1, 1}, skb_gdim = {1, 1, 1};
#define threadIdx skb_tid
#define blockIdx skb_bid
#define blockDim skb_bdim
#define gridDim skb_gdim
#undef __global__
#define __global__
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __shared__
#define __shared__ static
#undef __grid_constant__
#define __grid_constant__
static double skb_dyn_smem[1 << 18];

// one host thread per CUDA thread, blocks one after the other
template <class F> static void skb_run_grid(unsigned grid, unsigned block, F kernel) {
  skb_bdim = {block, 1, 1};
  skb_gdim = {grid, 1, 1};
  for (unsigned b = 0; b < grid; ++b) {
    std::barrier<> bar((std::ptrdiff_t)block);
    skb_block_barrier = &bar;
    std::vector<std::thread> threads;
    for (unsigned t = 0; t < block; ++t)
      threads.emplace_back([=]() {
        skb_tid = {t, 0, 0};
        skb_bid = {b, 0, 0};
        kernel();
      });
    for (auto &th : threads) th.join();
  }
}
"""

DRIVERS = r"""
extern "C" int host_local_hex(const skb_space_t *s, int form, int bilinear, int grid, double *out,
                              int *err) {
  const skb_space_t sp = *s;
  if (bilinear) skb_run_grid(grid, 256, [=]() { skb::local_hex_kernel<true>(sp, form, out, err); });
  else skb_run_grid(grid, 256, [=]() { skb::local_hex_kernel<false>(sp, form, out, err); });
  return 0;
}

extern "C" int host_hex_sumfact(const skb_space_t *s, int form, const int32_t *qstride,
                                const double *pp, const double *g, const uint8_t *bnode,
                                const uint8_t *vtx, int em, int shape, int grid, double *out,
                                int *err) {
  using namespace skb;
  const skb_space_t sp = *s;
  HexSfTab tb;                              // as skb_local_hex_sumfact fills it
  memcpy(tb.pp, pp, sizeof(tb.pp));
  memcpy(tb.g, g, sizeof(tb.g));
  memset(tb.bnode, 0, sizeof(tb.bnode));
  memcpy(tb.bnode, bnode, 27);
  memcpy(tb.vtx, vtx, 8);
  for (int k = 0; k < 3; ++k) tb.qs[k] = qstride[k];
  if (form == SKB_FORM_MASS)
    skb_run_grid(grid, 256, [=]() { local_hex_sf_kernel<SKB_FORM_MASS, 4, 256, 1>(sp, tb, em, out, err); });
  else if (shape == 1)
    skb_run_grid(grid, 256, [=]() { local_hex_sf_kernel<SKB_FORM_LAPLACE, 4, 256, 1>(sp, tb, em, out, err); });
  else
    skb_run_grid(grid, 128, [=]() { local_hex_sf_kernel<SKB_FORM_LAPLACE, 2, 128, 2>(sp, tb, em, out, err); });
  return 0;
}
"""


def _host_source():
    dyn = "double *smem = skb_dyn_smem;"
    local = open(SRC_LOCAL).read().split('#include "skb_common.cuh"', 1)[1]
    local = local[:local.index("static int grid_for(")]
    local = local.replace("extern __shared__ double smem[];  // inv[9][nqp], dx[nqp]", dyn)
    local = local.replace("extern __shared__ double smem[];", dyn)
    local += "}\n" * (local.count("namespace skb {") - local.count("}  // namespace skb"))
    sf = open(SRC_SF).read().split('#include "skb_common.cuh"', 1)[1]
    sf = sf[:sf.index("// launched by skb_local_hex_sumfact")]
    sf = sf.replace("extern __shared__ double sm[];", "double *sm = skb_dyn_smem;")
    sf += "}\n" * (sf.count("namespace skb {") - sf.count("}  // namespace skb"))
    assert "extern __shared__" not in local + sf
    return PRELUDE % {"hdr": HDR} + local + sf + DRIVERS


def build():
    os.makedirs(OUT_DIR, exist_ok=True)
    deps = [SRC_LOCAL, SRC_SF, HDR, __file__]
    if os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(d) for d in deps):
        return LIB
    cpp = os.path.join(OUT_DIR, "block_host.cpp")
    with open(cpp, "w") as f:
        f.write(_host_source())
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.run(["g++", "-O2", "-std=c++20", "-pthread", "-ffp-contract=off", "-shared",
                    "-fPIC", "-Wno-unknown-pragmas", "-I", cuda_inc, "-o", LIB, cpp],
                   check=True, capture_output=True)
    return LIB


def lib():
    return C.CDLL(build())

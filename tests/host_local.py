"""Test infrastructure: the element-local kernels of affine meshes on the host.

``local_affine_kernel`` / ``local_affine_cached_kernel`` (csrc/skb_local.cu) are one-thread-
per-element kernels: apart from staging the reference tables in shared memory they contain no
inter-thread communication.  Run as a single thread of a single block (``blockDim = gridDim =
1``: the staging loop copies every entry, the grid-stride loop visits every element) the same
source compiles with g++ - CUDA intrinsics spelled in standard C++ as in tests/host_arith.py,
``extern __shared__`` replaced by a static buffer, ``-ffp-contract=off`` for nvcc's
``-fmad=false``.  tests/test_local_kernels_cpu.py compares the output bit for bit with the
reference's element-local data (tests/golden).  The block-cooperative hexahedral kernels run on
the host through tests/host_block.py (one host thread per CUDA thread); everything with TMA /
tensor cores stays GPU-only.  Not product code.
"""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "scikit-fem_b200", "csrc")
SRC = os.path.join(CSRC, "skb_local.cu")
HDR = os.path.join(CSRC, "skb_common.cuh")
OUT_DIR = os.path.join(ROOT, "oracle", "_build")
LIB = os.path.join(OUT_DIR, "liblocal_host.so")

PRELUDE = r"""
#include <cmath>
#include <cstdint>
#include <cstring>
static inline double __drcp_rn(double x) { return 1.0 / x; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline int __double2hiint(double v) { int64_t b; std::memcpy(&b, &v, 8); return (int)(b >> 32); }
static inline int __double2loint(double v) { int64_t b; std::memcpy(&b, &v, 8); return (int)(b & 0xffffffff); }
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline void __syncthreads() {}
#include "%(hdr)s"
struct skb_idx3 { unsigned x, y, z; };
static skb_idx3 skb_one = {1, 1, 1}, skb_zero = {0, 0, 0};
#define threadIdx skb_zero
#define blockIdx skb_zero
#define blockDim skb_one
#define gridDim skb_one
#undef __global__
#define __global__
#undef __launch_bounds__
#define __launch_bounds__(...)
static double skb_host_smem[1 << 18];
"""

DRIVERS = r"""
extern "C" int host_local_affine(const skb_space_t *s, int form, double lambda, double two_mu,
                                 double *out, int bilinear, int cached) {
  const bool vec = s->ncomp > 1;
  if (cached == 3) {     // cached kernel of vector elements with a compile-time rule size
    if (!vec || !bilinear) return -1;
#define FIX(D, Q) if (s->dim == D && s->nqp == Q) { skb::local_affine_cached_kernel<D, true, Q>(*s, form, lambda, two_mu, out); return 0; }
    FIX(3, 4) FIX(3, 11) FIX(2, 3) FIX(2, 6)
#undef FIX
    return -1;
  }
  if (cached == 5) {     // the same, element-major output (nel, Nbv, Nbu)
    if (!vec || !bilinear) return -1;
#define FIX(D, Q) if (s->dim == D && s->nqp == Q) { skb::local_affine_cached_kernel<D, true, Q, true>(*s, form, lambda, two_mu, out); return 0; }
    FIX(3, 4) FIX(3, 11) FIX(2, 3) FIX(2, 6)
#undef FIX
    return -1;
  }
  if (cached == 4) {     // symmetric kernel of scalar elements, element-major output
    if (vec || !bilinear) return -1;
#define SYM(D, Q) if (s->dim == D && s->nqp == Q) { skb::local_affine_sym_kernel<D, Q, true>(*s, form, out); return 0; }
    SYM(3, 4) SYM(3, 11) SYM(2, 3) SYM(2, 6)
#undef SYM
    return -1;
  }
  if (cached == 2) {     // register-cached symmetric kernel of scalar elements
    if (vec || !bilinear) return -1;
#define SYM(D, Q) if (s->dim == D && s->nqp == Q) { skb::local_affine_sym_kernel<D, Q>(*s, form, out); return 0; }
    SYM(3, 4) SYM(3, 11) SYM(2, 3) SYM(2, 6)
#undef SYM
    return -1;
  }
#define CALL(D, V)                                                                     \
  do {                                                                                 \
    if (cached) skb::local_affine_cached_kernel<D, V>(*s, form, lambda, two_mu, out);  \
    else if (bilinear) skb::local_affine_kernel<D, V, true>(*s, form, lambda, two_mu, out);  \
    else skb::local_affine_kernel<D, V, false>(*s, form, lambda, two_mu, out);         \
  } while (0)
  if (s->dim == 2 && !vec) CALL(2, false);
  else if (s->dim == 2) CALL(2, true);
  else if (!vec) CALL(3, false);
  else CALL(3, true);
#undef CALL
  return 0;
}
"""


def _host_source():
    src = open(SRC).read()
    body = src.split('#include "skb_common.cuh"', 1)[1]
    # keep everything up to the block-cooperative hexahedral kernel
    cut = body.index("__device__ __forceinline__ void hex_jacobian(")
    body = body[:cut].replace("extern __shared__ double smem[];", "double *smem = skb_host_smem;")
    open_ns = body.count("namespace skb {") - body.count("}  // namespace skb")
    return PRELUDE % {"hdr": HDR} + body + "}\n" * open_ns + DRIVERS


def build():
    os.makedirs(OUT_DIR, exist_ok=True)
    if os.path.exists(LIB) and os.path.getmtime(LIB) >= max(
            os.path.getmtime(SRC), os.path.getmtime(HDR), os.path.getmtime(__file__)):
        return LIB
    cpp = os.path.join(OUT_DIR, "local_host.cpp")
    with open(cpp, "w") as f:
        f.write(_host_source())
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC",
                    "-Wno-unknown-pragmas", "-I", cuda_inc, "-o", LIB, cpp],
                   check=True, capture_output=True)
    return LIB


def lib():
    return C.CDLL(build())

"""Test infrastructure: the arithmetic-contract header of the kernels on the host.

``csrc/skb_common.cuh`` holds the device functions every kernel builds its numbers from: the
affine geometry (``affine_load`` / ``affine_invert``: determinant, cofactors, the 9 quotients),
``exact_div`` (one reciprocal + Markstein corrections instead of 9 IEEE divisions) and numpy's
pairwise summation as a streaming sum (``pw_sum``).  They are plain C++ apart from five CUDA
intrinsics, so g++ compiles the *shipped header* once those are spelled in standard C++
(``__drcp_rn`` = ``1.0 / x``, ``__fma_rn`` = ``std::fma``, ``__ldg`` = a load ...), with
``-ffp-contract=off`` standing in for nvcc's ``-fmad=false``.  tests/test_arith_contract_cpu.py
compares the results bit for bit with numpy / the oracle.  Not product code.
"""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "scikit-fem_b200", "csrc", "skb_common.cuh")
OUT_DIR = os.path.join(ROOT, "oracle", "_build")
LIB = os.path.join(OUT_DIR, "libarith_host.so")

SOURCE = r"""
#include <cmath>
#include <cstdint>
#include <cstring>
// the CUDA intrinsics the header uses, in standard C++ (IEEE binary64, round to nearest)
static inline double __drcp_rn(double x) { return 1.0 / x; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline int __double2hiint(double v) { int64_t b; std::memcpy(&b, &v, 8); return (int)(b >> 32); }
static inline int __double2loint(double v) { int64_t b; std::memcpy(&b, &v, 8); return (int)(b & 0xffffffff); }
template <class T> static inline T __ldg(const T *p) { return *p; }
#include "%(hdr)s"

extern "C" void host_exact_div(const double *a, const double *b, double *q, int64_t n) {
  for (int64_t i = 0; i < n; ++i) q[i] = skb::exact_div(a[i], b[i], __drcp_rn(b[i]));
}
extern "C" int host_exp_in_safe_range(double v) { return skb::exp_in_safe_range(v) ? 1 : 0; }
// per element: A (9), b (3), inv (9) row-major [i][j], det
extern "C" void host_affine3(const double *p, int64_t npts, const int32_t *t, int64_t nel,
                             double *A, double *b, double *inv, double *det) {
  for (int64_t e = 0; e < nel; ++e) {
    skb::Affine<3> g;
    skb::affine_load<3>(g, p, npts, t, nel, e);
    skb::affine_invert(g);
    for (int i = 0; i < 3; ++i) {
      b[3 * e + i] = g.b[i];
      for (int j = 0; j < 3; ++j) {
        A[9 * e + 3 * i + j] = g.A[i][j];
        inv[9 * e + 3 * i + j] = g.inv[i][j];
      }
    }
    det[e] = g.det;
  }
}
extern "C" void host_affine2(const double *p, int64_t npts, const int32_t *t, int64_t nel,
                             double *inv, double *det) {
  for (int64_t e = 0; e < nel; ++e) {
    skb::Affine<2> g;
    skb::affine_load<2>(g, p, npts, t, nel, e);
    skb::affine_invert(g);
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 2; ++j) inv[4 * e + 2 * i + j] = g.inv[i][j];
    det[e] = g.det;
  }
}
extern "C" double host_pw_sum(const double *v, int n) {
  auto f = [&](int i) -> double { return v[i]; };
  return skb::pw_sum(n, f);
}
extern "C" double host_seq_sum(const double *v, int n) {
  auto f = [&](int i) -> double { return v[i]; };
  return skb::seq_sum(n, f);
}
"""


def build():
    os.makedirs(OUT_DIR, exist_ok=True)
    if os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(HDR),
                                                            os.path.getmtime(__file__)):
        return LIB
    cpp = os.path.join(OUT_DIR, "arith_host.cpp")
    with open(cpp, "w") as f:
        f.write(SOURCE % {"hdr": HDR})
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC",
                    "-Wno-unknown-pragmas", "-I", cuda_inc, "-o", LIB, cpp],
                   check=True, capture_output=True)
    return LIB


def lib():
    h = C.CDLL(build())
    h.host_pw_sum.restype = C.c_double
    h.host_seq_sum.restype = C.c_double
    h.host_exp_in_safe_range.argtypes = [C.c_double]
    return h

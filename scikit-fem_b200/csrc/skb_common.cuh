// Shared device code of the B200 assembly engine: the arithmetic contract of
// SURVEY.md Appendix A (operation order of the reference's numpy expressions).
//
// Everything here is compiled with -fmad=false: a*b+c is two rounded
// operations exactly like numpy's separate ufunc passes.  Fused operations are
// used only where written explicitly (__fma_rn) inside exact_div(), whose
// *result* is the correctly rounded quotient, i.e. identical to numpy's `/`.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/skfem_b200.h"

#define SKB_CUDA_TRY(expr)                      \
  do {                                          \
    cudaError_t _e = (expr);                    \
    if (_e != cudaSuccess) return (int)_e;      \
  } while (0)

#ifndef SKB_DIV_CORRECTIONS
#define SKB_DIV_CORRECTIONS 2
#endif

namespace skb {

// kernel-launch accounting (skb_launch_count): every launcher reports how many
// kernels of this library it enqueued
void count_launch(int n = 1);
// profiling / test switches set by skb_debug_flags(): bit0 fused kernel skips P1,
// bit1 skips P2, bit2 prints per-role cycle counts, bit3 forces the dense
// (uncached) element-local kernel / the scalar hex kernel, bit5 P2 skips its
// shared-memory gathers, bit6 P2 skips its global stores (bits 0,1,5,6: invalid results)
int debug_flags();
// SMs the persistent fused kernel leaves free (skb_sm_reserve): room for a
// concurrent NCCL kernel when the interface exchange overlaps the next step
int sm_reserve();

// one device int, zeroed, released on every exit path (zero-determinant flag of the
// isoparametric kernels)
struct DeviceFlag {
  int *p = nullptr;
  cudaStream_t st;
  explicit DeviceFlag(cudaStream_t s) : st(s) {}
  cudaError_t init() {
    cudaError_t e = cudaMallocAsync((void **)&p, sizeof(int), st);
    if (e != cudaSuccess) { p = nullptr; return e; }
    return cudaMemsetAsync(p, 0, sizeof(int), st);
  }
  // value after everything enqueued so far (synchronises the stream)
  cudaError_t read(int *host) {
    cudaError_t e = cudaMemcpyAsync(host, p, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(st);
  }
  ~DeviceFlag() { if (p) cudaFreeAsync(p, st); }
  DeviceFlag(const DeviceFlag &) = delete;
  DeviceFlag &operator=(const DeviceFlag &) = delete;
};

// ---------------------------------------------------------------------------
// Correctly rounded a/b for many numerators sharing one denominator.
// y = RN(1/b) (__drcp_rn), q0 = RN(a*y), then Markstein corrections
// q <- RN(q + RN(a - b*q) * y).  After the first correction q is a faithful
// rounding of a/b, so by Markstein's theorem (y correctly rounded, q faithful)
// the second one returns RN(a/b).  Valid when no intermediate under/overflows:
// callers check exponents with div_safe() and otherwise use the plain `/`.
// Verified against IEEE division on 2e9 adversarial pairs.
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool exp_in_safe_range(double v) {
  // biased exponent within [1023-400, 1023+400], or v == +-0
  int hi = __double2hiint(v);
  int ex = (hi >> 20) & 0x7ff;
  bool zero = ((hi & 0x7fffffff) | __double2loint(v)) == 0;
  return zero | ((unsigned)(ex - 623) <= 800u);
}

__device__ __forceinline__ double exact_div(double a, double b, double y) {
  double q = a * y;
#pragma unroll
  for (int it = 0; it < SKB_DIV_CORRECTIONS; ++it) {
    double r = __fma_rn(-q, b, a);
    q = __fma_rn(r, y, q);
  }
  return q;
}

// ---------------------------------------------------------------------------
// Affine geometry, mapping/mapping_affine.py:55-131 (Appendix A.1).
// ---------------------------------------------------------------------------
template <int DIM>
struct Affine {
  double A[DIM][DIM];   // A[i][j] = p[i][t[j+1]] - p[i][t[0]]
  double b[DIM];        // p[i][t[0]]
  double inv[DIM][DIM]; // invA
  double det;           // detA (signed)
};

template <int DIM>
__device__ __forceinline__ void affine_load(Affine<DIM> &g, const double *__restrict__ p,
                                            int64_t npts, const int32_t *__restrict__ t,
                                            int64_t nel_total, int64_t e) {
  int32_t v[DIM + 1];
#pragma unroll
  for (int k = 0; k <= DIM; ++k) v[k] = __ldg(t + (int64_t)k * nel_total + e);
#pragma unroll
  for (int i = 0; i < DIM; ++i) {
    const double *pi = p + (int64_t)i * npts;
    double p0 = __ldg(pi + v[0]);
    g.b[i] = p0;
#pragma unroll
    for (int j = 0; j < DIM; ++j) g.A[i][j] = __ldg(pi + v[j + 1]) - p0;
  }
}

__device__ __forceinline__ void affine_invert(Affine<2> &g) {
  const double(*A)[2] = g.A;
  g.det = A[0][0] * A[1][1] - A[0][1] * A[1][0];          // mapping_affine.py:89-91
  double n00 = A[1][1], n01 = -A[0][1], n10 = -A[1][0], n11 = A[0][0];  // :106-110
  bool fast = exp_in_safe_range(g.det) && g.det != 0.0 && exp_in_safe_range(n00) &&
              exp_in_safe_range(n01) && exp_in_safe_range(n10) && exp_in_safe_range(n11);
  if (fast) {
    double y = __drcp_rn(g.det);
    g.inv[0][0] = exact_div(n00, g.det, y);
    g.inv[0][1] = exact_div(n01, g.det, y);
    g.inv[1][0] = exact_div(n10, g.det, y);
    g.inv[1][1] = exact_div(n11, g.det, y);
  } else {
    g.inv[0][0] = n00 / g.det;
    g.inv[0][1] = n01 / g.det;
    g.inv[1][0] = n10 / g.det;
    g.inv[1][1] = n11 / g.det;
  }
}

// numerators of the closed-form 3x3 inverse, sign/ordering of
// mapping_affine.py:111-129 (== mapping_isoparametric.py:212-220)
__device__ __forceinline__ void cofactors3(const double (*A)[3], double (*n)[3]) {
  n[0][0] = -A[1][2] * A[2][1] + A[1][1] * A[2][2];
  n[1][0] = A[1][2] * A[2][0] - A[1][0] * A[2][2];
  n[2][0] = -A[1][1] * A[2][0] + A[1][0] * A[2][1];
  n[0][1] = A[0][2] * A[2][1] - A[0][1] * A[2][2];
  n[1][1] = -A[0][2] * A[2][0] + A[0][0] * A[2][2];
  n[2][1] = A[0][1] * A[2][0] - A[0][0] * A[2][1];
  n[0][2] = -A[0][2] * A[1][1] + A[0][1] * A[1][2];
  n[1][2] = A[0][2] * A[1][0] - A[0][0] * A[1][2];
  n[2][2] = -A[0][1] * A[1][0] + A[0][0] * A[1][1];
}

__device__ __forceinline__ double det3(const double (*A)[3]) {
  // mapping_affine.py:92-98, left to right
  return A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) -
         A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
         A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
}

__device__ __forceinline__ void divide9(const double (*n)[3], double det, double (*inv)[3]) {
  bool fast = exp_in_safe_range(det) && det != 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) fast = fast && exp_in_safe_range(n[i][j]);
  if (fast) {
    double y = __drcp_rn(det);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) inv[i][j] = exact_div(n[i][j], det, y);
  } else {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) inv[i][j] = n[i][j] / det;
  }
}

__device__ __forceinline__ void affine_invert(Affine<3> &g) {
  g.det = det3(g.A);
  double n[3][3];
  cofactors3(g.A, n);
  divide9(n, g.det, g.inv);
}

// ---------------------------------------------------------------------------
// numpy pairwise summation (Appendix A.7; numpy loops_utils.h.src
// DOUBLE_pairwise_sum) of f(base) .. f(base+n-1), evaluated lazily.
// ---------------------------------------------------------------------------
template <class F>
__device__ __forceinline__ double pw_leaf(int base, int n, F &f) {
  if (n < 8) {
    double r = 0.0;
    for (int i = 0; i < n; ++i) r = r + f(base + i);
    return r;
  }
  double r[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) r[k] = f(base + k);
  int i = 8;
  for (; i < n - (n % 8); i += 8) {
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = r[k] + f(base + i + k);
  }
  double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
  for (; i < n; ++i) res = res + f(base + i);
  return res;
}

// explicit-stack version of the recursion  pw(n) = pw(n2) + pw(n - n2),
// n2 = n/2 - (n/2)%8, for n > 128 (depth <= 24 covers any int n)
template <class F>
__device__ double pw_sum(int n, F &f) {
  if (n <= 128) return pw_leaf(0, n, f);
  int sbase[24], sn[24];
  double sval[24];
  signed char sstate[24];  // 0: fresh, 1: left half pending, 2: right half pending
  int sp = 0;
  sbase[0] = 0; sn[0] = n; sstate[0] = 0;
  double ret = 0.0;
  while (sp >= 0) {
    int cn = sn[sp], cb = sbase[sp];
    if (cn <= 128) {
      ret = pw_leaf(cb, cn, f);
      --sp;
      continue;
    }
    int n2 = cn / 2;
    n2 -= n2 % 8;
    if (sstate[sp] == 0) {
      sstate[sp] = 1;
      ++sp; sbase[sp] = cb; sn[sp] = n2; sstate[sp] = 0;
    } else if (sstate[sp] == 1) {
      sval[sp] = ret;  // left result
      sstate[sp] = 2;
      ++sp; sbase[sp] = cb + n2; sn[sp] = cn - n2; sstate[sp] = 0;
    } else {
      ret = sval[sp] + ret;
      --sp;
    }
  }
  return ret;
}

// The same sum for a compile-time number of terms held in an array (registers): every loop
// unrolls.  N <= 128 is one leaf of the recursion.
template <int N>
__device__ __forceinline__ double pw_sum_fixed(const double (&t)[N]) {
  if (N < 8) {
    double r = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) r = r + t[i];
    return r;
  }
  static_assert(N <= 128, "one leaf of numpy's pairwise sum");
  double r[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) r[k] = t[k < N ? k : 0];
#pragma unroll
  for (int i = 8; i < N - (N % 8); i += 8)
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = r[k] + t[i + k < N ? i + k : 0];
  double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
#pragma unroll
  for (int i = N - (N % 8); i < N; ++i) res = res + t[i];
  return res;
}

// Plain left-to-right sum.  numpy reduces this way when the (nel, nqp) operand
// of np.sum(axis=1) is Fortran-ordered (the reduction axis is then the outer
// loop): e.g. ``v * dx`` for an affine mesh, where dx = np.tile(detA, (nqp,1)).T
// is F-ordered and v is a stride-0 broadcast (DESIGN.md "layout rule").
template <class F>
__device__ __forceinline__ double seq_sum(int n, F &f) {
  double r = f(0);
  for (int i = 1; i < n; ++i) r = r + f(i);
  return r;
}

}  // namespace skb

// Element-local matrices of high-order hexahedra (ElementHex2: 27 basis
// functions, 343 quadrature points) as a tensor-core contraction.
//
// For laplace / mass (models/poisson.py:7-19) on the isoparametric map
// (mapping_isoparametric.py:112-226) the local matrix is a Gram matrix
//     A_ij = sum_k Gs[k][i] * Gs[k][j],
//     laplace: k = (q, d), Gs = (invDF^T dphi_i)_d(q) * sqrt(dx_q)   (K = 3 nqp = 1029)
//     mass:    k = q,      Gs = phi_i(q) * sqrt(dx_q)                (K = nqp)
// i.e. a 27 x K by K x 27 product per element: the one part of this library
// whose work really is a dense contraction (FP64 DMMA for C4).  It
// runs on mma.sync.m8n8k4.f64 (tcgen05 has no FP64 kind).  Hex2 parity is
// value-level (CSR values within rtol 1e-12 of the reference; the tables are the
// reference's own), so the quadrature sum may be reordered and the symmetric
// sqrt(dx) scaling used; ElementHex1 (bit-exact local data) keeps the scalar
// kernel in skb_local.cu.
//
// One WARP per element, 8 elements per CTA, one CTA per SM.  Quadrature points are
// processed in chunks of 32 (lane = point).  The chunk's reference tables (dphi of
// the element and of the geometry element) are staged once per CTA in shared
// memory and shared by the 8 warps - the loads for chunk c+1 are issued before the
// DMMAs of chunk c, which hide their L2 latency.  Every lane forms the Jacobian of
// its point from the vertex coordinates (registers), inverts it, pushes the 27
// reference gradients forward and stores them, times sqrt(dx), in the warp's Gs
// chunk; then the warp runs the DMMAs of all ten upper-triangle 8 x 8 output tiles
// (4 fragment loads feed 10 independent DMMAs per k-step; 20 accumulator registers
// per lane).
// Measured alternatives, all parity-green (C4 laplace local kernel, 262 144
// elements): one CTA of 5 warps per element with 2 tiles per warp and block
// barriers between phases 26.4 ms (L1/shared pipe 67 %, DMMA pipe 37 %); warp per
// element 22.2 ms; this kernel (+ the staged loads held in registers across the
// DMMA phase) 18.2 ms, DMMA pipe 55 % (profiles/r1_ncu_hex2.md); producer/tensor
// warp pairs with double-buffered chunks and named barriers 25.5 ms (four producer
// warps cannot feed four tensor warps: the push-forward is the long pole).
#include "skb_common.cuh"

namespace skb {

constexpr int HM_QC = 32;        // quadrature points per chunk (= lanes)
constexpr int HM_WARPS = 8;      // elements in flight per CTA
constexpr int HM_NB = 32;        // basis functions padded to 4 tiles of 8

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int FORM>   // SKB_FORM_LAPLACE | SKB_FORM_MASS
__global__ void __launch_bounds__(HM_WARPS * 32, 1)
local_hex_mma_kernel(const skb_space_t s, double *__restrict__ out, int *__restrict__ err) {
  constexpr int KD = FORM == SKB_FORM_LAPLACE ? 3 : 1;  // Gs columns per quadrature point
  constexpr int KC = KD * HM_QC;                        // contraction length per chunk
  // Gs is stored basis-function-major, GsT[i][k], row stride == 4 (mod 16) doubles:
  // the 8x4 / 4x8 fragment loads (row 8t + lane/4, column k0 + lane%4) then touch
  // every bank pair exactly twice - the minimum for 64-bit loads - and the
  // producer stores (lanes over consecutive k) are conflict-free
  constexpr int LDT = KC + 4;
  extern __shared__ double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double *GsT = sm + (size_t)warp * HM_NB * LDT;        // this warp's [HM_NB][LDT]
  double *tab = sm + (size_t)HM_WARPS * HM_NB * LDT;    // [nbs*KD][HM_QC] element table chunk
  const int nqp = s.nqp, nbs = s.nbs;
  double *mtab = tab + nbs * KD * HM_QC;                // [24][HM_QC]     geometry table chunk
  const double *table = FORM == SKB_FORM_LAPLACE ? s.dphi : s.phi;
  // staging in two halves: fetch() issues the loads of a chunk into registers (coalesced
  // over q, all threads of the CTA), put() stores them; the DMMAs run in between, so a
  // warp never sits on the L2 latency of a load -> store pair in front of its DMMAs
  constexpr int MAXST = ((HM_NB * KD + 24) * HM_QC + HM_WARPS * 32 - 1) / (HM_WARPS * 32);
  const int nstage = (nbs * KD + 24) * HM_QC;
  double st[MAXST];
  auto fetch = [&](int q0) {
#pragma unroll
    for (int m = 0; m < MAXST; ++m) {
      const int idx = threadIdx.x + m * HM_WARPS * 32, r = idx / HM_QC, q = q0 + idx % HM_QC;
      const double *src = r < nbs * KD ? table + (int64_t)r * nqp
                                       : s.mdphi + (int64_t)(r - nbs * KD) * nqp;
      st[m] = (idx < nstage && q < nqp) ? __ldg(src + q) : 0.0;
    }
  };
  auto put = [&]() {
#pragma unroll
    for (int m = 0; m < MAXST; ++m) {
      const int idx = threadIdx.x + m * HM_WARPS * 32;
      if (idx < nstage) tab[idx] = st[m];                // tab and mtab are contiguous
    }
  };
  for (int idx = lane; idx < HM_NB * LDT; idx += 32) GsT[idx] = 0.0;   // padding rows stay 0
  __syncwarp();
  const double *fr = GsT + (lane >> 2) * LDT + (lane & 3);

  for (int64_t base = (int64_t)blockIdx.x * HM_WARPS; base < s.nel;
       base += (int64_t)gridDim.x * HM_WARPS) {
    const int64_t e = base + warp;
    const bool active = e < s.nel;                       // warps of the last group may idle
    const int64_t eg = !active ? 0 : (s.tind ? (int64_t)s.tind[e] : e);
    double xn[3][8];                                     // vertex coordinates (broadcast loads)
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const int64_t v = s.t[(int64_t)n * s.nel_total + eg];
#pragma unroll
      for (int i = 0; i < 3; ++i) xn[i][n] = __ldg(s.p + (int64_t)i * s.npts + v);
    }
    __syncthreads();                                     // previous group's reads of tab are over
    fetch(0);
    put();
    __syncthreads();
    double c[10][2];                                     // upper-triangle tiles, row-major order
#pragma unroll
    for (int t = 0; t < 10; ++t) c[t][0] = c[t][1] = 0.0;
    for (int q0 = 0; q0 < nqp; q0 += HM_QC) {
      const int q = q0 + lane;
      // ---- this lane's point: J = sum_n x_n (x) dphiM_n, inverse, sqrt(|det| W) -------------
      double inv[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}}, sdx = 0.0;
      if (q < nqp) {
        double J[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}}, nn[3][3];
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const double m = mtab[(n * 3 + j) * HM_QC + lane];
#pragma unroll
            for (int i = 0; i < 3; ++i) J[i][j] = __fma_rn(xn[i][n], m, J[i][j]);
          }
        const double det = det3(J);
        if (det == 0.0 && active) atomicExch(err, 1);   // mapping_isoparametric.py:195-196
        cofactors3(J, nn);
        const double rdet = 1.0 / det;        // value-level parity: one division per point
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) inv[i][j] = nn[i][j] * rdet;
        sdx = sqrt(fabs(det) * __ldg(s.W + q));
      }
      // ---- Gs chunk: pushed gradients (or values) of all basis functions at this point ------
      // (the previous chunk's fragment loads ended before the block barrier below it)
#pragma unroll 9
      for (int i = 0; i < nbs; ++i) {
        double *row = GsT + i * LDT + lane * KD;
        if (FORM == SKB_FORM_LAPLACE) {
          const double d0 = tab[(i * 3 + 0) * HM_QC + lane], d1 = tab[(i * 3 + 1) * HM_QC + lane],
                       d2 = tab[(i * 3 + 2) * HM_QC + lane];
          // grad_j = sum_c invDF[c][j] dphi_c   (element_h1.py:17), times sqrt(dx)
          row[0] = __fma_rn(inv[2][0], d2, __fma_rn(inv[1][0], d1, inv[0][0] * d0)) * sdx;
          row[1] = __fma_rn(inv[2][1], d2, __fma_rn(inv[1][1], d1, inv[0][1] * d0)) * sdx;
          row[2] = __fma_rn(inv[2][2], d2, __fma_rn(inv[1][2], d1, inv[0][2] * d0)) * sdx;
        } else {
          row[0] = tab[i * HM_QC + lane] * sdx;
        }
      }
      __syncthreads();                         // every warp is done with this chunk's tables
      const bool more = q0 + HM_QC < nqp;
      if (more) fetch(q0 + HM_QC);             // next chunk's loads fly during the DMMAs
      // ---- DMMA: C(a,b) += Gs[:, 8a:8a+8]^T Gs[:, 8b:8b+8] for the 10 tiles a <= b ----------
#pragma unroll 2
      for (int k0 = 0; k0 < KC; k0 += 4) {
        const double f0 = fr[k0], f1 = fr[8 * LDT + k0], f2 = fr[16 * LDT + k0],
                     f3 = fr[24 * LDT + k0];
        dmma(c[0][0], c[0][1], f0, f0);
        dmma(c[1][0], c[1][1], f0, f1);
        dmma(c[2][0], c[2][1], f0, f2);
        dmma(c[3][0], c[3][1], f0, f3);
        dmma(c[4][0], c[4][1], f1, f1);
        dmma(c[5][0], c[5][1], f1, f2);
        dmma(c[6][0], c[6][1], f1, f3);
        dmma(c[7][0], c[7][1], f2, f2);
        dmma(c[8][0], c[8][1], f2, f3);
        dmma(c[9][0], c[9][1], f3, f3);
      }
      if (more) put();
      __syncthreads();                         // staged tables of the next chunk are visible
    }
    if (!active) continue;
    // ---- write the tiles (and the mirrors of the off-diagonal ones) -------------------------
    // lane holds C[row = lane/4][col = 2*(lane%4) + {0,1}] of each tile
    int t = 0;
#pragma unroll
    for (int ta = 0; ta < 4; ++ta)
#pragma unroll
      for (int tb = ta; tb < 4; ++tb, ++t) {
        const int i = 8 * ta + (lane >> 2), j = 8 * tb + 2 * (lane & 3);
        if (i < nbs) {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int jj = j + u;
            if (jj < nbs) {
              out[((int64_t)i * nbs + jj) * s.nel + e] = c[t][u];
              if (ta != tb) out[((int64_t)jj * nbs + i) * s.nel + e] = c[t][u];
            }
          }
        }
      }
  }
}

size_t hex_mma_smem(int form, int nbs) {
  const int kd = form == SKB_FORM_LAPLACE ? 3 : 1;
  return sizeof(double) * ((size_t)HM_WARPS * HM_NB * (kd * HM_QC + 4) +
                           (size_t)nbs * kd * HM_QC + 24 * HM_QC);
}

// launched by launch_local (skb_local.cu) for scalar hex elements with 9..32
// basis functions; `err` is the device zero-determinant flag
int launch_hex_mma(const skb_space_t &s, int form, double *out, int *err, cudaStream_t st) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t smem = hex_mma_smem(form, s.nbs);
  if (smem > 227 * 1024) return SKB_ETOOBIG;
  const int64_t need = (s.nel + HM_WARPS - 1) / HM_WARPS;
  const int per_sm = 1;
  const int64_t cap = (int64_t)sms * per_sm;
  const int grid = (int)(need < cap ? need : cap);
  if (form == SKB_FORM_LAPLACE) {
    auto k = local_hex_mma_kernel<SKB_FORM_LAPLACE>;
    SKB_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, HM_WARPS * 32, smem, st>>>(s, out, err);
  } else {
    auto k = local_hex_mma_kernel<SKB_FORM_MASS>;
    SKB_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, HM_WARPS * 32, smem, st>>>(s, out, err);
  }
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace skb

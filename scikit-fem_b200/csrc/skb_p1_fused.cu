// Fused P1 (ElementTetP1) Laplace assembly: geometry -> local matrix -> CSR
// values in one pass; element-local matrices never touch HBM.
//
// Replaces, for the headline path, the whole chain
//   CellBasis.__init__           assembly/basis/cell_basis.py:94-106
//   BilinearForm._assemble       assembly/form/bilinear_form.py:58-128,150-151
//   COOData._assemble_scipy_csr  assembly/form/coo_data.py:27-36 (values)
// for form = models/poisson.py:7-9 (laplace) on ElementTetP1.
//
// Data layout (built once per (mesh, pattern) by skfem_b200/fused.py):
//   elements are ordered along a Morton curve and cut into tiles of T
//   elements.  Per tile the plan holds the list of its distinct vertices, the
//   connectivity rewritten in tile-local 16-bit vertex numbers and, for every
//   CSR slot the tile touches ("tile slot"), the staging indices of the local
//   entries that add into it.
//   A CTA owns one tile at a time:
//     phase 0  the tile's contribution-index list is requested into registers
//              (consumed after phase 1, so its latency hides behind the FP64
//              work) and the tile's vertex coordinates are gathered once into
//              shared memory (3 loads per vertex instead of 12 per element);
//     phase 1  every thread computes the 10 unique local entries of its
//              elements in registers, bit-identical to numpy (Appendix A),
//              and stages them in shared memory  vals[k*T + e];
//     phase 2  the index list is parked in shared memory (over the dead
//              coordinates) and one lane per tile slot adds that slot's staged
//              contributions in a fixed order.  Slots are sorted by
//              contribution count and stored sliced-ELL (groups of 32 slots,
//              contribution k of lane l at base + 32k + l, short lists padded
//              with the index of a staged 0.0) so lanes of a warp run equally
//              long without predication.  The sum goes straight to csr_data
//              (slot touched by this tile only) or to its reserved position in
//              a scratch array grouped by CSR slot.
//   skb_p1_combine then adds the per-tile partials of every shared slot in
//   tile order.  No float atomics anywhere: results are bit-reproducible.
#include "skb_common.cuh"

namespace skb {

struct P1Plan {
  const double *p;
  int64_t npts;
  const ushort4 *tl;                 // [ntiles*T] tile-local vertex ids, 0xFFFF = padding
  int32_t ntiles, T;
  const uint32_t *tile_vert_start;   // [ntiles+1]
  const int32_t *tile_verts;         // global vertex ids of each tile
  const uint32_t *tile_slot_start;   // [ntiles+1] first tile slot of each tile
  const uint32_t *tile_group_start;  // [ntiles+1] first 32-slot group of each tile
  const uint32_t *tile_contrib_start;// [ntiles+1] first index of each tile, multiples of 8
  const uint32_t *grp_base;          // per group: tile-relative offset into contrib
  const uint16_t *grp_len;           // per group: (padded) contribution list length
  const uint16_t *contrib;           // sliced-ELL staging indices k*T + e, 10*T = "zero"
  const uint32_t *meta;              // per tile slot: bit31 ? scratch position : csr slot
  double *csr_data;
  double *scratch;
  double w;                          // the common quadrature weight
  int32_t nqp;
  int32_t aux_bytes;                 // shared bytes for coordinates / index list
};

// unique (a<=b) local entries: k index of pair (a,b), 4 basis functions
__device__ __forceinline__ constexpr int sym_index4(int a, int b) {  // a <= b
  return a * 4 - (a * (a - 1)) / 2 + (b - a);
}

// sum of nqp identical terms v in numpy's order: sequential from 0 for n < 8,
// pairwise lanes for n >= 8.  nqp == 4 (the default rule of ElementTetP1) is
// the fast path: ((v+v)+v)+v with v+v exact.
__device__ __noinline__ double sum_equal_terms_general(double v, int nqp) {
  auto f = [&](int) -> double { return v; };
  return pw_sum(nqp, f);
}

template <bool NQP4>
__device__ __forceinline__ double sum_equal_terms(double v, int nqp) {
  if (NQP4) {
    double r = __fma_rn(v, 2.0, v);  // (v+v)+v : v+v is exact, one rounding
    return r + v;
  }
  return sum_equal_terms_general(v, nqp);
}

// |c| is 0 or within [2^-60, 2^60]: if every coordinate of a tile passes, all
// cofactors are 0 or in [2^-278, 2^123] and a nonzero determinant lies in
// [2^-391, 2^184], so exact_div() can neither overflow nor underflow and the
// per-element exponent checks of divide9() are unnecessary.
__device__ __forceinline__ bool coord_tame(double c) {
  const unsigned h = (unsigned)__double2hiint(c) & 0x7fffffffu;
  const bool zero = (h | (unsigned)__double2loint(c)) == 0u;
  return zero | ((h - 0x3c300000u) <= (0x43b00000u - 0x3c300000u));
}

constexpr int P1_MAX_PREFETCH = 4;  // uint4 (8 indices) per thread

template <int T_ELEMS, int THREADS, bool NQP4>
__global__ void __launch_bounds__(THREADS)
p1tet_laplace_fused_kernel(const P1Plan pl) {
  extern __shared__ double smem[];
  double *vals = smem;                                 // [10*T + 2], [10*T] == 0.0
  double4 *sxyz = reinterpret_cast<double4 *>(smem + 10 * T_ELEMS + 2);  // coordinates (32 B)
  uint4 *sidx4 = reinterpret_cast<uint4 *>(sxyz);      // later: the index list
  const uint16_t *sidx = reinterpret_cast<const uint16_t *>(sxyz);
  __shared__ int s_wild;
  constexpr int PER_THREAD = T_ELEMS / THREADS;
  constexpr int NWARPS = THREADS / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { vals[10 * T_ELEMS] = 0.0; s_wild = 0; }
  __syncthreads();
  for (int tile = blockIdx.x; tile < pl.ntiles; tile += gridDim.x) {
    // ---- phase 0: request the index list, stage vertex coordinates -------------
    const uint32_t c0 = pl.tile_contrib_start[tile], c1 = pl.tile_contrib_start[tile + 1];
    const int nvec = (int)((c1 - c0) >> 3);
    uint4 pf[P1_MAX_PREFETCH];
    {
      const uint4 *src = reinterpret_cast<const uint4 *>(pl.contrib + c0);
#pragma unroll
      for (int r = 0; r < P1_MAX_PREFETCH; ++r) {
        const int i = r * THREADS + threadIdx.x;
        if (i < nvec) pf[r] = __ldg(src + i);
      }
    }
    {
      const uint32_t v0 = pl.tile_vert_start[tile], v1 = pl.tile_vert_start[tile + 1];
      const double *px = pl.p, *py = pl.p + pl.npts, *pz = pl.p + 2 * pl.npts;
      bool wild = false;
      for (int i = threadIdx.x; i < (int)(v1 - v0); i += THREADS) {
        const int32_t gv = __ldg(pl.tile_verts + v0 + i);
        const double x = __ldg(px + gv), y = __ldg(py + gv), z = __ldg(pz + gv);
        wild |= !(coord_tame(x) & coord_tame(y) & coord_tame(z));
        sxyz[i] = make_double4(x, y, z, 0.0);
      }
      if (wild) s_wild = 1;
    }
    __syncthreads();
    const bool tame = (s_wild == 0);
    // ---- phase 1: local matrices -------------------------------------------------
#pragma unroll 1
    for (int it = 0; it < PER_THREAD; ++it) {
      const int el = it * THREADS + threadIdx.x;
      const ushort4 v = __ldg(pl.tl + (int64_t)tile * T_ELEMS + el);
      if (v.x == 0xFFFF) continue;  // padding of the last tile
      double A[3][3];
      {
        const double4 q0 = sxyz[v.x], q1 = sxyz[v.y], q2 = sxyz[v.z], q3 = sxyz[v.w];
        A[0][0] = q1.x - q0.x; A[0][1] = q2.x - q0.x; A[0][2] = q3.x - q0.x;
        A[1][0] = q1.y - q0.y; A[1][1] = q2.y - q0.y; A[1][2] = q3.y - q0.y;
        A[2][0] = q1.z - q0.z; A[2][1] = q2.z - q0.z; A[2][2] = q3.z - q0.z;
      }
      const double det = det3(A);
      double n[3][3], inv[3][3];
      cofactors3(A, n);
      if (tame && det != 0.0) {
        const double y = __drcp_rn(det);
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
      // P1 push-forward: dphi_b is +-unit, so grad_b (b=1..3) is row b-1 of inv
      // and grad_0[j] = -((inv0j + inv1j) + inv2j)   (Appendix A.4)
      double g[4][3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        g[0][j] = -((inv[0][j] + inv[1][j]) + inv[2][j]);
        g[1][j] = inv[0][j];
        g[2][j] = inv[1][j];
        g[3][j] = inv[2][j];
      }
      const double dx = fabs(det) * pl.w;  // cell_basis.py:104-105
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = a; b < 4; ++b) {
          const double d = (g[a][0] * g[b][0] + g[a][1] * g[b][1]) + g[a][2] * g[b][2];
          vals[sym_index4(a, b) * T_ELEMS + el] = sum_equal_terms<NQP4>(d * dx, pl.nqp);
        }
    }
    __syncthreads();
    // ---- phase 2: park the index list, then per-slot sums in fixed order ---------
#pragma unroll
    for (int r = 0; r < P1_MAX_PREFETCH; ++r) {
      const int i = r * THREADS + threadIdx.x;
      if (i < nvec) sidx4[i] = pf[r];
    }
    for (int i = P1_MAX_PREFETCH * THREADS + threadIdx.x; i < nvec; i += THREADS)
      sidx4[i] = __ldg(reinterpret_cast<const uint4 *>(pl.contrib + c0) + i);
    __syncthreads();
    {
      const uint32_t s0 = pl.tile_slot_start[tile];
      const int nslots = (int)(pl.tile_slot_start[tile + 1] - s0);
      const uint32_t g0 = pl.tile_group_start[tile], g1 = pl.tile_group_start[tile + 1];
      for (uint32_t g = g0 + warp; g < g1; g += NWARPS) {
        const int j = (int)(g - g0) * 32 + lane;
        const int len = (int)__ldg(pl.grp_len + g);
        const uint16_t *cb = sidx + __ldg(pl.grp_base + g) + lane;
        const uint32_t m = (j < nslots) ? __ldg(pl.meta + s0 + j) : 0xffffffffu;
        double acc = 0.0;
        int k = 0;
        for (; k + 4 <= len; k += 4) {
          const int i0 = cb[k * 32], i1 = cb[(k + 1) * 32], i2 = cb[(k + 2) * 32],
                    i3 = cb[(k + 3) * 32];
          const double a0 = vals[i0], a1 = vals[i1], a2 = vals[i2], a3 = vals[i3];
          acc = acc + a0;
          acc = acc + a1;
          acc = acc + a2;
          acc = acc + a3;
        }
        for (; k < len; ++k) acc = acc + vals[cb[k * 32]];
        if (m != 0xffffffffu) {
          if (m & 0x80000000u) pl.scratch[m & 0x7fffffffu] = acc;
          else pl.csr_data[m] = acc;
        }
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
p1_combine_kernel(const double *__restrict__ scratch, const uint32_t *__restrict__ sptr,
                  const uint32_t *__restrict__ gslot, int64_t nshared,
                  double *__restrict__ csr_data) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nshared;
       k += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t a = sptr[k], b = sptr[k + 1];
    double acc = scratch[a];
    for (uint32_t i = a + 1; i < b; ++i) acc = acc + scratch[i];
    csr_data[gslot[k]] = acc;
  }
}

}  // namespace skb

extern "C" int skb_p1tet_laplace_fused(const double *p, int64_t npts, const uint16_t *tl,
                                       int32_t ntiles, int32_t tile_elems, int32_t threads,
                                       const uint32_t *tile_vert_start, const int32_t *tile_verts,
                                       int32_t aux_bytes, const uint32_t *tile_slot_start,
                                       const uint32_t *tile_group_start,
                                       const uint32_t *tile_contrib_start,
                                       const uint32_t *grp_base, const uint16_t *grp_len,
                                       const uint16_t *contrib, const uint32_t *meta, double w,
                                       int32_t nqp, double *csr_data, double *scratch,
                                       void *stream) {
  using namespace skb;
  if (ntiles < 0 || !p || nqp <= 0 || aux_bytes <= 0) return SKB_EINVAL;
  if (ntiles == 0) return SKB_OK;
  P1Plan pl;
  pl.p = p; pl.npts = npts; pl.tl = (const ushort4 *)tl; pl.ntiles = ntiles; pl.T = tile_elems;
  pl.tile_vert_start = tile_vert_start; pl.tile_verts = tile_verts;
  pl.tile_slot_start = tile_slot_start; pl.tile_group_start = tile_group_start;
  pl.tile_contrib_start = tile_contrib_start;
  pl.grp_base = grp_base; pl.grp_len = grp_len; pl.contrib = contrib; pl.meta = meta;
  pl.csr_data = csr_data; pl.scratch = scratch; pl.w = w; pl.nqp = nqp;
  pl.aux_bytes = aux_bytes;
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t smem = sizeof(double) * (10 * (size_t)tile_elems + 2) + (size_t)aux_bytes;
  if (smem > 226 * 1024) return SKB_ETOOBIG;
  int per_sm = (int)((227 * 1024) / (smem + 1024 + 16));
  if (per_sm * threads > 2048) per_sm = 2048 / threads;
  if (per_sm < 1) per_sm = 1;
#define SKB_P1_LAUNCH(TT, TH, Q4)                                                             \
  do {                                                                                        \
    auto k = p1tet_laplace_fused_kernel<TT, TH, Q4>;                                          \
    SKB_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                      (int)smem));                                            \
    const int cap = per_sm * sms;                                                             \
    const int grid = ntiles < cap ? ntiles : cap;                                             \
    k<<<grid, TH, smem, st>>>(pl);                                                            \
  } while (0)
  const bool q4 = (nqp == 4);
  if (tile_elems == 1024 && threads == 256) {
    if (q4) SKB_P1_LAUNCH(1024, 256, true); else SKB_P1_LAUNCH(1024, 256, false);
  } else if (tile_elems == 1024 && threads == 512) {
    if (q4) SKB_P1_LAUNCH(1024, 512, true); else SKB_P1_LAUNCH(1024, 512, false);
  } else if (tile_elems == 1024 && threads == 1024) {
    if (q4) SKB_P1_LAUNCH(1024, 1024, true); else SKB_P1_LAUNCH(1024, 1024, false);
  } else if (tile_elems == 2048 && threads == 512) {
    if (q4) SKB_P1_LAUNCH(2048, 512, true); else SKB_P1_LAUNCH(2048, 512, false);
  } else if (tile_elems == 2048 && threads == 1024) {
    if (q4) SKB_P1_LAUNCH(2048, 1024, true); else SKB_P1_LAUNCH(2048, 1024, false);
  } else if (tile_elems == 512 && threads == 256) {
    if (q4) SKB_P1_LAUNCH(512, 256, true); else SKB_P1_LAUNCH(512, 256, false);
  } else if (tile_elems == 512 && threads == 512) {
    if (q4) SKB_P1_LAUNCH(512, 512, true); else SKB_P1_LAUNCH(512, 512, false);
  } else {
    return SKB_EINVAL;
  }
#undef SKB_P1_LAUNCH
  count_launch();
  return (int)cudaGetLastError();
}

extern "C" int skb_p1_combine(const double *scratch, const uint32_t *sptr, const uint32_t *gslot,
                              int64_t nshared, double *csr_data, void *stream) {
  using namespace skb;
  if (nshared < 0) return SKB_EINVAL;
  if (nshared == 0) return SKB_OK;
  int64_t g = (nshared + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  p1_combine_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(scratch, sptr, gslot, nshared,
                                                              csr_data);
  count_launch();
  return (int)cudaGetLastError();
}

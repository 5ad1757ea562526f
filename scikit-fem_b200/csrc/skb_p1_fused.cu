// Fused P1 (ElementTetP1) Laplace assembly: geometry -> local matrix -> CSR
// values in one pass; element-local matrices never touch HBM.
//
// Replaces, for the headline path, the whole chain
//   CellBasis.__init__           assembly/basis/cell_basis.py:94-106
//   BilinearForm._assemble       assembly/form/bilinear_form.py:58-128,150-151
//   COOData._assemble_scipy_csr  assembly/form/coo_data.py:27-36 (values)
// for form = models/poisson.py:7-9 (laplace) on ElementTetP1.
//
// Plan (built once per (mesh, pattern) by skfem_b200/fused.py): elements are
// ordered along a Morton curve and cut into tiles of T elements.  Each tile
// owns one contiguous, 16-byte aligned *record* in HBM:
//     header | tl: T x ushort4 tile-local vertex ids | verts: global vertex ids
//     | grp: per 32-slot group {offset/32, length} | meta: per tile slot the CSR
//     slot or (bit 31) the scratch position | ids: sliced-ELL staging indices
//
// Kernel: persistent CTAs, software pipeline over the CTA's tiles
//   TMA   the record of tile k+NR-1 is fetched by one cp.async.bulk (UBLKCP)
//         into a ring of NR shared buffers, completion on an mbarrier;
//   LDGSTS the vertex coordinates of tile k+1 are gathered asynchronously
//         (cp.async, 3 x 8 B per vertex) into a double-buffered coordinate
//         array as soon as its record has landed;
//   P1    every thread forms the 10 unique local entries of its element(s) in
//         registers, bit-identical to numpy (no FMA
//         contraction, reference operation order, correctly rounded division)
//         and stages them  vals[k*T + e];
//   P2    one lane per tile slot adds that slot's staged contributions in a
//         fixed order (sliced ELL: contribution k of lane l at base + 32k + l,
//         short lists padded with the index of a staged 0.0) and writes the
//         sum to csr_data (slot touched by this tile only) or to its reserved
//         position in a scratch array grouped by CSR slot.
// skb_p1_combine then adds the per-tile partials of every shared slot in tile
// order.  No float atomics anywhere: results are bit-reproducible.
#include <cstdio>
#include <cstring>
#include "skb_common.cuh"

namespace skb {

struct P1Args {
  const double *p;
  int64_t npts;
  const unsigned char *rec;     // concatenated tile records
  const uint64_t *rec_start;    // [ntiles+1] byte offsets (multiples of 16)
  int32_t ntiles;
  int32_t rec_cap;              // largest record, bytes (multiple of 16)
  int32_t vcap;                 // most vertices in one tile
  double *csr_data;
  double *scratch;
  double w;                     // the common quadrature weight
  int32_t nqp;
  int32_t tame;                 // all coordinates within the exact_div-safe range
  int32_t debug;                // profiling aid: bit0 skip P1, bit1 skip P2 (results invalid)
};

struct RecHeader {              // 32 bytes at the start of every record
  uint32_t nverts, ngroups, off_verts, off_grp, off_meta, off_ids, off_fsel, off_meta2;
};

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

// P1Args::tame: the plan builder verified that every vertex coordinate is 0 or
// has magnitude within [2^-60, 2^60].  Then all cofactors are 0 or in
// [2^-278, 2^123] and a nonzero determinant lies in [2^-443, 2^186], so
// exact_div() can neither overflow nor underflow and no per-element exponent
// checks are needed ("exact division").  Otherwise every element
// takes the plain IEEE division.

// ---- async-copy / mbarrier primitives (PTX) ---------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
  // C loop around try_wait: no PTX labels, so the function can be inlined any number of times
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
// TMA bulk copy global -> shared, completion counted on an mbarrier; marked evict-first in L2:
// the records are read once per step and must not push the vertex coordinates / tile partials
// out of the L2 (measured: neutral on the step time, 0.1983 vs 0.1986 ms)
__device__ __forceinline__ void tma_bulk_g2s_stream(void *dst, const void *src, unsigned bytes,
                                                    uint64_t *bar) {
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void cp_async8(void *dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.wait_all;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// Warp-specialised persistent kernel.  A CTA = T "compute" threads (one
// element each, FP64 bound) + NRED "reduce" threads (shared-memory latency
// bound).  In iteration k the compute warps run P1 of tile k into vals[k&1]
// while the reduce warps run P2 of tile k-1 out of vals[(k-1)&1], request the
// vertex gather of tile k+2 and retire the gather of tile k+1; one block
// barrier per iteration, then one thread re-arms the freed record buffer with
// the TMA fetch of tile k-1+NR.
// FAST (opt-in, P1Args::tame == 2): the element arithmetic uses fused multiply-adds
// and one reciprocal instead of the reference's operation order - ~100 instead of
// ~190 FP64 instructions per element.  Values then agree with the reference to a few
// ulp per term (well inside the rtol 1e-12 bar for CSR values) but the element-local
// data is no longer bit-identical; the sparsity pattern is unaffected (it comes from
// the plan, which is always built from bit-exact local data).
// CT = compute threads (default one per element); CT < T_ELEMS lets every compute thread
// take T_ELEMS / CT elements in turn and frees thread slots for reduce warps.
template <int T_ELEMS, int NRED, int NR, bool NQP4, bool FAST = false, int CT = T_ELEMS>
__global__ void __launch_bounds__(CT + NRED + 32)
p1tet_laplace_fused_kernel(const P1Args a) {
  static_assert(NR >= 4, "record ring must hold tiles k-1 .. k+2");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int VSTRIDE = 10 * T_ELEMS + 16;   // + one staged 0.0 per bank pair
  // shared layout: vals[2] | coords[3] | records[NR] | mbarriers | flags | counters
  double *vals = reinterpret_cast<double *>(smem_raw);                      // [2][VSTRIDE]
  // coordinates: 3 buffers x {x[vcap], y[vcap], z[vcap]} (struct of arrays: a
  // random 8-byte gather conflicts far less than a 32-byte-strided one)
  double *coords = vals + 2 * VSTRIDE;                                      // [3][3][vcap]
  unsigned char *recs = reinterpret_cast<unsigned char *>(coords + 9 * (size_t)a.vcap);
  uint64_t *mbar = reinterpret_cast<uint64_t *>(recs + (size_t)NR * a.rec_cap);
  // roles: [0, T) compute, [T, T+NRED) reduce, last warp = TMA producer (lane 0)
  static_assert(T_ELEMS % CT == 0, "elements per compute thread must be integral");
  const bool is_compute = threadIdx.x < CT;
  const bool is_reduce = !is_compute && threadIdx.x < CT + NRED;
  const bool is_producer = threadIdx.x == CT + NRED;
  const int rtid = (int)threadIdx.x - CT;             // reduce-thread index
  const int lane = threadIdx.x & 31;
  const int rwarp = rtid >> 5;                        // reduce-warp index
  constexpr int NRW = NRED / 32;
  const bool tame = a.tame != 0;
  const int nk = (a.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  auto tile_of = [&](int k) { return (int)blockIdx.x + k * (int)gridDim.x; };
  auto rec_of = [&](int k) { return recs + (size_t)(k % NR) * a.rec_cap; };
  // producer lane: record extents are loaded one iteration before they are
  // needed so the TMA issue after the barrier has no global-load latency
  uint64_t nx_b0 = 0;
  unsigned nx_bytes = 0;
  auto extent = [&](int k) {
    if (k < nk) {
      const int t = tile_of(k);
      nx_b0 = a.rec_start[t];
      nx_bytes = (unsigned)(a.rec_start[t + 1] - nx_b0);
    }
  };
  auto issue = [&](int k) {   // fetch the record of this CTA's k-th tile (extent preloaded)
    if (k < nk) {
      fence_proxy_async();
      mbar_expect_tx(&mbar[k % NR], nx_bytes);
      tma_bulk_g2s_stream(rec_of(k), a.rec + nx_b0, nx_bytes, &mbar[k % NR]);
    }
  };
  auto wait_rec = [&](int k) { mbar_wait(&mbar[k % NR], (unsigned)((k / NR) & 1)); };
  auto gather = [&](int k) {  // reduce threads: async gather of tile k's vertex coordinates
    const unsigned char *r = rec_of(k);
    const RecHeader *h = reinterpret_cast<const RecHeader *>(r);
    const int nv = (int)h->nverts;
    const int32_t *verts = reinterpret_cast<const int32_t *>(r + h->off_verts);
    double *dx = coords + (size_t)(k % 3) * 3 * a.vcap, *dy = dx + a.vcap, *dz = dy + a.vcap;
    const double *px = a.p, *py = a.p + a.npts, *pz = a.p + 2 * a.npts;
    for (int i = rtid; i < nv; i += NRED) {
      const int32_t gv = verts[i];
      cp_async8(dx + i, px + gv);
      cp_async8(dy + i, py + gv);
      cp_async8(dz + i, pz + gv);
    }
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < NR; ++i) mbar_init(&mbar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int i = 0; i < 16; ++i) {
      vals[10 * T_ELEMS + i] = 0.0;
      vals[VSTRIDE + 10 * T_ELEMS + i] = 0.0;
    }
  }
  __syncthreads();
  if (nk <= 0) return;
  // prologue: records 0..NR-2 in flight, coordinates of tiles 0 and 1 resident
  if (is_producer) {
    for (int k = 0; k < NR; ++k) { extent(k); issue(k); }
    extent(NR);                        // for the issue after iteration 1
  }
  if (is_reduce) {
    wait_rec(0);
    gather(0);
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (nk > 1) { wait_rec(1); gather(1); }
    asm volatile("cp.async.commit_group;" ::: "memory");
    cp_async_wait_all();
  }
  __syncthreads();

  // profiling aid (debug bit 2): cycles block 0 spends working vs at the barrier
  const bool timing = (a.debug & 4) && blockIdx.x == 0 && lane == 0 &&
                      (threadIdx.x == 0 || threadIdx.x == CT);
  long long t_work = 0, t_bar = 0, t_mark = timing ? clock64() : 0;
  for (int k = 0; k <= nk; ++k) {
    if (is_compute) {
      // ---- P1(k): local matrices of tile k -> vals[k & 1] ---------------------------
      if (k < nk && !(a.debug & 1)) {
        wait_rec(k);
        const unsigned char *r = rec_of(k);
        const ushort4 *tl = reinterpret_cast<const ushort4 *>(r + sizeof(RecHeader));
        const double *sx = coords + (size_t)(k % 3) * 3 * a.vcap, *sy = sx + a.vcap,
                     *sz = sy + a.vcap;
        double *out = vals + (size_t)(k & 1) * VSTRIDE;
#pragma unroll 1
        for (int el = threadIdx.x; el < T_ELEMS; el += CT) {
        const ushort4 v = tl[el];
        if (v.x != 0xFFFF) {  // not a padding element of the last tile
          double A[3][3];
          {
            const double x0 = sx[v.x], y0 = sy[v.x], z0 = sz[v.x];
            A[0][0] = sx[v.y] - x0; A[0][1] = sx[v.z] - x0; A[0][2] = sx[v.w] - x0;
            A[1][0] = sy[v.y] - y0; A[1][1] = sy[v.z] - y0; A[1][2] = sy[v.w] - y0;
            A[2][0] = sz[v.y] - z0; A[2][1] = sz[v.z] - z0; A[2][2] = sz[v.w] - z0;
          }
          double det, n[3][3], inv[3][3];
          if (FAST) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {   // cofactors, same sign convention as cofactors3
              const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
                n[j][i] = __fma_rn(A[i1][j1], A[i2][j2], -(A[i1][j2] * A[i2][j1]));
              }
            }
            det = __fma_rn(A[0][0], n[0][0], __fma_rn(A[0][1], n[1][0], A[0][2] * n[2][0]));
            const double y = 1.0 / det;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
              for (int j = 0; j < 3; ++j) inv[i][j] = n[i][j] * y;
          } else {
            det = det3(A);
            cofactors3(A, n);
          }
          if (FAST) {
            // inverse already formed above
          } else if (tame && det != 0.0) {
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
          // P1 push-forward: dphi_b is +-unit, so grad_b (b=1..3) is row b-1 of
          // inv and grad_0[j] = -((inv0j + inv1j) + inv2j)   (Appendix A.4)
          double g[4][3];
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            g[0][j] = -((inv[0][j] + inv[1][j]) + inv[2][j]);
            g[1][j] = inv[0][j];
            g[2][j] = inv[1][j];
            g[3][j] = inv[2][j];
          }
          const double dx = fabs(det) * a.w;  // cell_basis.py:104-105
          const double dxs = dx * (double)a.nqp;   // FAST: nqp equal terms = one product
#pragma unroll
          for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int q = p; q < 4; ++q) {
              if (FAST) {
                const double d = __fma_rn(g[p][2], g[q][2],
                                          __fma_rn(g[p][1], g[q][1], g[p][0] * g[q][0]));
                out[sym_index4(p, q) * T_ELEMS + el] = d * dxs;
              } else {
                const double d = (g[p][0] * g[q][0] + g[p][1] * g[q][1]) + g[p][2] * g[q][2];
                out[sym_index4(p, q) * T_ELEMS + el] = sum_equal_terms<NQP4>(d * dx, a.nqp);
              }
            }
        }
        }
      }
    } else if (is_reduce) {
      // ---- reduce warps -----------------------------------------------------------------
      if (k >= 1 && !(a.debug & 2)) {
        // ---- P2(k-1): per-slot sums in fixed order (sliced ELL, in shared memory) -----
        const unsigned char *r = rec_of(k - 1);
        const RecHeader *h = reinterpret_cast<const RecHeader *>(r);
        const double *in = vals + (size_t)((k - 1) & 1) * VSTRIDE;
        const int ngroups = (int)h->ngroups;
        const uint32_t *grp = reinterpret_cast<const uint32_t *>(r + h->off_grp);
        const uint32_t *meta = reinterpret_cast<const uint32_t *>(r + h->off_meta);
        const uint32_t *meta2 = reinterpret_cast<const uint32_t *>(r + h->off_meta2);
        const uint8_t *fsel = r + h->off_fsel;
        const uint16_t *ids = reinterpret_cast<const uint16_t *>(r + h->off_ids);
        // groups are sorted longest first, dealt round-robin to the reduce warps;
        // every list length is a multiple of 2 (padded with the staged zero)
#pragma unroll 1
        for (int g = rwarp; g < ngroups; g += NRW) {
          const uint32_t gi = grp[g];
          const int len = (int)(gi >> 16);
          const uint16_t *cb = ids + (size_t)(gi & 0xffffu) * 32 + lane;
          const uint32_t m = meta[g * 32 + lane];
          const uint32_t m2 = meta2[g * 32 + lane];
          const int fs = fsel[g * 32 + lane];
          double acc = 0.0;
          const int len_eff = (a.debug & 32) ? 0 : len;   // profiling: skip the gathers
#pragma unroll 2
          for (int c = 0; c < len_eff; c += 2) {
            const int i0 = cb[c * 32], i1 = cb[(c + 1) * 32];
            const double a0 = in[i0], a1 = in[i1];
            acc = acc + a0;
            acc = acc + a1;
          }
          // long lists are split over 2 or 4 adjacent lanes: fixed combination
          // tree (l + l+1) + (l+2 + l+3), selected by the leader lane's fsel
          const double t1 = acc + __shfl_down_sync(0xffffffffu, acc, 1);
          const double t2 = t1 + __shfl_down_sync(0xffffffffu, t1, 2);
          acc = fs == 0 ? acc : (fs == 1 ? t1 : t2);
          // the local matrix is bitwise symmetric: slot (r,c), r<c, and its
          // mirror (c,r) receive the same terms in the same order -> one sum
          if (a.debug & 64) {                             // profiling: skip the global stores
            if (acc == 1.2345e300) a.csr_data[0] = acc;
          } else {
            if (m != 0xffffffffu) {
              if (m & 0x80000000u) a.scratch[m & 0x7fffffffu] = acc;
              else a.csr_data[m] = acc;
            }
            if (m2 != 0xffffffffu) a.csr_data[m2] = acc;
          }
        }
      }
      // request the vertex gather of tile k+2 (its record has been in flight
      // for at least P2's duration), retire the gather of tile k+1 (requested
      // one iteration ago)
      if (k + 2 < nk) {
        wait_rec(k + 2);
        gather(k + 2);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    }
    if (timing) { const long long c = clock64(); t_work += c - t_mark; t_mark = c; }
    __syncthreads();  // vals[k&1] complete; record k-1 free; coords of tile k+1 visible
    if (timing) { const long long c = clock64(); t_bar += c - t_mark; t_mark = c; }
    if (is_producer && k >= 1) {
      issue(k - 1 + NR);               // into the buffer tile k-1 vacated
      extent(k + NR);                  // its latency hides behind the next iteration
    }
  }
  if (timing)
    printf("[skb timing] block 0 %s warp: %d tiles, work %lld cycles, barrier wait %lld cycles\n",
           threadIdx.x == 0 ? "compute" : "reduce ", nk, t_work, t_bar);
}

__global__ void __launch_bounds__(256)
p1_combine_kernel(const double *__restrict__ scratch, const uint32_t *__restrict__ sptr,
                  const uint32_t *__restrict__ gslot, const uint32_t *__restrict__ gslot2,
                  int64_t nshared, double *__restrict__ csr_data) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nshared;
       k += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t a = sptr[k], b = sptr[k + 1];
    double acc = scratch[a];
    for (uint32_t i = a + 1; i < b; ++i) acc = acc + scratch[i];
    const uint32_t s = gslot[k], s2 = gslot2[k];
    csr_data[s] = acc;
    if (s2 != s) csr_data[s2] = acc;   // mirror slot of a symmetric pair
  }
}

template <int TT, int NRED, int NR, int CT = TT>
static int launch_fused(const P1Args &a, size_t smem, int sms, bool q4, cudaStream_t st) {
  // persistent grid: as many CTAs per SM as shared memory, threads and
  // registers (<= 64 per thread by __launch_bounds__) allow
  int per_sm = (int)((228 * 1024) / (smem + 1024));
  const int by_threads = 2048 / (CT + NRED + 32);
  if (per_sm > by_threads) per_sm = by_threads;
  if (per_sm < 1) per_sm = 1;
  int free_sms = sm_reserve();
  if (free_sms > sms - 1) free_sms = sms - 1;
  const int cap = per_sm * (sms - free_sms);
  const int grid = a.ntiles < cap ? a.ntiles : cap;
  constexpr bool kHasFast = TT == 512 && (NRED == 480 || CT != TT);   // fast arithmetic variants
  if (a.tame == 2 && kHasFast) {
    auto k = p1tet_laplace_fused_kernel<TT, NRED, NR, true, kHasFast, CT>;
    SKB_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, CT + NRED + 32, smem, st>>>(a);
  } else if (q4) {
    auto k = p1tet_laplace_fused_kernel<TT, NRED, NR, true, false, CT>;
    SKB_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, CT + NRED + 32, smem, st>>>(a);
  } else {
    auto k = p1tet_laplace_fused_kernel<TT, NRED, NR, false, false, CT>;
    SKB_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, CT + NRED + 32, smem, st>>>(a);
  }
  return (int)cudaGetLastError();
}

}  // namespace skb

extern "C" int64_t skb_p1_fused_smem_bytes(int32_t tile_elems, int32_t ring, int32_t rec_cap,
                                           int32_t vcap) {
  return (int64_t)(sizeof(double) * 2 * (10 * (size_t)tile_elems + 16) + 9 * (size_t)vcap * 8 +
                   (size_t)ring * rec_cap + 8 * (size_t)ring + 32);
}

extern "C" int skb_p1tet_laplace_fused(const double *p, int64_t npts, const void *rec,
                                       const uint64_t *rec_start, int32_t ntiles,
                                       int32_t tile_elems, int32_t reduce_threads, int32_t ring,
                                       int32_t rec_cap, int32_t vcap, int32_t tame, double w,
                                       int32_t nqp,
                                       double *csr_data, double *scratch, void *stream) {
  using namespace skb;
  if (ntiles < 0 || !p || nqp <= 0 || vcap <= 0 || rec_cap <= 0 || (rec_cap & 15)) return SKB_EINVAL;
  if (ntiles == 0) return SKB_OK;
  P1Args a;
  a.p = p; a.npts = npts; a.rec = (const unsigned char *)rec; a.rec_start = rec_start;
  a.ntiles = ntiles; a.rec_cap = rec_cap; a.vcap = vcap;
  a.csr_data = csr_data; a.scratch = scratch; a.w = w; a.nqp = nqp; a.tame = tame;
  a.debug = debug_flags();
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t smem = (size_t)skb_p1_fused_smem_bytes(tile_elems, ring, rec_cap, vcap);
  if (smem > 227 * 1024) return SKB_ETOOBIG;
  const bool q4 = (nqp == 4);
  int rc = SKB_EINVAL;
#define SKB_P1_CASE(TT, NRED)                                                  \
  if (tile_elems == TT && reduce_threads == NRED) {                            \
    if (ring == 4) rc = launch_fused<TT, NRED, 4>(a, smem, sms, q4, st);       \
    else if (ring == 5) rc = launch_fused<TT, NRED, 5>(a, smem, sms, q4, st);  \
  }
  SKB_P1_CASE(128, 96)
  SKB_P1_CASE(256, 128)
  SKB_P1_CASE(256, 224)
  SKB_P1_CASE(256, 256)
  SKB_P1_CASE(384, 224)
  SKB_P1_CASE(384, 352)
  SKB_P1_CASE(512, 128)
  SKB_P1_CASE(512, 256)
  SKB_P1_CASE(512, 384)
  SKB_P1_CASE(512, 480)
  SKB_P1_CASE(768, 224)
#undef SKB_P1_CASE
  // two elements per compute thread (256 compute threads), more reduce warps
#define SKB_P1_CASE2(TT, NRED, CT)                                                 \
  if (tile_elems == TT && reduce_threads == NRED) {                                \
    if (ring == 4) rc = launch_fused<TT, NRED, 4, CT>(a, smem, sms, q4, st);       \
    else if (ring == 5) rc = launch_fused<TT, NRED, 5, CT>(a, smem, sms, q4, st);  \
  }
  SKB_P1_CASE2(512, 736, 256)
  SKB_P1_CASE2(512, 608, 256)
  SKB_P1_CASE2(512, 640, 128)
#undef SKB_P1_CASE2
  if (rc == SKB_OK) count_launch();
  return rc;
}

// L2 residency for the per-tile partial sums: `scratch` is written by the fused kernel and
// read once by p1_combine right after it; pinning that window in the 126 MB L2 keeps the
// round trip out of HBM.  bytes == 0 resets the stream to the default policy.
extern "C" int skb_l2_window(const void *ptr, int64_t bytes, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  cudaStreamAttrValue v;
  memset(&v, 0, sizeof(v));
  if (ptr && bytes > 0) {
    int dev = 0, max_persist = 0, max_window = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    if (max_persist <= 0 || max_window <= 0) return SKB_OK;   // feature absent: nothing to do
    static int64_t limit_set[64] = {0};                 // per device (the limit is per device)
    const int64_t want = bytes < max_persist ? bytes : max_persist;
    const int slot = dev >= 0 && dev < 64 ? dev : 0;
    if (want > limit_set[slot] || dev >= 64) {
      SKB_CUDA_TRY(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)want));
      limit_set[slot] = want;
    }
    v.accessPolicyWindow.base_ptr = const_cast<void *>(ptr);
    v.accessPolicyWindow.num_bytes = (size_t)(bytes < max_window ? bytes : max_window);
    v.accessPolicyWindow.hitRatio = want >= bytes ? 1.0f : (float)want / (float)bytes;
    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  } else {
    v.accessPolicyWindow.num_bytes = 0;
    v.accessPolicyWindow.hitRatio = 0.0f;
    v.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
    v.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
  }
  SKB_CUDA_TRY(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v));
  return SKB_OK;
}

extern "C" int skb_p1_combine(const double *scratch, const uint32_t *sptr, const uint32_t *gslot,
                              const uint32_t *gslot2, int64_t nshared, double *csr_data,
                              void *stream) {
  using namespace skb;
  if (nshared < 0) return SKB_EINVAL;
  if (nshared == 0) return SKB_OK;
  int64_t g = (nshared + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  p1_combine_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(scratch, sptr, gslot, gslot2,
                                                              nshared, csr_data);
  count_launch();
  return (int)cudaGetLastError();
}

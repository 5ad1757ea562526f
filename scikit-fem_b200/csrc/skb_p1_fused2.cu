// Fused P1 (ElementTetP1) Laplace assembly, second generation: geometry -> local
// matrix -> CSR values in one pass; element-local matrices never touch HBM.
//
// Replaces, for the headline path, the whole chain
//   CellBasis.__init__           assembly/basis/cell_basis.py:94-106
//   BilinearForm._assemble       assembly/form/bilinear_form.py:58-128,150-151
//   COOData._assemble_scipy_csr  assembly/form/coo_data.py:27-36 (values)
// for form = models/poisson.py:7-9 (laplace) on ElementTetP1.
//
// Plan (skfem_b200/fused2.py).  Elements are ordered by a balanced k-d tree into
// *super-tiles* (compact boxes of S tiles) of *tiles* (T elements).  Every tile owns
// one contiguous 16-byte aligned record in HBM (one TMA bulk copy):
//     header | tl: T x ushort4 tile-local vertex ids (+ the expected zero mask of the
//     10 unique local entries in the spare bits) | verts: global vertex ids |
//     grp: per 32-lane group {offset, length} | lane: per lane the pool index of its
//     slot | ids: sliced-ELL staging indices, two per 32-bit word
// and every super-tile a flush table {CSR slot | scratch position, mirror slot} per
// accumulator of its *pool*.
//
// Kernel: persistent CTAs, several per SM, each walking whole super-tiles:
//   TMA    records are prefetched NR-1 tiles ahead (cp.async.bulk + mbarrier);
//   LDGSTS the vertex coordinates of tile k+1 are gathered (cp.async) while tile k
//          is reduced;
//   P1     one element per thread: the 10 unique local entries in registers, bit
//          identical to numpy, staged  vals[k*T + e]; the zero
//          mask of the entries is compared with the plan's (pattern validation);
//   P2     one lane per tile slot adds that slot's staged contributions in a fixed
//          order and accumulates into the super-tile's pool in shared memory;
//   flush  after the last tile of a super-tile the pool is written out in CSR order
//          (coalesced runs): slots complete inside the super-tile go to csr_data (and
//          their mirror), the others to a scratch array that skb_p1_combine adds in
//          super-tile order.
// No float atomics: results are bit-reproducible run to run.
#include <cstdio>
#include <cstring>
#include "skb_common.cuh"

namespace skb {

struct P1v2Args {
  const double *p;
  int64_t npts;
  const unsigned char *rec;     // concatenated tile records
  const uint64_t *rec_start;    // [ntiles+1] byte offsets (multiples of 16)
  const int64_t *st_fl0;        // [nst+1] first flush entry of every super-tile (even)
  const uint2 *fl;              // flush table: {target, mirror target}
  int32_t nst, ntiles, S;       // super-tiles, tiles, tiles per super-tile
  int32_t rec_cap;              // largest record, bytes (multiple of 16)
  int32_t vcap;                 // most vertices in one tile (even)
  int32_t pool_cap;             // most accumulators in one super-tile (even)
  int32_t ring;                 // record buffers
  double *csr_data;
  double *scratch;
  double w;                     // the common quadrature weight
  int32_t nqp;
  int32_t debug;
  int32_t *flag;                // device int: bit0 set when the zero mask of an element changed
  uint16_t *nz_out;             // plan time only: receives the zero mask of every element
  // MODE 4 (mass): phi_j(x_q) * phi_i(x_q) for the 10 unique pairs (row-major upper triangle)
  // at the 4 quadrature points, and the weights
  double cq[10][4];
  double wq[4];
};

struct RecHeader2 {             // 32 bytes at the start of every record
  uint32_t nverts, ngroups, off_verts, off_grp, off_lane, off_ids, nelems, pad;
};

// ---- async-copy / mbarrier primitives (PTX) ---------------------------------
__device__ __forceinline__ uint32_t smem_u32_2(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init2(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32_2(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx2(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32_2(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait2(uint64_t *bar, unsigned parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32_2(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait2(uint64_t *bar, unsigned parity) {
  while (!mbar_try_wait2(bar, parity)) {
  }
}
// TMA bulk copy global -> shared, completion counted on an mbarrier; evict-first in L2 (the
// records and flush tables are read once per step)
__device__ __forceinline__ void tma_bulk_g2s2(void *dst, const void *src, unsigned bytes,
                                              uint64_t *bar) {
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32_2(dst)), "l"(src), "r"(bytes), "r"(smem_u32_2(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void cp_async8_2(uint32_t dst_smem, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_smem), "l"(src) : "memory");
}

__device__ __forceinline__ unsigned nonzero_bits(double v) {   // v != +-0, integer pipe only
  return (((unsigned)__double2hiint(v) & 0x7fffffffu) | (unsigned)__double2loint(v)) != 0u;
}

// MODE 0: any mesh, any equal-weight rule: IEEE division, numpy's quadrature sum evaluated
//         term by term.
// MODE 1: nqp == 4, coordinates within the exact_div-safe range: one reciprocal + Markstein
//         corrections per quotient (skb_common.cuh exact_div).
// MODE 2: additionally coordinates within [2^-28, 2^28] and the weight within [2^-20, 1]: then
//         d * dx is a normal number or exactly 0 for every entry, so numpy's
//         ((v + v) + v) + v  (v = d * dx) equals 4 v = d * (4 dx) bit for bit - one
//         multiplication instead of three operations (proof in DESIGN.md).
// MODE 3: opt-in fast arithmetic (FMA + one reciprocal; values within a few ulp per term).
// MODE 4: the mass form u * v (models/poisson.py:17-19) instead of the Laplace form: no inverse,
//         entry = numpy's sum over the 4 points of (phi_j phi_i) * (|det| W_q).
//
// Roles: threads [0, T) compute (one element each in P1, the tile's slot groups in P2, the
// flush), one more warp whose lane 0 is the producer: it waits for records / flush tables and
// issues the TMA copies, so the latency of its global loads never delays a compute warp.
// The CTA processes the super-tiles blockIdx.x, blockIdx.x + gridDim.x, ...; its j-th tile
// uses record buffer j % NR.
// NT compute threads handle T / NT elements each (NT == T: one element per thread).
template <int T, int MODE, int NT = T, bool PROF = false>
__global__ void __launch_bounds__(NT + 32, (T >= 512 ? 2 : (T >= 256 ? 3 : 6)))
p1tet_laplace_fused2_kernel(const P1v2Args a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int NVALS = 10 * T + 16;     // + one staged 0.0 per bank pair
  double *vals = reinterpret_cast<double *>(smem_raw);             // [NVALS]
  double *coords = vals + NVALS;                                   // [2][vcap][3] (AoS)
  double *pool = coords + 6 * (size_t)a.vcap;                      // [pool_cap]
  uint2 *flbuf = reinterpret_cast<uint2 *>(pool + a.pool_cap);     // [pool_cap] flush table
  unsigned char *recs = reinterpret_cast<unsigned char *>(flbuf + a.pool_cap);
  const int NR = a.ring;
  uint64_t *mbar = reinterpret_cast<uint64_t *>(recs + (size_t)NR * a.rec_cap);   // [NR + 1]
  int *fl_np = reinterpret_cast<int *>(mbar + NR + 1);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = NT / 32;
  static_assert(T % NT == 0, "elements per compute thread must be integral");
  const bool is_compute = tid < NT;
  const bool is_producer = tid == NT;
  const int S = a.S, G = (int)gridDim.x;

  // this CTA's tile sequence
  int st = (int)blockIdx.x;
  if (st >= a.nst) return;
  const int nst_mine = (a.nst - st + G - 1) / G;
  const bool owns_last = ((a.nst - 1 - st) % G) == 0;
  const int nmine = nst_mine * S - (owns_last ? a.nst * S - a.ntiles : 0);
  int tile = st * S, tend = min(tile + S, a.ntiles);

  // producer state: issue cursor, wait cursor
  int p_st = st, p_tile = tile, p_tend = tend, p_slot = 0, p_issued = 0;
  int w_slot = 2 % NR;
  unsigned w_par = 0, fl_par = 0;
  auto issue_record = [&]() {         // producer: fetch the record of the next tile in sequence
    if (p_issued < nmine) {
      const uint64_t b0 = a.rec_start[p_tile];
      const unsigned bytes = (unsigned)(a.rec_start[p_tile + 1] - b0);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx2(&mbar[p_slot], bytes);
      tma_bulk_g2s2(recs + (size_t)p_slot * a.rec_cap, a.rec + b0, bytes, &mbar[p_slot]);
      if (++p_tile >= p_tend) {
        p_st += G;
        p_tile = p_st * S;
        p_tend = min(p_tile + S, a.ntiles);
      }
    }
    ++p_issued;
    if (++p_slot == NR) p_slot = 0;
  };
  auto issue_flush_table = [&](int s) {   // producer: fetch the flush table of super-tile s
    const int64_t f0 = a.st_fl0[s];
    const int np = (int)(a.st_fl0[s + 1] - f0);          // even
    *fl_np = np;
    if (np > 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx2(&mbar[NR], (unsigned)np * 8u);
      tma_bulk_g2s2(flbuf, a.fl + f0, (unsigned)np * 8u, &mbar[NR]);
    }
  };
  auto gather = [&](const unsigned char *r, int par) {   // async gather of a tile's vertices
    const RecHeader2 *h = reinterpret_cast<const RecHeader2 *>(r);
    const int nv = (int)h->nverts;
    const int32_t *verts = reinterpret_cast<const int32_t *>(r + h->off_verts);
    const uint32_t dst = smem_u32_2(coords + (size_t)par * 3 * a.vcap);
    const double *px = a.p, *py = a.p + a.npts, *pz = a.p + 2 * a.npts;
#pragma unroll 1
    for (int i = tid; i < nv; i += NT) {
      const int32_t gv = verts[i];
      cp_async8_2(dst + 24 * i, px + gv);
      cp_async8_2(dst + 24 * i + 8, py + gv);
      cp_async8_2(dst + 24 * i + 16, pz + gv);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  if (tid == 0) {
    for (int i = 0; i <= NR; ++i) mbar_init2(&mbar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int i = 0; i < 16; ++i) vals[10 * T + i] = 0.0;
  }
  __syncthreads();
  if (is_producer) {                    // prologue: NR records and the first flush table in flight
    for (int i = 0; i < NR; ++i) issue_record();
    issue_flush_table(st);
  }
  mbar_wait2(&mbar[0], 0);
  if (nmine > 1) mbar_wait2(&mbar[1], 0);
  if (is_compute) gather(recs, 0);
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();

  const double w1 = a.w;
  const double w4 = a.w * 4.0;
  unsigned bad = 0;
  // PROF (skb_debug_flags bit 2, separate instantiation: the counters cost registers): lane 0
  // of every warp of CTA 0 accumulates the cycles it spends per phase.  BAR.SYNC defers its
  // blocking to the next memory instruction, hence the dummy shared-memory load.
  const bool prof = PROF && blockIdx.x == 0 && lane == 0;
  long long pc[PROF ? 6 : 1] = {0}, t0 = 0;
#define SKB_TICK(i)                                  \
  if (PROF && prof) {                                \
    (void)*reinterpret_cast<volatile int *>(fl_np);  \
    const long long t1_ = clock64();                 \
    pc[PROF ? (i) : 0] += t1_ - t0;                  \
    t0 = t1_;                                        \
  }
  if (PROF && prof) t0 = clock64();
  int slot = 0, par = 0;
  bool first_of_st = false;             // the first super-tile's flush table is already in flight
  // Invariant at the top of iteration `it`: the records of this tile and of the next one have
  // landed and are visible to every thread; coords[par] holds this tile's vertices.
#pragma unroll 1
  for (int it = 0; it < nmine; ++it) {
    const int slot1 = slot + 1 == NR ? 0 : slot + 1;
    const unsigned char *r = recs + (size_t)slot * a.rec_cap;
    const RecHeader2 *h = reinterpret_cast<const RecHeader2 *>(r);
    const bool has_next = it + 1 < nmine;
    const bool last_of_st = tile + 1 >= tend;
    // the next tile's vertex coordinates travel while this tile is computed and reduced
    if (is_compute && has_next) gather(recs + (size_t)slot1 * a.rec_cap, par ^ 1);
    SKB_TICK(0)
    // ---- P1: local matrix of element `tid` -> vals -------------------------------------
    if (is_compute && !(a.debug & 1)) {
#pragma unroll 1
     for (int el = tid; el < T; el += NT) {
      const ushort4 v = reinterpret_cast<const ushort4 *>(r + sizeof(RecHeader2))[el];
      if (v.x != 0xFFFF) {   // not a padding element of a short tile
        const unsigned keep = (unsigned)(v.x >> 10) | ((unsigned)(v.y >> 10) << 6);
        const double *cb = coords + (par ? 3 * a.vcap : 0);
        const double *c0 = cb + 3u * (v.x & 0x3ffu), *c1 = cb + 3u * (v.y & 0x3ffu),
                     *c2 = cb + 3u * (v.z & 0x3ffu), *c3 = cb + 3u * (v.w & 0x3ffu);
        double A[3][3];
        {
          const double x0 = c0[0], y0 = c0[1], z0 = c0[2];
          A[0][0] = c1[0] - x0; A[0][1] = c2[0] - x0; A[0][2] = c3[0] - x0;
          A[1][0] = c1[1] - y0; A[1][1] = c2[1] - y0; A[1][2] = c3[1] - y0;
          A[2][0] = c1[2] - z0; A[2][1] = c2[2] - z0; A[2][2] = c3[2] - z0;
        }
        if (MODE == 4) {
          // |det A| only (mapping_affine.py:92-98); phi_j * phi_i is element independent
          const double m0 = A[1][1] * A[2][2] - A[1][2] * A[2][1];
          const double m1 = A[1][0] * A[2][2] - A[1][2] * A[2][0];
          const double m2 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
          const double adet = fabs(A[0][0] * m0 - A[0][1] * m1 + A[0][2] * m2);
          double dxq[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) dxq[q] = adet * a.wq[q];       // cell_basis.py:104-105
          unsigned nz = 0;
          double *out = vals + el;
#pragma unroll
          for (int k = 0; k < 10; ++k) {
            double val = 0.0;                                        // np.sum, n < 8: left to right
#pragma unroll
            for (int q = 0; q < 4; ++q) val = val + a.cq[k][q] * dxq[q];
            nz |= nonzero_bits(val) << k;
            out[k * T] = val;
          }
          bad |= (nz ^ keep);
          if (a.nz_out) a.nz_out[(size_t)tile * T + el] = (uint16_t)nz;
        } else {
        double det, n[3][3], inv[3][3];
        if (MODE == 3) {
#pragma unroll
          for (int i = 0; i < 3; ++i) {   // cofactors, same sign convention as cofactors3
            const int i1_ = (i + 1) % 3, i2_ = (i + 2) % 3;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
              n[j][i] = __fma_rn(A[i1_][j1], A[i2_][j2], -(A[i1_][j2] * A[i2_][j1]));
            }
          }
          det = __fma_rn(A[0][0], n[0][0], __fma_rn(A[0][1], n[1][0], A[0][2] * n[2][0]));
          const double y = 1.0 / det;
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) inv[i][j] = n[i][j] * y;
        } else {
          // minors of the first row are shared between det (mapping_affine.py:92-98) and the
          // first column of the inverse (:111-129): -b + a == a - b and b - a == -(a - b)
          // bit for bit
          const double m0 = A[1][1] * A[2][2] - A[1][2] * A[2][1];
          const double m1 = A[1][0] * A[2][2] - A[1][2] * A[2][0];
          const double m2 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
          det = A[0][0] * m0 - A[0][1] * m1 + A[0][2] * m2;
          n[0][0] = m0;
          n[1][0] = -m1;
          n[2][0] = m2;
          n[0][1] = A[0][2] * A[2][1] - A[0][1] * A[2][2];
          n[1][1] = -A[0][2] * A[2][0] + A[0][0] * A[2][2];
          n[2][1] = A[0][1] * A[2][0] - A[0][0] * A[2][1];
          n[0][2] = -A[0][2] * A[1][1] + A[0][1] * A[1][2];
          n[1][2] = A[0][2] * A[1][0] - A[0][0] * A[1][2];
          n[2][2] = -A[0][1] * A[1][0] + A[0][0] * A[1][1];
          if (MODE >= 1 && det != 0.0) {
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
        }
        // P1 push-forward: dphi_b is +-unit, so grad_b (b=1..3) is row b-1 of inv and
        // grad_0[j] = -((inv0j + inv1j) + inv2j)   (Appendix A.4)
        double g[4][3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          g[0][j] = -((inv[0][j] + inv[1][j]) + inv[2][j]);
          g[1][j] = inv[0][j];
          g[2][j] = inv[1][j];
          g[3][j] = inv[2][j];
        }
        const double adet = fabs(det);
        const double dx = adet * w1;            // cell_basis.py:104-105
        const double dx4 = adet * w4;           // == 4 * dx exactly (power-of-two scaling)
        unsigned nz = 0;
        double *out = vals + el;
        int k = 0;
#pragma unroll
        for (int pp = 0; pp < 4; ++pp)
#pragma unroll
          for (int q = pp; q < 4; ++q, ++k) {
            double val;
            if (MODE == 3) {
              const double d = __fma_rn(g[pp][2], g[q][2],
                                        __fma_rn(g[pp][1], g[q][1], g[pp][0] * g[q][0]));
              val = d * dx4;
            } else {
              const double d = (g[pp][0] * g[q][0] + g[pp][1] * g[q][1]) + g[pp][2] * g[q][2];
              if (MODE == 2) {
                val = d * dx4;
              } else if (MODE == 1) {
                const double t = d * dx;
                val = __fma_rn(t, 2.0, t) + t;   // (t+t)+t in one rounding, then + t
              } else {
                const double t = d * dx;
                auto f = [&](int) -> double { return t; };
                val = pw_sum(a.nqp, f);
              }
            }
            nz |= nonzero_bits(val) << k;
            out[k * T] = val;
          }
        bad |= (nz ^ keep);
        if (a.nz_out) a.nz_out[(size_t)tile * T + el] = (uint16_t)nz;
        }
      }
     }
    }
    SKB_TICK(1)
    __syncthreads();   // (A) vals complete
    SKB_TICK(2)
    if (is_producer) {
      // the record two tiles ahead must be visible at the top of the next iteration (its
      // vertex list feeds the gather); it has been in flight for NR - 2 iterations
      if (it + 2 < nmine) mbar_wait2(&mbar[w_slot], w_par);
      if (++w_slot == NR) { w_slot = 0; w_par ^= 1u; }
      // every compute thread has left the previous super-tile's flush: its table buffer is free
      if (first_of_st) issue_flush_table(st);
      if (last_of_st && *fl_np > 0) { mbar_wait2(&mbar[NR], fl_par); fl_par ^= 1u; }
    }
    // ---- P2: per-slot sums in fixed order (sliced ELL), accumulated into the pool ---------
    if (is_compute && !(a.debug & 2)) {
      const int ngroups = (int)h->ngroups;
      const uint32_t *grp = reinterpret_cast<const uint32_t *>(r + h->off_grp);
      const uint16_t *lanew = reinterpret_cast<const uint16_t *>(r + h->off_lane) + lane;
      const unsigned char *ids = r + h->off_ids + 4 * lane;
      const unsigned char *vb = reinterpret_cast<const unsigned char *>(vals);
      // the id words hold two byte offsets into vals; a lane's even and odd columns are summed
      // separately (s0, s1), rows 2 apart alternate between two more accumulators
#define SKB_V(off) (*reinterpret_cast<const double *>(vb + (off)))
#define SKB_W(row) (*reinterpret_cast<const uint32_t *>(cb + 128u * (row)))
#pragma unroll 1
      for (int gi = warp; gi < ngroups; gi += NW) {
        const uint32_t gw = grp[gi];
        int rows = (int)((gw >> 16) & 0x7fffu);            // two ELL columns per row, >= 1
        const unsigned char *cb = ids + ((gw & 0xffffu) << 7);
        const unsigned lw = lanew[gi * 32];
        double acc;
        if (rows == 1) {
          const uint32_t w0 = SKB_W(0);
          acc = SKB_V(w0 & 0xffffu) + SKB_V(w0 >> 16);
        } else if (rows == 2) {
          const uint32_t w0 = SKB_W(0), w1 = SKB_W(1);
          const double a0 = SKB_V(w0 & 0xffffu), a1 = SKB_V(w0 >> 16);
          const double b0 = SKB_V(w1 & 0xffffu), b1 = SKB_V(w1 >> 16);
          acc = (a0 + b0) + (a1 + b1);
        } else {
          double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll 1
          for (; rows >= 2; rows -= 2, cb += 256) {
            const uint32_t w0 = SKB_W(0), w1 = SKB_W(1);
            const double a0 = SKB_V(w0 & 0xffffu), a1 = SKB_V(w0 >> 16);
            const double b0 = SKB_V(w1 & 0xffffu), b1 = SKB_V(w1 >> 16);
            s0 = s0 + a0; s1 = s1 + a1; s2 = s2 + b0; s3 = s3 + b1;
          }
          if (rows) {
            const uint32_t w0 = SKB_W(0);
            s0 = s0 + SKB_V(w0 & 0xffffu);
            s1 = s1 + SKB_V(w0 >> 16);
          }
          acc = (s0 + s2) + (s1 + s3);
        }
        if (gw & 0x80000000u) {
          // long lists are split over 2 or 4 adjacent lanes: fixed combination tree
          // (l + l+1) + (l+2 + l+3), selected by the leader lane
          const double t1 = acc + __shfl_down_sync(0xffffffffu, acc, 1);
          const double t2 = t1 + __shfl_down_sync(0xffffffffu, t1, 2);
          const unsigned fs = (lw >> 13) & 3u;
          acc = fs == 0 ? acc : (fs == 1 ? t1 : t2);
        }
        if (lw != 0xFFFFu) {
          double *pa = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(pool) +
                                                  ((lw & 0x1fffu) << 3));
          if (!(lw & 0x8000u)) acc = *pa + acc;            // not the first tile touching it
          *pa = acc;
        }
      }
#undef SKB_V
#undef SKB_W
    }
    SKB_TICK(3)
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();   // (B) vals, record `slot` free; coords of the next tile visible; pool updated
    SKB_TICK(4)
    if (is_producer) issue_record();
    // ---- flush: the super-tile's pool -> csr_data / scratch, in CSR order --------------------
    if (last_of_st && is_compute && !(a.debug & 64)) {
      const int np = *fl_np;
#pragma unroll 2
      for (int i = tid; i < np; i += NT) {
        const uint2 m = flbuf[i];
        const double val = pool[i];
        if (m.x != 0xffffffffu) {
          if (m.x & 0x80000000u) a.scratch[m.x & 0x7fffffffu] = val;
          else a.csr_data[m.x] = val;
          if (m.y != 0xffffffffu) a.csr_data[m.y] = val;
        }
      }
      // the next super-tile's first pool write comes after barrier (A) of its first tile
    }
    SKB_TICK(5)
    // next tile of the sequence
    first_of_st = last_of_st;
    if (last_of_st) {
      st += G;
      tile = st * S;
      tend = min(tile + S, a.ntiles);
    } else {
      ++tile;
    }
    slot = slot1;
    par ^= 1;
  }
  if (bad & 0x3ffu) atomicOr(a.flag, 1);
  if (PROF && prof)
    printf("fused2 cta0 warp %d: tiles %d  cycles/tile: gather issue %lld  P1 %lld  "
           "barrier A %lld  P2 %lld  cp.async wait + barrier B %lld  flush / record issue %lld\n",
           warp, nmine, pc[0] / nmine, pc[PROF ? 1 : 0] / nmine, pc[PROF ? 2 : 0] / nmine,
           pc[PROF ? 3 : 0] / nmine, pc[PROF ? 4 : 0] / nmine, pc[PROF ? 5 : 0] / nmine);
#undef SKB_TICK
}

__global__ void __launch_bounds__(256)
p1_combine2_kernel(const double *__restrict__ scratch, const uint32_t *__restrict__ sptr,
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

template <int T, int MODE, int NT = T, bool PROF = false>
static int launch_fused2(const P1v2Args &a, size_t smem, int sms, int ctas_per_sm,
                         cudaStream_t st) {
  auto k = p1tet_laplace_fused2_kernel<T, MODE, NT, PROF>;
  SKB_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  SKB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, NT + 32, smem));
  if (occ < 1) return SKB_ETOOBIG;
  if (ctas_per_sm > 0 && occ > ctas_per_sm) occ = ctas_per_sm;
  int free_sms = sm_reserve();
  if (free_sms > sms - 1) free_sms = sms - 1;
  const int cap = occ * (sms - free_sms);
  const int grid = a.nst < cap ? a.nst : cap;
  k<<<grid, NT + 32, smem, st>>>(a);
  return (int)cudaGetLastError();
}

}  // namespace skb

extern "C" int64_t skb_p1_fused2_smem_bytes(int32_t tile_elems, int32_t ring, int32_t rec_cap,
                                            int32_t vcap, int32_t pool_cap) {
  return (int64_t)(sizeof(double) * ((10 * (size_t)tile_elems + 16) + 6 * (size_t)vcap +
                                     2 * (size_t)pool_cap) +
                   (size_t)ring * rec_cap + 8 * ((size_t)ring + 1) + 32);
}

// Warm fused P1-tet Laplace assembly (see the top of this file).  mode: 0 generic, 1 exact
// division via reciprocal (tame coordinates, nqp == 4), 2 additionally the 4 * dx shortcut,
// 3 fast (FMA) arithmetic.  `flag` (device int32) gets bit 0 set if the zero mask of any
// element's local matrix differs from the plan's: the cached CSR pattern is then no longer the
// reference's pattern and the caller must re-plan (coo_data.py:35, eliminate_zeros).
// nz_out != NULL (plan time): only the local matrices are formed and the 10-bit zero mask of
// element e of tile t is stored at nz_out[t * tile_elems + e]; nothing else is written.
static int p1tet_fused2_launch(const double *tab_host, const double *p, int64_t npts, const void *rec,
                                        const uint64_t *rec_start, const int64_t *st_fl0,
                                        const void *fl, int32_t nst, int32_t ntiles,
                                        int32_t tiles_per_super, int32_t tile_elems,
                                        int32_t ring, int32_t rec_cap,
                                        int32_t vcap, int32_t pool_cap, int32_t ctas_per_sm,
                                        int32_t mode, double w, int32_t nqp, double *csr_data,
                                        double *scratch, int32_t *flag, uint16_t *nz_out,
                                        void *stream) {
  using namespace skb;
  if (nst < 0 || !p || nqp <= 0 || vcap <= 0 || (vcap & 1) || pool_cap <= 0 || (pool_cap & 1) ||
      rec_cap <= 0 || (rec_cap & 15) || ring < 3 || ring > 8 || !flag || tiles_per_super < 1 ||
      ntiles > (int64_t)nst * tiles_per_super || ntiles <= (int64_t)(nst - 1) * tiles_per_super)
    return SKB_EINVAL;
  if (mode < 0 || mode > 4 || ((mode == 1 || mode == 2 || mode == 4) && nqp != 4) ||
      (mode == 4 && !tab_host))
    return SKB_EINVAL;
  if (nst == 0) return SKB_OK;
  P1v2Args a;
  a.p = p; a.npts = npts; a.rec = (const unsigned char *)rec; a.rec_start = rec_start;
  a.st_fl0 = st_fl0; a.fl = (const uint2 *)fl; a.nst = nst; a.ntiles = ntiles;
  a.S = tiles_per_super;
  a.rec_cap = rec_cap; a.vcap = vcap; a.pool_cap = pool_cap; a.ring = ring;
  a.csr_data = csr_data; a.scratch = scratch; a.w = w; a.nqp = nqp;
  a.debug = debug_flags(); a.flag = flag; a.nz_out = nz_out;
  for (int k = 0; k < 10; ++k)
    for (int q = 0; q < 4; ++q) a.cq[k][q] = 0.0;
  for (int q = 0; q < 4; ++q) a.wq[q] = 0.0;
  if (mode == 4) {
    // phi (4 basis functions x 4 points), then W (4): the products phi_j * phi_i are rounded
    // here exactly as numpy rounds u * v (bilinear_form.py:151 multiplies by dx afterwards)
    int k = 0;
    for (int i = 0; i < 4; ++i)
      for (int j = i; j < 4; ++j, ++k)
        for (int q = 0; q < 4; ++q) a.cq[k][q] = tab_host[j * 4 + q] * tab_host[i * 4 + q];
    for (int q = 0; q < 4; ++q) a.wq[q] = tab_host[16 + q];
  }
  if (nz_out) a.debug |= 2 | 64;     // plan-time mask pass: local matrices only
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t smem = (size_t)skb_p1_fused2_smem_bytes(tile_elems, ring, rec_cap, vcap, pool_cap);
  if (smem > 227 * 1024) return SKB_ETOOBIG;
  int rc = SKB_EINVAL;
#define SKB_P1V2_CASE(TT)                                                                \
  if (tile_elems == TT) {                                                                \
    if (mode == 0) rc = launch_fused2<TT, 0>(a, smem, sms, ctas_per_sm, st);             \
    else if (mode == 1) rc = launch_fused2<TT, 1>(a, smem, sms, ctas_per_sm, st);        \
    else if (mode == 2) rc = launch_fused2<TT, 2>(a, smem, sms, ctas_per_sm, st);        \
    else if (mode == 3) rc = launch_fused2<TT, 3>(a, smem, sms, ctas_per_sm, st);        \
    else rc = launch_fused2<TT, 4>(a, smem, sms, ctas_per_sm, st);                       \
  }
  const int ept = ctas_per_sm >> 8;        // bits 8..: elements per compute thread (0/1 = one)
  ctas_per_sm &= 0xff;
  if ((a.debug & 4) && tile_elems == 256 && mode == 2 && ept <= 1) {
    rc = launch_fused2<256, 2, 256, true>(a, smem, sms, ctas_per_sm, st);   // phase counters
  } else if (ept <= 1) {
    SKB_P1V2_CASE(128)
    SKB_P1V2_CASE(256)
    SKB_P1V2_CASE(512)
  } else if (ept == 2 && tile_elems == 512) {
    if (mode == 0) rc = launch_fused2<512, 0, 256>(a, smem, sms, ctas_per_sm, st);
    else if (mode == 1) rc = launch_fused2<512, 1, 256>(a, smem, sms, ctas_per_sm, st);
    else if (mode == 2) rc = launch_fused2<512, 2, 256>(a, smem, sms, ctas_per_sm, st);
    else if (mode == 3) rc = launch_fused2<512, 3, 256>(a, smem, sms, ctas_per_sm, st);
  }
#undef SKB_P1V2_CASE
  if (rc == SKB_OK) count_launch();
  return rc;
}


extern "C" int skb_p1tet_laplace_fused2(const double *p, int64_t npts, const void *rec,
                                        const uint64_t *rec_start, const int64_t *st_fl0,
                                        const void *fl, int32_t nst, int32_t ntiles,
                                        int32_t tiles_per_super, int32_t tile_elems,
                                        int32_t ring, int32_t rec_cap,
                                        int32_t vcap, int32_t pool_cap, int32_t ctas_per_sm,
                                        int32_t mode, double w, int32_t nqp, double *csr_data,
                                        double *scratch, int32_t *flag, uint16_t *nz_out,
                                        void *stream) {
  if (mode == 4) return SKB_EINVAL;
  return p1tet_fused2_launch(nullptr, p, npts, rec, rec_start, st_fl0, fl, nst, ntiles,
                             tiles_per_super, tile_elems, ring, rec_cap, vcap, pool_cap,
                             ctas_per_sm, mode, w, nqp, csr_data, scratch, flag, nz_out, stream);
}

// The same pipeline for the mass form u * v on ElementTetP1 with the 4-point rule
// (models/poisson.py:17-19): tab_host = phi[4][4] (basis function x quadrature point) followed
// by W[4], host doubles.  The mass matrix has the full graph pattern; plan and records are built
// exactly as for the Laplace form.
extern "C" int skb_p1tet_mass_fused2(const double *tab_host, const double *p, int64_t npts,
                                     const void *rec, const uint64_t *rec_start,
                                     const int64_t *st_fl0, const void *fl, int32_t nst,
                                     int32_t ntiles, int32_t tiles_per_super, int32_t tile_elems,
                                     int32_t ring, int32_t rec_cap, int32_t vcap,
                                     int32_t pool_cap, int32_t ctas_per_sm, double *csr_data,
                                     double *scratch, int32_t *flag, uint16_t *nz_out,
                                     void *stream) {
  return p1tet_fused2_launch(tab_host, p, npts, rec, rec_start, st_fl0, fl, nst, ntiles,
                             tiles_per_super, tile_elems, ring, rec_cap, vcap, pool_cap,
                             ctas_per_sm, 4, 1.0, 4, csr_data, scratch, flag, nz_out, stream);
}

extern "C" int skb_p1_combine2(const double *scratch, const uint32_t *sptr, const uint32_t *gslot,
                               const uint32_t *gslot2, int64_t nshared, double *csr_data,
                               void *stream) {
  using namespace skb;
  if (nshared < 0) return SKB_EINVAL;
  if (nshared == 0) return SKB_OK;
  int64_t g = (nshared + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  p1_combine2_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(scratch, sptr, gslot, gslot2,
                                                               nshared, csr_data);
  count_launch();
  return (int)cudaGetLastError();
}

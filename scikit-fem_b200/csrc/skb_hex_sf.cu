// Element-local Laplace / mass matrices of ElementHex2 by sum factorisation.
//
// The reference evaluates  A_ij = sum_q dot(grad phi_j, grad phi_i) dx  over the
// 343 points of the default rule for all 27 x 27 pairs (bilinear_form.py:86-98 on
// element_hex2.py:1255-1260, mapping_isoparametric.py:112-226): 27*27*343*3
// multiply-adds per element, which is what the Gram-matrix kernel of
// csrc/skb_hex_mma.cu still performs (on the FP64 tensor cores).  Basis functions
// and rule are tensor products, so with G = |det J| W J^-1 J^-T (6 components per
// point)
//     A_ij = sum_{d,e} sum_{q1,q2,q3} G_de(q) u1(q1) u2(q2) u3(q3),
//     u_k = pp[type_k(d,e)][3 i_k + j_k]  (products of the 1-D functions / derivatives)
// and the sum is taken one axis at a time:
//     T1[c][i1 j1][q2 q3]     = sum_q1 G_c  * pp[type1(c)][i1 j1][q1]     9*49*7 per combo
//     T2[g][i1 j1][i2 j2][q3] = sum_q2 T1_c * pp[type2(c)][i2 j2][q2]     81*7*7, combos with the
//                                                                        same third-axis table summed
//     B[g'][i1 j1][i2 j2][i3 j3] = sum_q3 T2_g * pp[type3(g)][i3 j3][q3]  729*7 per group
// over the 6 combos d <= e (G is symmetric; the off-diagonal combos enter as B_ij + B_ji):
// 63 k multiply-adds per element instead of 750 k.  Hex2 parity is value-level (rtol
// 1e-12, tests/test_gpu_parity.py::test_hex2_value_parity; the formulation is checked
// against the reference on the CPU in tests/test_hex2_sumfact_cpu.py), so the
// quadrature sum may be re-associated; ElementHex1 (bit-exact) keeps the scalar kernel.
//
// Mapping onto the SM: E elements per CTA pass (2, two CTAs per SM, 88 kB of shared memory
// each).  Every stage is a flat list of work items (element, row) spread over the CTA; an item
// loads its 7 inputs of one combo from shared memory once and runs 9 x 7 FMAs whose second
// operand is a *kernel parameter at a compile-time offset* - the 1-D tables live in the
// constant bank (LDCU into uniform registers, DFMA R, R, UR, R), no shared-memory traffic for
// them.  All loops over combos, groups and table rows are unrolled.  Geometry: the 12 edge
// differences of the trilinear map are fetched one pass ahead (vertex numbers and coordinates
// at two different points of the pass, so neither latency is exposed); Jacobian column f
// depends on the two other axes only (3 x 49 values per element), G needs 343 points.
// The last stage adds the symmetric and the two off-diagonal parts and writes either the
// reference layout (27, 27, nel) or, for warm re-assembly, element-major (nel, 27, 27) -
// 729 consecutive doubles per element, which skb_csr_reduce_em gathers sector by sector.
// Measured (BASELINE configs[3], 262 144 elements, one B200): 4.1 ms against 18.2 ms for the
// Gram kernel; ncu stage split and the CTA shapes tried: profiles/r2_hex_sumfact.md.
#include "skb_common.cuh"

namespace skb {

constexpr int SF_NQ = 7;

struct HexSfTab {
  double pp[4][9][SF_NQ];   // [type][3 i + j][q]: bit 0 of type = derivative on i, bit 1 on j
  double g[2][SF_NQ];       // geometry element: 1 - x_q, x_q
  int32_t qs[3];            // stride of the 1-D index of axis k in the external point index
  uint8_t bnode[32];        // a + 3 b + 9 c of basis function k
  uint8_t vtx[8];           // local vertex at corner (a, b, c): vtx[4 a + 2 b + c]
};

// combos c = 0..5: (d, e) = (0,0) (1,1) (2,2) (0,1) (0,2) (1,2)
__host__ __device__ constexpr int sf_d(int c) { return c < 3 ? c : (c == 5 ? 1 : 0); }
__host__ __device__ constexpr int sf_e(int c) { return c < 3 ? c : (c == 3 ? 1 : 2); }
__host__ __device__ constexpr int sf_type(int c, int k) {
  return (sf_d(c) == k ? 1 : 0) + (sf_e(c) == k ? 2 : 0);
}
// groups of combos sharing the third-axis table and the symmetric / transposed-add
// treatment: 0 = {00, 11} (l l, sym), 1 = {22} (l' l', sym), 2 = {01} (l l, off),
// 3 = {02, 12} (l l', off)
__host__ __device__ constexpr int sf_grp(int c) { return c < 2 ? 0 : (c == 2 ? 1 : (c == 3 ? 2 : 3)); }
__host__ __device__ constexpr int sf_gtype3(int g) { return g == 0 ? 0 : (g == 1 ? 3 : (g == 2 ? 0 : 2)); }

template <int FORM> struct SfShape {
  static constexpr int NQ = SF_NQ, NQ2 = NQ * NQ, NQ3 = NQ2 * NQ;
  static constexpr int NC = FORM == SKB_FORM_LAPLACE ? 6 : 1;    // combos
  static constexpr int NG = FORM == SKB_FORM_LAPLACE ? 4 : 1;    // groups
  static constexpr int NB = FORM == SKB_FORM_LAPLACE ? 2 : 1;    // B arrays (sym, off)
  // row strides == NQ (mod 16) doubles: the items of stage 2 / 3 (consecutive rows, NQ
  // consecutive points each) then fall into distinct 8-byte banks
  static constexpr int LD1 = NQ2 + (((NQ - NQ2) % 16) + 16) % 16;          // 55
  static constexpr int LD2 = 9 * NQ + (((NQ - 9 * NQ) % 16) + 16) % 16;    // 71
  static constexpr int SZ_G = NC * NQ3, SZ_T2 = NG * 9 * LD2;
  static constexpr int SZ_JC = 9 * NQ2, SZ_T1 = NC * 9 * LD1, SZ_B = NB * 729;
  static constexpr int R0 = SZ_G > SZ_T2 ? SZ_G : SZ_T2;                   // G, then T2
  static constexpr int R1a = SZ_T1 > SZ_B ? SZ_T1 : SZ_B;
  static constexpr int R1 = R1a > SZ_JC ? R1a : SZ_JC;                     // J columns, T1, B
  static constexpr int PER = R0 + R1;
};
// per-element stride: == 16 / E (mod 16) doubles, so that the E elements a warp of the last
// stage reads side by side sit in different 8-byte banks
template <int FORM, int E> __host__ __device__ constexpr int sf_per() {
  constexpr int want = (16 / E) % 16, have = SfShape<FORM>::PER % 16;
  return SfShape<FORM>::PER + (want - have + 16) % 16;
}

template <int FORM, int E, int T, int MINB>
__global__ void __launch_bounds__(T, MINB)
local_hex_sf_kernel(const skb_space_t s, const __grid_constant__ HexSfTab tb, const int em,
                    double *__restrict__ out, int *__restrict__ err) {
  using S = SfShape<FORM>;
  constexpr int NQ = S::NQ, NQ2 = S::NQ2, NQ3 = S::NQ3, NC = S::NC, NG = S::NG;
  constexpr int LD1 = S::LD1, LD2 = S::LD2;
  constexpr bool LAP = FORM == SKB_FORM_LAPLACE;
  static_assert(E * 36 <= T && T % E == 0, "one edge difference per thread");
  extern __shared__ double sm[];
  __shared__ double sg[2][NQ];
  __shared__ double sd[E * 36];          // edge differences [el][f][2 a + b][i] of the batch
  __shared__ uint16_t sinv[730];         // 27 i + j of the B entry (i1 j1, i2 j2, i3 j3)
  const int tid = threadIdx.x;
  if (tid < 2 * NQ) sg[tid / NQ][tid % NQ] = tb.g[tid / NQ][tid % NQ];
  for (int ij = tid; ij < 729; ij += T) {
    const int bi = tb.bnode[ij / 27], bj = tb.bnode[ij % 27];
    const int i1 = bi % 3, i2 = (bi / 3) % 3, i3 = bi / 9, j1 = bj % 3, j2 = (bj / 3) % 3,
              j3 = bj / 9;
    sinv[((i1 * 3 + j1) * 9 + i2 * 3 + j2) * 9 + i3 * 3 + j3] = (uint16_t)ij;
  }
  constexpr int PER = sf_per<FORM, E>();
  auto r0 = [&](int el) { return sm + el * PER; };
  auto r1 = [&](int el) { return sm + el * PER + S::R0; };
  // thread tid < 36 E owns one component of one edge difference of the trilinear map:
  // x(f = 1, a, b) - x(f = 0, a, b), (a, b) the corner bits of the two other axes.  The two
  // dependent loads (vertex numbers, coordinates) of the next pass are issued at different
  // points of the current one, so that neither latency is waited for
  int vhi = 0, vlo = 0;
  auto load_verts = [&](int64_t b0) {
    if (tid >= E * 36 || b0 >= s.nel) return;
    const int el = tid / 36, r = tid % 36, f = r / 12, a = (r / 6) & 1, b = (r / 3) & 1;
    const int wf = 4 >> f, wu = f == 0 ? 2 : 4, wv = f == 2 ? 2 : 1;     // corner bit weights
    int64_t e = b0 + el;
    if (e >= s.nel) e = s.nel - 1;                        // slots past the end repeat the last
    const int64_t eg = s.tind ? (int64_t)s.tind[e] : e;
    vlo = s.t[(int64_t)tb.vtx[a * wu + b * wv] * s.nel_total + eg];
    vhi = s.t[(int64_t)tb.vtx[wf + a * wu + b * wv] * s.nel_total + eg];
  };
  auto load_diff = [&]() -> double {
    const double *pi = s.p + (int64_t)(tid % 3) * s.npts;   // component i = (tid % 36) % 3
    return tid < E * 36 ? __ldg(pi + vhi) - __ldg(pi + vlo) : 0.0;
  };
  load_verts((int64_t)blockIdx.x * E);
  double dnext = load_diff();

  for (int64_t base = (int64_t)blockIdx.x * E; base < s.nel; base += (int64_t)gridDim.x * E) {
    if (tid < E * 36) sd[tid] = dnext;
    __syncthreads();           // also: the previous pass' reads of B are over (J columns alias it)
    // ---- Jacobian columns: column f depends on the two other axes only -------------------
    // Jc[f][i][qu qv] = sum_{a,b} (x(f = 1, a, b) - x(f = 0, a, b))_i g_a(qu) g_b(qv)
    for (int item = tid; item < E * 3 * NQ2; item += T) {
      const int el = item / (3 * NQ2), r = item % (3 * NQ2), f = r / NQ2, uv = r % NQ2;
      const int qu = uv / NQ, qv = uv % NQ;
      const double *d = sd + el * 36 + f * 12;
      const double gu0 = sg[0][qu], gu1 = sg[1][qu], gv0 = sg[0][qv], gv1 = sg[1][qv];
      const double w00 = gu0 * gv0, w01 = gu0 * gv1, w10 = gu1 * gv0, w11 = gu1 * gv1;
      double *jc = r1(el);
#pragma unroll
      for (int i = 0; i < 3; ++i)
        jc[(f * 3 + i) * NQ2 + uv] =
            __fma_rn(d[9 + i], w11, __fma_rn(d[6 + i], w10, __fma_rn(d[3 + i], w01, d[i] * w00)));
    }
    __syncthreads();
    load_verts(base + (int64_t)gridDim.x * E);             // in flight during the G stage
    // ---- G at every point ----------------------------------------------------------------
#pragma unroll 2
    for (int item = tid; item < E * NQ3; item += T) {
      const int el = item / NQ3, q = item % NQ3, q1 = q / NQ2, q2 = (q / NQ) % NQ, q3 = q % NQ;
      const double *jc = r1(el);
      double J[3][3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        J[i][0] = jc[(0 + i) * NQ2 + q2 * NQ + q3];
        J[i][1] = jc[(3 + i) * NQ2 + q1 * NQ + q3];
        J[i][2] = jc[(6 + i) * NQ2 + q1 * NQ + q2];
      }
      const double det = det3(J);
      if (det == 0.0 && base + el < s.nel) atomicExch(err, 1);   // mapping_isoparametric.py:195-196
      const double w = __ldg(s.W + q1 * tb.qs[0] + q2 * tb.qs[1] + q3 * tb.qs[2]);
      double *G = r0(el);
      if (LAP) {
        double nn[3][3];                                // adjugate: nn / det = J^-1
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int r1_ = (r + 1) % 3, r2_ = (r + 2) % 3;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const int c1 = (c + 1) % 3, c2 = (c + 2) % 3;
            // cofactor of J[c][r] (cyclic form, no sign flips)
            nn[r][c] = __fma_rn(J[c1][r1_], J[c2][r2_], -(J[c1][r2_] * J[c2][r1_]));
          }
        }
        const double sc = w * __drcp_rn(fabs(det));
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int d = sf_d(c), e = sf_e(c);
          G[c * NQ3 + q] = __fma_rn(nn[d][2], nn[e][2],
                                    __fma_rn(nn[d][1], nn[e][1], nn[d][0] * nn[e][0])) * sc;
        }
      } else {
        G[q] = fabs(det) * w;
      }
    }
    dnext = load_diff();                                   // in flight during stages 1 - 3
    __syncthreads();
    // ---- stage 1: contract q1 ---------------------------------------------------------------
    for (int item = tid; item < E * NQ2; item += T) {
      const int el = item / NQ2, q23 = item % NQ2;
      const double *G = r0(el);
      double *T1 = r1(el);                              // the J columns are dead
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int ty = LAP ? sf_type(c, 0) : 0;
        double gq[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) gq[q] = G[c * NQ3 + q * NQ2 + q23];
#pragma unroll
        for (int ij = 0; ij < 9; ++ij) {
          double acc = gq[0] * tb.pp[ty][ij][0];
#pragma unroll
          for (int q = 1; q < NQ; ++q) acc = __fma_rn(gq[q], tb.pp[ty][ij][q], acc);
          T1[(c * 9 + ij) * LD1 + q23] = acc;
        }
      }
    }
    __syncthreads();
    // ---- stage 2: contract q2, combos of one group summed ---------------------------------
    for (int item = tid; item < E * 9 * NQ; item += T) {
      const int el = item / (9 * NQ), r = item % (9 * NQ), a = r / NQ, q3 = r % NQ;
      const double *T1 = r1(el);
      double *T2 = r0(el);                              // G is dead
      double acc[NG][9];
#pragma unroll
      for (int g = 0; g < NG; ++g)
#pragma unroll
        for (int b = 0; b < 9; ++b) acc[g][b] = 0.0;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int ty = LAP ? sf_type(c, 1) : 0, g = LAP ? sf_grp(c) : 0;
        double tq[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) tq[q] = T1[(c * 9 + a) * LD1 + q * NQ + q3];
#pragma unroll
        for (int b = 0; b < 9; ++b)
#pragma unroll
          for (int q = 0; q < NQ; ++q) acc[g][b] = __fma_rn(tq[q], tb.pp[ty][b][q], acc[g][b]);
      }
#pragma unroll
      for (int g = 0; g < NG; ++g)
#pragma unroll
        for (int b = 0; b < 9; ++b) T2[(g * 9 + a) * LD2 + b * NQ + q3] = acc[g][b];
    }
    __syncthreads();
    // ---- stage 3: contract q3 into the symmetric and the off-diagonal part -------------------
    for (int item = tid; item < E * 81; item += T) {
      const int el = item / 81, ab = item % 81, a = ab / 9, b = ab % 9;
      const double *T2 = r0(el);
      double *B = r1(el);                               // T1 is dead
      double acc[2][9];
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int k = 0; k < 9; ++k) acc[h][k] = 0.0;
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        const int ty = LAP ? sf_gtype3(g) : 0, h = g < 2 ? 0 : 1;
        double tq[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) tq[q] = T2[(g * 9 + a) * LD2 + b * NQ + q];
#pragma unroll
        for (int k = 0; k < 9; ++k)
#pragma unroll
          for (int q = 0; q < NQ; ++q) acc[h][k] = __fma_rn(tq[q], tb.pp[ty][k][q], acc[h][k]);
      }
      // stored at the position of the output entry (27 i + j): the last stage then reads
      // consecutive values (and a stride-27 transpose), both free of bank conflicts
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const int pos = sinv[ab * 9 + k];
#pragma unroll
        for (int h = 0; h < S::NB; ++h) B[h * 729 + pos] = acc[h][k];
      }
    }
    __syncthreads();
    // ---- A_ij = Bsym_ij + (Boff_ij + Boff_ji), E consecutive elements per (i, j) ---------------
    if (em) {
      // element-major output (nel, 27, 27): 729 consecutive values per element
      for (int item = tid; item < E * 729; item += T) {
        const int el = item / 729, ij = item % 729;
        if (base + el >= s.nel) break;
        const double *B = r1(el);
        double v = B[ij];
        if (LAP) v = v + (B[729 + ij] + B[729 + (ij % 27) * 27 + ij / 27]);
        out[(base + el) * 729 + ij] = v;
      }
    } else if (base + tid % E < s.nel) {                  // T % E == 0: el is fixed per thread
      const double *B = r1(tid % E);
      double *o = out + base + tid % E;
#pragma unroll 4
      for (int ij = tid / E; ij < 729; ij += T / E) {
        double v = B[ij];
        if (LAP) v = v + (B[729 + ij] + B[729 + (ij % 27) * 27 + ij / 27]);
        o[(int64_t)ij * s.nel] = v;                       // ij = 27 i + j
      }
    }
    // the next pass' first barrier orders these reads of B before the J columns are written
    // over them
  }
}

// launched by skb_local_hex_sumfact; `err` is the device zero-determinant flag
template <int FORM, int E, int T, int MINB>
static int launch_hex_sf(const skb_space_t &s, const HexSfTab &tb, int em, double *out, int *err,
                         cudaStream_t st) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t smem = sizeof(double) * (size_t)E * sf_per<FORM, E>();
  auto k = local_hex_sf_kernel<FORM, E, T, MINB>;
  SKB_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t need = (s.nel + E - 1) / E, cap = (int64_t)sms * MINB;
  const int grid = (int)(need < cap ? need : cap);
  k<<<grid, T, smem, st>>>(s, tb, em, out, err);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace skb

extern "C" int skb_local_hex_sumfact(const skb_space_t *space, int form, int32_t nq,
                                     const int32_t *qstride_host, const double *pp_host,
                                     const double *g_host, const uint8_t *bnode_host,
                                     const uint8_t *vtx_host, int32_t element_major,
                                     double *out_local, void *stream) {
  using namespace skb;
  const int em = element_major != 0;
  if (!space || !qstride_host || !pp_host || !g_host || !bnode_host || !vtx_host || !out_local)
    return SKB_EINVAL;
  const skb_space_t &s = *space;
  if (nq != SF_NQ || s.nqp != nq * nq * nq || s.nbs != 27 || s.ncomp != 1 || s.dim != 3 ||
      s.nnodes != 8 || s.mapping != SKB_MAP_ISO_HEX1 ||
      (form != SKB_FORM_LAPLACE && form != SKB_FORM_MASS))
    return SKB_EINVAL;
  if (s.nel <= 0) return SKB_OK;
  HexSfTab tb;
  memcpy(tb.pp, pp_host, sizeof(tb.pp));
  memcpy(tb.g, g_host, sizeof(tb.g));
  memset(tb.bnode, 0, sizeof(tb.bnode));
  memcpy(tb.bnode, bnode_host, 27);
  memcpy(tb.vtx, vtx_host, 8);
  for (int k = 0; k < 3; ++k) tb.qs[k] = qstride_host[k];
  for (int k = 0; k < 27; ++k)
    if (tb.bnode[k] >= 27) return SKB_EINVAL;
  for (int k = 0; k < 8; ++k)
    if (tb.vtx[k] >= 8) return SKB_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  DeviceFlag flag(st);
  SKB_CUDA_TRY(flag.init());
  // two CTAs of 2 elements per SM: 5.49 ms per C4 step against 5.85 ms for one CTA of 4 (the
  // barriers of one CTA are covered by the other); debug bits 4, 5 select the other shapes
  // that were measured (elements, threads, CTAs per SM) - profiles/r2_hex_sumfact.md
  const int var = (debug_flags() >> 4) & 3;
  int rc;
  if (form == SKB_FORM_MASS)
    rc = launch_hex_sf<SKB_FORM_MASS, 4, 256, 1>(s, tb, em, out_local, flag.p, st);
  else if (var == 1)
    rc = launch_hex_sf<SKB_FORM_LAPLACE, 4, 256, 1>(s, tb, em, out_local, flag.p, st);
  else if (var == 2)
    rc = launch_hex_sf<SKB_FORM_LAPLACE, 1, 96, 5>(s, tb, em, out_local, flag.p, st);
  else if (var == 3)
    rc = launch_hex_sf<SKB_FORM_LAPLACE, 1, 96, 4>(s, tb, em, out_local, flag.p, st);
  else
    rc = launch_hex_sf<SKB_FORM_LAPLACE, 2, 128, 2>(s, tb, em, out_local, flag.p, st);
  if (rc != SKB_OK) return rc;
  int herr = 0;
  SKB_CUDA_TRY(flag.read(&herr));
  if (herr) return SKB_EZERODET;
  return (int)cudaGetLastError();
}

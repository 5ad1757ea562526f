// Element-local matrices / vectors for tabulated H1 elements.
//
// Replaces CellBasis.__init__ (assembly/basis/cell_basis.py:94-106) and the
// Nbfun^2 Python loop of BilinearForm._assemble
// (assembly/form/bilinear_form.py:86-98,150-151): geometry, push-forward,
// integrand and the quadrature reduction are fused per element; the
// (Nbfun, dim, nel, nqp) basis arrays of the reference never exist.
//
// Two kernels:
//   local_affine_kernel : one thread per element (tri/tet, MappingAffine)
//   local_hex_kernel    : one CTA per element (MeshHex1 isoparametric map),
//                         per-qp inverse Jacobians staged in shared memory
#include "skb_common.cuh"

namespace skb {

// ---- integrands in the reference's operation order (Appendix A.6) ----------
// Scalar field at one quadrature point: value + gradient.
template <int DIM>
struct FS {
  double v;
  double g[DIM];
};
// Vector field phi_b * e_n (element_vector.py:36-48): dense, zero padded.
template <int DIM>
struct FV {
  double v[DIM];
  double g[DIM][DIM];
};

template <int DIM>
__device__ __forceinline__ double form_scalar(int form, const FS<DIM> &u, const FS<DIM> &v) {
  if (form == SKB_FORM_LAPLACE) {
    // helpers.dot = einsum('i...,i...') : sequential over i  (helpers.py:108-110)
    double acc = u.g[0] * v.g[0];
#pragma unroll
    for (int i = 1; i < DIM; ++i) acc = acc + u.g[i] * v.g[i];
    return acc;
  }
  return u.v * v.v;  // mass (models/poisson.py:17-19)
}

template <int DIM>
__device__ __forceinline__ double form_vector(int form, const FV<DIM> &u, const FV<DIM> &v,
                                              double lambda, double two_mu) {
  if (form == SKB_FORM_MASS) {  // dot(u, v)
    double acc = u.v[0] * v.v[0];
#pragma unroll
    for (int i = 1; i < DIM; ++i) acc = acc + u.v[i] * v.v[i];
    return acc;
  }
  if (form == SKB_FORM_VECTOR_LAPLACE) {
    // helpers.ddot = einsum('ij...,ij...') row-major (helpers.py:113-115)
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < DIM; ++i)
#pragma unroll
      for (int j = 0; j < DIM; ++j) acc = acc + u.g[i][j] * v.g[i][j];
    return acc;
  }
  // linear elasticity (models/elasticity.py:35-53):
  //   ddot(C(sym_grad(u)), sym_grad(v)),  C(T) = 2.*Mu*T + Lambda*eye(trace(T), d)
  double Su[DIM][DIM], Sv[DIM][DIM];
#pragma unroll
  for (int i = 0; i < DIM; ++i)
#pragma unroll
    for (int j = 0; j < DIM; ++j) {
      Su[i][j] = 0.5 * (u.g[i][j] + u.g[j][i]);  // helpers.py:71-73
      Sv[i][j] = 0.5 * (v.g[i][j] + v.g[j][i]);
    }
  double tr = Su[0][0];  // einsum('ii...')
#pragma unroll
  for (int i = 1; i < DIM; ++i) tr = tr + Su[i][i];
  double acc = 0.0;
#pragma unroll
  for (int i = 0; i < DIM; ++i)
#pragma unroll
    for (int j = 0; j < DIM; ++j) {
      double e = (i == j) ? tr : 0.0 * tr;     // helpers.eye (helpers.py:147-150)
      double c = two_mu * Su[i][j] + lambda * e;
      acc = acc + c * Sv[i][j];
    }
  return acc;
}

// push-forward grad_b[j] = sum_i inv[i][j]*dphi[i]  (element_h1.py:17,
// einsum('ijkl,il->jkl'): sequential in i, Appendix A.4)
template <int DIM>
__device__ __forceinline__ void push_grad(const double (*inv)[DIM], const double *dphi_b, int nqp,
                                          int q, double *g) {
#pragma unroll
  for (int j = 0; j < DIM; ++j) {
    double acc = inv[0][j] * dphi_b[q];
#pragma unroll
    for (int i = 1; i < DIM; ++i) acc = acc + inv[i][j] * dphi_b[i * nqp + q];
    g[j] = acc;
  }
}

template <int DIM>
__device__ __forceinline__ void make_vector_field(FV<DIM> &f, int n, double phi, const double *g) {
#pragma unroll
  for (int a = 0; a < DIM; ++a) {
    f.v[a] = (a == n) ? phi : 0.0;
#pragma unroll
    for (int b = 0; b < DIM; ++b) f.g[a][b] = (a == n) ? g[b] : 0.0;
  }
}

// shared-memory table layout: phi[nbs][nqp], dphi[nbs][DIM][nqp], W[nqp]
struct Tables {
  const double *phi, *dphi, *W;
};

__device__ __forceinline__ Tables stage_tables(double *smem, const skb_space_t &s) {
  const int nphi = s.nbs * s.nqp, ndphi = s.nbs * s.dim * s.nqp;
  for (int i = threadIdx.x; i < nphi; i += blockDim.x) smem[i] = s.phi[i];
  for (int i = threadIdx.x; i < ndphi; i += blockDim.x) smem[nphi + i] = s.dphi[i];
  for (int i = threadIdx.x; i < s.nqp; i += blockDim.x) smem[nphi + ndphi + i] = s.W[i];
  __syncthreads();
  Tables t;
  t.phi = smem;
  t.dphi = smem + nphi;
  t.W = smem + nphi + ndphi;
  return t;
}

template <int DIM, bool VEC, bool BILINEAR>
__global__ void __launch_bounds__(128)
local_affine_kernel(const skb_space_t s, int form, double lambda, double two_mu,
                    double *__restrict__ out) {
  extern __shared__ double smem[];
  const Tables tab = stage_tables(smem, s);
  const int nqp = s.nqp, nbs = s.nbs;
  constexpr int NC = VEC ? DIM : 1;
  const int nb = nbs * NC;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < s.nel;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t eg = s.tind ? (int64_t)s.tind[e] : e;
    Affine<DIM> g;
    affine_load<DIM>(g, s.p, s.npts, s.t, s.nel_total, eg);
    affine_invert(g);
    const double absdet = fabs(g.det);
    if (!BILINEAR) {
      // LinearForm._assemble (linear_form.py:41-44), unit_load: v.
      // v is a stride-0 broadcast and the affine dx is Fortran-ordered, so
      // numpy's v*dx is F-ordered and np.sum(axis=1) adds left to right.
      for (int ib = 0; ib < nbs; ++ib)
        for (int nv = 0; nv < NC; ++nv) {
          auto f = [&](int q) -> double {
            double dx = absdet * tab.W[q];
            return tab.phi[ib * nqp + q] * dx;
          };
          out[(int64_t)(ib * NC + nv) * s.nel + e] = seq_sum(nqp, f);
        }
      continue;
    }
    for (int jb = 0; jb < nbs; ++jb)
      for (int ib = 0; ib < nbs; ++ib) {
        const double *dj = tab.dphi + jb * DIM * nqp, *di = tab.dphi + ib * DIM * nqp;
#pragma unroll
        for (int nu = 0; nu < NC; ++nu)
#pragma unroll
          for (int nv = 0; nv < NC; ++nv) {
            auto f = [&](int q) -> double {
              double gu[DIM], gv[DIM];
              push_grad<DIM>(g.inv, dj, nqp, q, gu);
              push_grad<DIM>(g.inv, di, nqp, q, gv);
              double val;
              if (VEC) {
                FV<DIM> u, v;
                make_vector_field<DIM>(u, nu, tab.phi[jb * nqp + q], gu);
                make_vector_field<DIM>(v, nv, tab.phi[ib * nqp + q], gv);
                val = form_vector<DIM>(form, u, v, lambda, two_mu);
              } else {
                FS<DIM> u, v;
                u.v = tab.phi[jb * nqp + q];
                v.v = tab.phi[ib * nqp + q];
#pragma unroll
                for (int k = 0; k < DIM; ++k) { u.g[k] = gu[k]; v.g[k] = gv[k]; }
                val = form_scalar<DIM>(form, u, v);
              }
              double dx = absdet * tab.W[q];  // cell_basis.py:104-105
              return val * dx;                // bilinear_form.py:151
            };
            const int J = jb * NC + nu, I = ib * NC + nv;
            out[((int64_t)J * nb + I) * s.nel + e] = pw_sum(nqp, f);
          }
      }
  }
}

// ---------------------------------------------------------------------------
// Faster variant for rules with nqp <= LOCAL_MAXQ: the pushed gradients of the
// trial function (per jb) and of the test function (per ib) are computed once
// per quadrature point and kept in per-thread local memory, and vector-valued
// integrands are evaluated on their *structurally non-zero* terms only.
//
// Skipping a structural zero is exact: in the reference those terms are
// (0 * x) products added to the running sum, i.e. +-0.0, which leaves every
// partial sum unchanged (up to the sign of zero; inf/nan gradients excepted -
// they occur only on zero-volume elements and still give non-finite results).
// The surviving terms are added in the reference's row-major (i, j) order.
// ---------------------------------------------------------------------------
constexpr int LOCAL_MAXQ = 16;

// ddot(C(sym_grad(u)), sym_grad(v)) for u = phi_j e_nu, v = phi_i e_nv, given the
// scalar gradients gu, gv (models/elasticity.py:35-53, helpers.py:71-73,113-150).
//   sym_grad(u)[nu][nu] = gu[nu],  [nu][k] = [k][nu] = 0.5*gu[k],  trace = gu[nu]
//   C(T)[a][b] = 2.*Mu*T[a][b] + Lambda*(a==b ? tr : 0.*tr)
template <int DIM>
__device__ __forceinline__ double elasticity_sparse(int nu, int nv, const double *gu,
                                                    const double *gv, double lambda,
                                                    double two_mu) {
  double acc = 0.0;
  if (nu == nv) {
    const int n = nu;
    const double D = (two_mu * gu[n] + lambda * gu[n]) * gv[n];
#pragma unroll
    for (int a = 0; a < DIM; ++a)
#pragma unroll
      for (int b = 0; b < DIM; ++b) {
        if (a == n && b == n) acc = acc + D;
        else if (a == n) acc = acc + (two_mu * (0.5 * gu[b])) * (0.5 * gv[b]);
        else if (b == n) acc = acc + (two_mu * (0.5 * gu[a])) * (0.5 * gv[a]);
      }
  } else {
    const double L = (lambda * gu[nu]) * gv[nv];                    // at (nv, nv)
    const double M = (two_mu * (0.5 * gu[nv])) * (0.5 * gv[nu]);    // at (nu, nv) and (nv, nu)
#pragma unroll
    for (int a = 0; a < DIM; ++a)
#pragma unroll
      for (int b = 0; b < DIM; ++b) {
        if (a == nv && b == nv) acc = acc + L;
        else if ((a == nu && b == nv) || (a == nv && b == nu)) acc = acc + M;
      }
  }
  return acc;
}

// NQP > 0: the rule size is a compile-time constant - the gradient / dx arrays and the NQP
// terms of an entry then live in registers and numpy's pairwise sum unrolls (pw_sum_fixed).
// EM: element-major output (nel, Nbv, Nbu) for skb_csr_reduce_em.  The roles of the two loops
// are swapped (test function outside, trial function inside) so that a thread writes the row
// of its element it is working on left to right; every entry is still its own pairwise sum of
// the same terms, i.e. the same bits.
template <int DIM, bool VEC, int NQP = 0, bool EM = false>
__global__ void __launch_bounds__(128)
local_affine_cached_kernel(const skb_space_t s, int form, double lambda, double two_mu,
                           double *__restrict__ out) {
  extern __shared__ double smem[];
  const Tables tab = stage_tables(smem, s);
  const int nqp = NQP ? NQP : s.nqp, nbs = s.nbs;
  constexpr int MAXQ = NQP ? NQP : LOCAL_MAXQ;
  constexpr int NC = VEC ? DIM : 1;
  const int nb = nbs * NC;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < s.nel;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t eg = s.tind ? (int64_t)s.tind[e] : e;
    Affine<DIM> g;
    affine_load<DIM>(g, s.p, s.npts, s.t, s.nel_total, eg);
    affine_invert(g);
    const double absdet = fabs(g.det);
    double go[MAXQ][DIM], gi[MAXQ][DIM], dxq[MAXQ];
#pragma unroll
    for (int q = 0; q < MAXQ; ++q)
      if (q < nqp) dxq[q] = absdet * tab.W[q];                      // cell_basis.py:104-105
    for (int ob = 0; ob < nbs; ++ob) {                // outer function: trial (test when EM)
#pragma unroll
      for (int q = 0; q < MAXQ; ++q)
        if (q < nqp) push_grad<DIM>(g.inv, tab.dphi + ob * DIM * nqp, nqp, q, go[q]);
      for (int nb_ = 0; nb_ < nbs; ++nb_) {           // inner function: test (trial when EM)
#pragma unroll
        for (int q = 0; q < MAXQ; ++q)
          if (q < nqp) push_grad<DIM>(g.inv, tab.dphi + nb_ * DIM * nqp, nqp, q, gi[q]);
        const int jb = EM ? nb_ : ob, ib = EM ? ob : nb_;
        const double(*gu)[DIM] = EM ? gi : go;
        const double(*gv)[DIM] = EM ? go : gi;
        const double *pj = tab.phi + jb * nqp, *pi = tab.phi + ib * nqp;
        // nu, nv become compile-time constants after unrolling, which lets the
        // compiler keep only the structurally non-zero terms of the integrand
#pragma unroll
        for (int n0 = 0; n0 < NC; ++n0)
#pragma unroll
          for (int n1 = 0; n1 < NC; ++n1) {
            const int nu = EM ? n1 : n0, nv = EM ? n0 : n1;
            auto f = [&](int q) -> double {
              double val;
              if (!VEC) {
                if (form == SKB_FORM_LAPLACE) {
                  val = gu[q][0] * gv[q][0];
#pragma unroll
                  for (int k = 1; k < DIM; ++k) val = val + gu[q][k] * gv[q][k];
                } else {
                  val = pj[q] * pi[q];
                }
              } else if (form == SKB_FORM_ELASTICITY) {
                val = elasticity_sparse<DIM>(nu, nv, gu[q], gv[q], lambda, two_mu);
              } else if (nu != nv) {
                val = 0.0;                      // vector Laplace / mass: disjoint components
              } else if (form == SKB_FORM_VECTOR_LAPLACE) {
                val = 0.0;                      // row nu of grad(u) against row nu of grad(v)
#pragma unroll
                for (int k = 0; k < DIM; ++k) val = val + gu[q][k] * gv[q][k];
              } else {
                val = pj[q] * pi[q];            // dot(u, v)
              }
              return val * dxq[q];              // bilinear_form.py:151
            };
            const int J = jb * NC + nu, I = ib * NC + nv;
            double r;
            if (NQP) {
              double term[MAXQ];
#pragma unroll
              for (int q = 0; q < MAXQ; ++q) term[q] = f(q);
              r = pw_sum_fixed<MAXQ>(term);
            } else {
              r = pw_sum(nqp, f);
            }
            if (EM) out[((int64_t)e * nb + I) * nb + J] = r;
            else out[((int64_t)J * nb + I) * s.nel + e] = r;
          }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Scalar elements, symmetric forms (laplace, mass), rules of NQP points known at
// compile time.  Two facts make this much cheaper than local_affine_kernel:
//  * both integrands are bitwise symmetric in (u, v) - products commute, the
//    sequence of additions is the same - so K[j][i] == K[i][j] bit for bit and
//    only the pairs i <= j are evaluated (the same observation the fused P1
//    path uses);
//  * with NQP a template parameter the pushed-forward gradients of row i at all
//    quadrature points (NQP x DIM doubles), dx and the NQP terms of one entry
//    live in registers; only the gradients of the column function are recomputed
//    per pair.  P2 tetrahedra: 14.4 k instead of 40.7 k FP64 operations.
// numpy's pairwise sum is replicated with compile-time loops (pw_sum_fixed).
// ---------------------------------------------------------------------------
template <int DIM, int NQP, bool EM = false>
__global__ void __launch_bounds__(128)
local_affine_sym_kernel(const skb_space_t s, int form, double *__restrict__ out) {
  extern __shared__ double smem[];
  const Tables tab = stage_tables(smem, s);
  const int nbs = s.nbs;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < s.nel;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t eg = s.tind ? (int64_t)s.tind[e] : e;
    Affine<DIM> g;
    affine_load<DIM>(g, s.p, s.npts, s.t, s.nel_total, eg);
    affine_invert(g);
    const double absdet = fabs(g.det);
    double dxq[NQP];
#pragma unroll
    for (int q = 0; q < NQP; ++q) dxq[q] = absdet * tab.W[q];      // cell_basis.py:104-105
    for (int ib = 0; ib < nbs; ++ib) {
      double gi[NQP][DIM];
      if (form == SKB_FORM_LAPLACE) {
#pragma unroll
        for (int q = 0; q < NQP; ++q) push_grad<DIM>(g.inv, tab.dphi + ib * DIM * NQP, NQP, q, gi[q]);
      }
      const double *pi = tab.phi + ib * NQP;
      for (int jb = ib; jb < nbs; ++jb) {
        const double *dj = tab.dphi + jb * DIM * NQP, *pj = tab.phi + jb * NQP;
        double term[NQP];
#pragma unroll
        for (int q = 0; q < NQP; ++q) {
          double val;
          if (form == SKB_FORM_LAPLACE) {
            double gj[DIM];
            push_grad<DIM>(g.inv, dj, NQP, q, gj);
            // u = trial function jb, v = test function ib (bilinear_form.py:88-91); the
            // products commute, so entry (ib, jb) has the same bits
            val = gj[0] * gi[q][0];
#pragma unroll
            for (int k = 1; k < DIM; ++k) val = val + gj[k] * gi[q][k];
          } else {
            val = pj[q] * pi[q];
          }
          term[q] = val * dxq[q];                                   // bilinear_form.py:151
        }
        const double v = pw_sum_fixed<NQP>(term);
        if (EM) {                                   // (nel, Nbv, Nbu), see the cached kernel
          out[((int64_t)e * nbs + ib) * nbs + jb] = v;
          if (jb != ib) out[((int64_t)e * nbs + jb) * nbs + ib] = v;
        } else {
          out[((int64_t)jb * nbs + ib) * s.nel + e] = v;
          if (jb != ib) out[((int64_t)ib * nbs + jb) * s.nel + e] = v;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Hexahedra: isoparametric trilinear map (mapping_isoparametric.py:112-226).
// One CTA per element.  Phase 1: threads over q compute J, det, inv(e,q),
// dx(e,q) into shared memory.  Phase 2: threads over local entries (j,i),
// each walks q in numpy's pairwise order.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void hex_jacobian(const skb_space_t &s, const double (*xn)[8], int q,
                                             double (*J)[3]) {
  const int nqp = s.nqp;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double acc = 0.0;  // out = zeros; out += p*dphi  (:115-119)
      for (int n = 0; n < 8; ++n) acc = acc + xn[i][n] * __ldg(s.mdphi + (n * 3 + j) * nqp + q);
      J[i][j] = acc;
    }
}

template <bool BILINEAR>
__global__ void __launch_bounds__(256)
local_hex_kernel(const skb_space_t s, int form, double *__restrict__ out, int *__restrict__ err) {
  extern __shared__ double smem[];  // inv[9][nqp], dx[nqp]
  const int nqp = s.nqp, nbs = s.nbs;
  double *s_inv = smem, *s_dx = smem + 9 * nqp;
  __shared__ double xn[3][8];
  for (int64_t e = blockIdx.x; e < s.nel; e += gridDim.x) {
    const int64_t eg = s.tind ? (int64_t)s.tind[e] : e;
    __syncthreads();
    if (threadIdx.x < 24) {
      int i = threadIdx.x / 8, n = threadIdx.x % 8;
      xn[i][n] = s.p[(int64_t)i * s.npts + s.t[(int64_t)n * s.nel_total + eg]];
    }
    __syncthreads();
    for (int q = threadIdx.x; q < nqp; q += blockDim.x) {
      double J[3][3], nn[3][3], inv[3][3];
      hex_jacobian(s, xn, q, J);
      double det = det3(J);
      if (det == 0.0) atomicExch(err, 1);  // mapping_isoparametric.py:195-196
      cofactors3(J, nn);
      divide9(nn, det, inv);
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) s_inv[(i * 3 + j) * nqp + q] = inv[i][j];
      s_dx[q] = fabs(det) * __ldg(s.W + q);
    }
    __syncthreads();
    if (!BILINEAR) {
      for (int ib = threadIdx.x; ib < nbs; ib += blockDim.x) {
        auto f = [&](int q) -> double { return __ldg(s.phi + ib * nqp + q) * s_dx[q]; };
        out[(int64_t)ib * s.nel + e] = pw_sum(nqp, f);
      }
      continue;
    }
    for (int ent = threadIdx.x; ent < nbs * nbs; ent += blockDim.x) {
      const int jb = ent / nbs, ib = ent % nbs;
      const double *dj = s.dphi + jb * 3 * nqp, *di = s.dphi + ib * 3 * nqp;
      auto f = [&](int q) -> double {
        FS<3> u, v;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          double au = s_inv[(0 * 3 + j) * nqp + q] * __ldg(dj + q);
          double av = s_inv[(0 * 3 + j) * nqp + q] * __ldg(di + q);
#pragma unroll
          for (int i = 1; i < 3; ++i) {
            au = au + s_inv[(i * 3 + j) * nqp + q] * __ldg(dj + i * nqp + q);
            av = av + s_inv[(i * 3 + j) * nqp + q] * __ldg(di + i * nqp + q);
          }
          u.g[j] = au;
          v.g[j] = av;
        }
        u.v = __ldg(s.phi + jb * nqp + q);
        v.v = __ldg(s.phi + ib * nqp + q);
        return form_scalar<3>(form, u, v) * s_dx[q];
      };
      out[(int64_t)ent * s.nel + e] = pw_sum(nqp, f);
    }
  }
}

int launch_hex_mma(const skb_space_t &s, int form, double *out, int *err, cudaStream_t st);

static int grid_for(int64_t work_items, int block, int per_sm) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t need = (work_items + block - 1) / block;
  int64_t cap = (int64_t)sms * per_sm;
  return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

template <bool BILINEAR, bool EM = false>
static int launch_local(const skb_space_t *sp, int form, const double *params, double *out,
                        cudaStream_t st) {
  if (!sp || sp->nel < 0) return SKB_EINVAL;
  if (sp->nel == 0) return SKB_OK;
  if (!out) return SKB_EINVAL;
  const skb_space_t s = *sp;
  const bool vec = s.ncomp > 1;
  if (vec && s.ncomp != s.dim) return SKB_EINVAL;
  const double lambda = params ? params[0] : 1.0, two_mu = params ? params[1] : 2.0;
  if (BILINEAR) {
    if (form < 0 || form > SKB_FORM_ELASTICITY) return SKB_EINVAL;
    if (!vec && form > SKB_FORM_MASS) return SKB_EINVAL;
    if (vec && form == SKB_FORM_LAPLACE) return SKB_EINVAL;
  } else if (form != SKB_LFORM_UNIT_LOAD || vec) {
    return SKB_EINVAL;
  }
  if (s.mapping == SKB_MAP_AFFINE) {
    if (s.nnodes != s.dim + 1) return SKB_EINVAL;
    size_t smem = sizeof(double) * ((size_t)s.nbs * (1 + s.dim) * s.nqp + s.nqp);
    if (smem > 200 * 1024) return SKB_ETOOBIG;
    const int block = 128;
    const int grid = grid_for(s.nel, block, 16);
#define SKB_LAUNCH_AFFINE(D, V)                                                              \
  do {                                                                                       \
    auto k = local_affine_kernel<D, V, BILINEAR>;                                            \
    if (smem > 48 * 1024)                                                                    \
      SKB_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                        (int)smem));                                         \
    k<<<grid, block, smem, st>>>(s, form, lambda, two_mu, out);                              \
    count_launch();                                                                          \
  } while (0)
#define SKB_LAUNCH_CACHED(D, V)                                                              \
  do {                                                                                       \
    auto k = local_affine_cached_kernel<D, V>;                                               \
    if (smem > 48 * 1024)                                                                    \
      SKB_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                        (int)smem));                                         \
    k<<<grid, block, smem, st>>>(s, form, lambda, two_mu, out);                              \
    count_launch();                                                                          \
  } while (0)
    // scalar symmetric forms with a rule size known at compile time: register-cached,
    // upper triangle only
    if (BILINEAR && !vec && !(debug_flags() & 8)) {
#define SKB_LAUNCH_SYM(D, Q)                                                                 \
  if (s.dim == D && s.nqp == Q) {                                                            \
    auto k = local_affine_sym_kernel<D, Q, EM>;                                                \
    if (smem > 48 * 1024)                                                                    \
      SKB_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                        (int)smem));                                         \
    k<<<grid, block, smem, st>>>(s, form, out);                                              \
    count_launch();                                                                          \
    return (int)cudaGetLastError();                                                          \
  }
      SKB_LAUNCH_SYM(3, 4)
      SKB_LAUNCH_SYM(3, 11)
      SKB_LAUNCH_SYM(2, 3)
      SKB_LAUNCH_SYM(2, 6)
#undef SKB_LAUNCH_SYM
    }
    // other scalar cases: recomputing the push-forward per pair is cheaper than
    // the local-memory round trip (measured), so only vector elements - whose
    // dense integrand is 7x more FP64 work - take the cached/sparse kernel
    if (BILINEAR && vec && !(debug_flags() & 8)) {
#define SKB_LAUNCH_CACHED_FIXED(D, Q)                                                        \
  if (s.dim == D && s.nqp == Q) {                                                            \
    auto k = local_affine_cached_kernel<D, true, Q, EM>;                                       \
    if (smem > 48 * 1024)                                                                    \
      SKB_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                        (int)smem));                                         \
    k<<<grid, block, smem, st>>>(s, form, lambda, two_mu, out);                              \
    count_launch();                                                                          \
    return (int)cudaGetLastError();                                                          \
  }
      SKB_LAUNCH_CACHED_FIXED(3, 11)
      SKB_LAUNCH_CACHED_FIXED(3, 4)
      SKB_LAUNCH_CACHED_FIXED(2, 6)
      SKB_LAUNCH_CACHED_FIXED(2, 3)
#undef SKB_LAUNCH_CACHED_FIXED
    }
    if (EM) return SKB_EINVAL;       // element-major output: the fixed-rule kernels above only
    if (BILINEAR && vec && s.nqp <= LOCAL_MAXQ && !(debug_flags() & 8)) {
      if (s.dim == 2 && !vec) SKB_LAUNCH_CACHED(2, false);
      else if (s.dim == 2 && vec) SKB_LAUNCH_CACHED(2, true);
      else if (s.dim == 3 && !vec) SKB_LAUNCH_CACHED(3, false);
      else if (s.dim == 3 && vec) SKB_LAUNCH_CACHED(3, true);
      else return SKB_EINVAL;
      return (int)cudaGetLastError();
    }
#undef SKB_LAUNCH_CACHED
    if (s.dim == 2 && !vec) SKB_LAUNCH_AFFINE(2, false);
    else if (s.dim == 2 && vec) SKB_LAUNCH_AFFINE(2, true);
    else if (s.dim == 3 && !vec) SKB_LAUNCH_AFFINE(3, false);
    else if (s.dim == 3 && vec) SKB_LAUNCH_AFFINE(3, true);
    else return SKB_EINVAL;
#undef SKB_LAUNCH_AFFINE
    return (int)cudaGetLastError();
  }
  if (s.mapping == SKB_MAP_ISO_HEX1) {
    if (EM) return SKB_EINVAL;
    if (s.dim != 3 || s.nnodes != 8 || vec || !s.mdphi) return SKB_EINVAL;
    size_t smem = sizeof(double) * 10 * (size_t)s.nqp;
    if (smem > 200 * 1024) return SKB_ETOOBIG;
    DeviceFlag flag(st);
    SKB_CUDA_TRY(flag.init());
    int *err = flag.p;
    if (BILINEAR && s.nbs > 8 && s.nbs <= 32 && (form == SKB_FORM_LAPLACE || form == SKB_FORM_MASS) &&
        !(debug_flags() & 8)) {
      // high-order hexes (value-level parity): Gram-matrix contraction on the FP64
      // tensor cores, csrc/skb_hex_mma.cu; debug bit 3 keeps the scalar kernel
      const int rc = launch_hex_mma(s, form, out, err, st);
      if (rc != SKB_OK) return rc;
    } else {
      auto k = local_hex_kernel<BILINEAR>;
      if (smem > 48 * 1024)
        SKB_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
      const int grid = grid_for(s.nel * 256, 256, 8);
      k<<<grid, 256, smem, st>>>(s, form, out, err);
      count_launch();
    }
    int herr = 0;
    SKB_CUDA_TRY(flag.read(&herr));
    if (herr) return SKB_EZERODET;
    return (int)cudaGetLastError();
  }
  return SKB_EINVAL;
}

}  // namespace skb

extern "C" int skb_local_bilinear(const skb_space_t *space, int form, const double *params_host,
                                  double *out_local, void *stream) {
  return skb::launch_local<true>(space, form, params_host, out_local, (cudaStream_t)stream);
}

extern "C" int skb_local_bilinear_em(const skb_space_t *space, int form,
                                     const double *params_host, double *out_local_em,
                                     void *stream) {
  return skb::launch_local<true, true>(space, form, params_host, out_local_em,
                                       (cudaStream_t)stream);
}

extern "C" int skb_local_linear(const skb_space_t *space, int form, const double *params_host,
                                double *out_local, void *stream) {
  return skb::launch_local<false>(space, form, params_host, out_local, (cudaStream_t)stream);
}

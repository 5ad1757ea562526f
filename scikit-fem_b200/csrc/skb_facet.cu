// FacetBasis on affine (tri / tet) meshes: quadrature on mesh facets.
//
// Replaces FacetBasis.__init__ (assembly/basis/facet_basis.py:76-116) and the
// facet part of MappingAffine (mapping/mapping_affine.py:154-181 B, c, detB;
// :234-246 G; :195-203 invF; :248-281 normals).  In the reference every facet
// gets its own local quadrature points Y = invF(G(X)), at which lbasis is
// re-evaluated, so basis values are genuine (nfacets, nqp) arrays.  Here:
//
//   skb_facet_geometry  per facet f (one thread): x = G(X), Y = invF(x), dx =
//                       |detB| W, unit normal n, |detB| - same operation order
//   skb_facet_basis     per (facet, q): evaluates the element's monomial tables
//                       at Y (the term order of the reference's polynomials, so
//                       values are bit-identical) and pushes the gradient
//                       forward through invA of the facet's element
// Forms on a FacetBasis run through the traced path (skfem_b200/form.py).
#include "skb_common.cuh"

namespace skb {

constexpr int POLY_MAXT = 12;  // terms per polynomial

// |detB| of the facet map, mapping_affine.py:170-181
__device__ __forceinline__ double facet_detB(const double (*B)[1]) {
  return sqrt(B[0][0] * B[0][0] + B[1][0] * B[1][0]);
}
__device__ __forceinline__ double facet_detB(const double (*B)[2]) {
  const double a = B[1][0] * B[2][1] - B[2][0] * B[1][1];
  const double b = -B[0][0] * B[2][1] + B[2][0] * B[0][1];
  const double c = B[0][0] * B[1][1] - B[1][0] * B[0][1];
  return sqrt((a * a + b * b) + c * c);
}

template <int DIM>
__global__ void __launch_bounds__(128)
facet_geometry_kernel(const skb_space_t s, const int32_t *__restrict__ facets,
                      int64_t nfacets_total, const int32_t *__restrict__ find,
                      const int32_t *__restrict__ tind, const int32_t *__restrict__ tind_n,
                      const int32_t *__restrict__ lfacet, int64_t nf, const double *__restrict__ Xb,
                      const double *__restrict__ Wb, int nqp, double *__restrict__ x,
                      double *__restrict__ Y, double *__restrict__ dx, double *__restrict__ nrm,
                      double *__restrict__ detabs) {
  // reference normals of the local facets (mapping_affine.py:249-262)
  const double nref2[3][2] = {{0., -1.}, {1., 1.}, {-1., 0.}};
  const double nref3[4][3] = {{0., 0., -1.}, {0., -1., 0.}, {-1., 0., 0.}, {1., 1., 1.}};
  for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < nf;
       f += (int64_t)gridDim.x * blockDim.x) {
    const int64_t gf = find[f];
    double B[DIM][DIM - 1], c[DIM];
    {
      const int32_t v0 = facets[gf];
#pragma unroll
      for (int i = 0; i < DIM; ++i) {
        const double *pi = s.p + (int64_t)i * s.npts;
        c[i] = pi[v0];
#pragma unroll
        for (int j = 0; j < DIM - 1; ++j)
          B[i][j] = pi[facets[(int64_t)(j + 1) * nfacets_total + gf]] - pi[v0];
      }
    }
    const double detB = facet_detB(B);
    Affine<DIM> g, gn;
    affine_load<DIM>(g, s.p, s.npts, s.t, s.nel_total, tind[f]);
    affine_invert(g);
    const bool same = tind_n[f] == tind[f];
    if (!same) {
      affine_load<DIM>(gn, s.p, s.npts, s.t, s.nel_total, tind_n[f]);
      affine_invert(gn);
    }
    const Affine<DIM> &gq = same ? g : gn;
    // unit normal: n = invDF^T Nref / |.|  (einsum('ijkl,ik->jkl'), then 1./length)
    double n[DIM];
    {
      const int lf = lfacet[f];
      double s2 = 0.0;
#pragma unroll
      for (int j = 0; j < DIM; ++j) {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < DIM; ++i) {
          const double N = (DIM == 2) ? nref2[lf][i] : nref3[lf][i];
          acc = (i == 0) ? gq.inv[i][j] * N : acc + gq.inv[i][j] * N;
        }
        n[j] = acc;
        s2 = (j == 0) ? acc * acc : s2 + acc * acc;
      }
      const double rl = 1. / sqrt(s2);
#pragma unroll
      for (int j = 0; j < DIM; ++j) n[j] = n[j] * rl;
    }
    for (int q = 0; q < nqp; ++q) {
      double xq[DIM];
#pragma unroll
      for (int i = 0; i < DIM; ++i) {
        double acc = B[i][0] * Xb[q];
#pragma unroll
        for (int j = 1; j < DIM - 1; ++j) acc = acc + B[i][j] * Xb[j * nqp + q];
        xq[i] = acc + c[i];                                   // G(X), :234-243
        if (x) x[((int64_t)i * nf + f) * nqp + q] = xq[i];
      }
      if (Y) {
#pragma unroll
        for (int i = 0; i < DIM; ++i) {                       // invF, :195-203
          double acc = g.inv[i][0] * (xq[0] - g.b[0]);
#pragma unroll
          for (int j = 1; j < DIM; ++j) acc = acc + g.inv[i][j] * (xq[j] - g.b[j]);
          Y[((int64_t)i * nf + f) * nqp + q] = acc;
        }
      }
      if (dx) dx[f * nqp + q] = fabs(detB) * Wb[q];
      if (detabs) detabs[f * nqp + q] = fabs(detB);
      if (nrm) {
#pragma unroll
        for (int j = 0; j < DIM; ++j) nrm[((int64_t)j * nf + f) * nqp + q] = n[j];
      }
    }
  }
}

// One polynomial: terms (coef, exponents) evaluated left to right exactly like
// skfem_b200/element.py::_poly (== the reference's written expression order).
template <int DIM>
__device__ __forceinline__ double eval_poly(const double *coef, const int32_t *expo, int nterm,
                                            const double *X) {
  double acc = 0.0;
  bool any_mono = false;
  for (int k = 0; k < nterm; ++k) {
    const int e = expo[k];
    double m = 0.0;
    bool has = false;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      const int p = (e >> (8 * d)) & 0xff;
      for (int r = 0; r < p; ++r) {
        m = has ? m * X[d] : X[d];
        has = true;
      }
    }
    const double t = has ? coef[k] * m : coef[k];
    acc = (k == 0) ? t : acc + t;
    any_mono |= has;
  }
  if (!any_mono) acc = acc + 0.0 * X[0];   // constant polynomial: c + 0*x
  return acc;
}

template <int DIM>
__global__ void __launch_bounds__(128)
facet_basis_kernel(const skb_space_t s, const int32_t *__restrict__ tind, int64_t nf, int nqp,
                   const double *__restrict__ Y, const double *__restrict__ coef,
                   const int32_t *__restrict__ expo, const int32_t *__restrict__ nterm, int b,
                   double *__restrict__ value, double *__restrict__ grad) {
  const int64_t total = nf * nqp;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t f = idx / nqp;
    const int q = (int)(idx - f * nqp);
    double X[DIM];
#pragma unroll
    for (int i = 0; i < DIM; ++i) X[i] = Y[((int64_t)i * nf + f) * nqp + q];
    const int base = b * (1 + DIM);
    value[idx] = eval_poly<DIM>(coef + (int64_t)base * POLY_MAXT, expo + (int64_t)base * POLY_MAXT,
                                nterm[base], X);
    if (grad) {
      Affine<DIM> g;
      affine_load<DIM>(g, s.p, s.npts, s.t, s.nel_total, tind[f]);
      affine_invert(g);
      double d[DIM];
#pragma unroll
      for (int i = 0; i < DIM; ++i)
        d[i] = eval_poly<DIM>(coef + (int64_t)(base + 1 + i) * POLY_MAXT,
                              expo + (int64_t)(base + 1 + i) * POLY_MAXT, nterm[base + 1 + i], X);
#pragma unroll
      for (int j = 0; j < DIM; ++j) {  // einsum('ijkl,ikl->jkl', invDF, dphi)
        double acc = g.inv[0][j] * d[0];
#pragma unroll
        for (int i = 1; i < DIM; ++i) acc = acc + g.inv[i][j] * d[i];
        grad[((int64_t)j * nf + f) * nqp + q] = acc;
      }
    }
  }
}

static inline int fblk(int64_t n) {
  int64_t g = (n + 127) / 128;
  if (g < 1) g = 1;
  if (g > 148 * 32) g = 148 * 32;
  return (int)g;
}

}  // namespace skb

extern "C" int skb_facet_geometry(const skb_space_t *space, const int32_t *facets,
                                  int64_t nfacets_total, const int32_t *find,
                                  const int32_t *tind, const int32_t *tind_normals,
                                  const int32_t *lfacet, int64_t nf, const double *Xb,
                                  const double *Wb, int32_t nqp, double *x, double *Y, double *dx,
                                  double *normals, double *detabs, void *stream) {
  using namespace skb;
  if (!space || nf < 0 || nqp <= 0) return SKB_EINVAL;
  if (nf == 0) return SKB_OK;
  const skb_space_t s = *space;
  if (s.mapping != SKB_MAP_AFFINE) return SKB_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (s.dim == 2)
    facet_geometry_kernel<2><<<fblk(nf), 128, 0, st>>>(s, facets, nfacets_total, find, tind,
                                                       tind_normals, lfacet, nf, Xb, Wb, nqp, x, Y,
                                                       dx, normals, detabs);
  else if (s.dim == 3)
    facet_geometry_kernel<3><<<fblk(nf), 128, 0, st>>>(s, facets, nfacets_total, find, tind,
                                                       tind_normals, lfacet, nf, Xb, Wb, nqp, x, Y,
                                                       dx, normals, detabs);
  else
    return SKB_EINVAL;
  count_launch();
  return (int)cudaGetLastError();
}

extern "C" int skb_facet_basis(const skb_space_t *space, const int32_t *tind, int64_t nf,
                               int32_t nqp, const double *Y, const double *poly_coef,
                               const int32_t *poly_expo, const int32_t *poly_nterm, int32_t b,
                               double *value, double *grad, void *stream) {
  using namespace skb;
  if (!space || nf < 0 || nqp <= 0) return SKB_EINVAL;
  if (nf == 0) return SKB_OK;
  if (!value) return SKB_EINVAL;
  const skb_space_t s = *space;
  if (s.mapping != SKB_MAP_AFFINE || b < 0 || b >= s.nbs) return SKB_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (s.dim == 2)
    facet_basis_kernel<2><<<fblk(nf * nqp), 128, 0, st>>>(s, tind, nf, nqp, Y, poly_coef,
                                                          poly_expo, poly_nterm, b, value, grad);
  else if (s.dim == 3)
    facet_basis_kernel<3><<<fblk(nf * nqp), 128, 0, st>>>(s, tind, nf, nqp, Y, poly_coef,
                                                          poly_expo, poly_nterm, b, value, grad);
  else
    return SKB_EINVAL;
  count_launch();
  return (int)cudaGetLastError();
}

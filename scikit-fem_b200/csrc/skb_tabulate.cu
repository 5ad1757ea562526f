// Materialised global basis / geometry at quadrature points, for the traced
// path of user-defined forms: the user's Python callable is evaluated with
// device arrays of the reference's shapes, so the arrays the reference builds
// in CellBasis.__init__ must exist for those forms (and only for those).
//
//   grad (dim, nel, nqp)  ElementH1.gbasis           element/element_h1.py:10-18
//   dx   (nel, nqp)       |detDF| * W                assembly/basis/cell_basis.py:104-105
//   x    (dim, nel, nqp)  Mapping.F(X)               mapping_affine.py:183-193,
//                                                    mapping_isoparametric.py:52-58,170-171
//   detabs (nel, nqp)     |detDF| (for w.h)          cell_basis.py:136-141
// plus the numpy-order quadrature reduction of bilinear_form.py:150-151.
#include <atomic>
#include "skb_common.cuh"

namespace skb {

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
static std::atomic<int> g_debug{0};
int debug_flags() { return g_debug.load(std::memory_order_relaxed); }
static std::atomic<int> g_sm_reserve{0};
int sm_reserve() { return g_sm_reserve.load(std::memory_order_relaxed); }

template <int DIM>
__global__ void __launch_bounds__(128)
tabulate_affine_kernel(const skb_space_t s, int b, double *__restrict__ grad,
                       double *__restrict__ dx, double *__restrict__ x,
                       double *__restrict__ detabs) {
  const int nqp = s.nqp;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < s.nel;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t eg = s.tind ? (int64_t)s.tind[e] : e;
    Affine<DIM> g;
    affine_load<DIM>(g, s.p, s.npts, s.t, s.nel_total, eg);
    affine_invert(g);
    const double absdet = fabs(g.det);
    for (int q = 0; q < nqp; ++q) {
      if (grad) {
        const double *d = s.dphi + (int64_t)b * DIM * nqp;
#pragma unroll
        for (int j = 0; j < DIM; ++j) {
          double acc = g.inv[0][j] * __ldg(d + q);
#pragma unroll
          for (int i = 1; i < DIM; ++i) acc = acc + g.inv[i][j] * __ldg(d + i * nqp + q);
          grad[((int64_t)j * s.nel + e) * nqp + q] = acc;
        }
      }
      if (dx) dx[e * nqp + q] = absdet * __ldg(s.W + q);
      if (detabs) detabs[e * nqp + q] = absdet;
      if (x) {
#pragma unroll
        for (int i = 0; i < DIM; ++i) {
          // einsum('ijk,jl', A, X) sequential in j, then + b
          double acc = g.A[i][0] * __ldg(s.X + q);
#pragma unroll
          for (int j = 1; j < DIM; ++j) acc = acc + g.A[i][j] * __ldg(s.X + j * nqp + q);
          x[((int64_t)i * s.nel + e) * nqp + q] = acc + g.b[i];
        }
      }
    }
  }
}

__global__ void __launch_bounds__(128)
tabulate_hex_kernel(const skb_space_t s, int b, double *__restrict__ grad, double *__restrict__ dx,
                    double *__restrict__ x, double *__restrict__ detabs, int *__restrict__ err) {
  const int nqp = s.nqp;
  const int64_t total = s.nel * nqp;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = idx / nqp;
    const int q = (int)(idx - e * nqp);
    const int64_t eg = s.tind ? (int64_t)s.tind[e] : e;
    double xn[3][8];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const int32_t v = s.t[(int64_t)n * s.nel_total + eg];
#pragma unroll
      for (int i = 0; i < 3; ++i) xn[i][n] = s.p[(int64_t)i * s.npts + v];
    }
    double J[3][3], nn[3][3], inv[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double acc = 0.0;
#pragma unroll
        for (int n = 0; n < 8; ++n) acc = acc + xn[i][n] * __ldg(s.mdphi + (n * 3 + j) * nqp + q);
        J[i][j] = acc;
      }
    const double det = det3(J);
    if (det == 0.0) atomicExch(err, 1);
    if (grad) {
      cofactors3(J, nn);
      divide9(nn, det, inv);
      const double *d = s.dphi + (int64_t)b * 3 * nqp;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double acc = inv[0][j] * __ldg(d + q);
#pragma unroll
        for (int i = 1; i < 3; ++i) acc = acc + inv[i][j] * __ldg(d + i * nqp + q);
        grad[((int64_t)j * s.nel + e) * nqp + q] = acc;
      }
    }
    if (dx) dx[idx] = fabs(det) * __ldg(s.W + q);
    if (detabs) detabs[idx] = fabs(det);
    if (x) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        double acc = 0.0;  // Fmap: out = zeros; out += p * phi   (:52-58)
#pragma unroll
        for (int n = 0; n < 8; ++n) acc = acc + xn[i][n] * __ldg(s.mphi + n * nqp + q);
        x[((int64_t)i * s.nel + e) * nqp + q] = acc;
      }
    }
  }
}

// Mapping.DF / invDF / detDF at the quadrature points (mapping/mapping.py:6-114): the arrays
// of shape (dim, dim, nel, nqp) and (nel, nqp) the reference's Mapping classes return
// (mapping_affine.py:205-232, mapping_isoparametric.py:173-226), signed determinant.
template <int DIM>
__global__ void __launch_bounds__(128)
mapping_affine_kernel(const skb_space_t s, double *__restrict__ DF, double *__restrict__ invDF,
                      double *__restrict__ det) {
  const int nqp = s.nqp;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < s.nel;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t eg = s.tind ? (int64_t)s.tind[e] : e;
    Affine<DIM> g;
    affine_load<DIM>(g, s.p, s.npts, s.t, s.nel_total, eg);
    affine_invert(g);
    for (int q = 0; q < nqp; ++q) {
#pragma unroll
      for (int i = 0; i < DIM; ++i)
#pragma unroll
        for (int j = 0; j < DIM; ++j) {
          const int64_t o = (((int64_t)(i * DIM + j)) * s.nel + e) * nqp + q;
          if (DF) DF[o] = g.A[i][j];
          if (invDF) invDF[o] = g.inv[i][j];
        }
      if (det) det[e * nqp + q] = g.det;
    }
  }
}

__global__ void __launch_bounds__(128)
mapping_hex_kernel(const skb_space_t s, double *__restrict__ DF, double *__restrict__ invDF,
                   double *__restrict__ detout, int *__restrict__ err) {
  const int nqp = s.nqp;
  const int64_t total = s.nel * nqp;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = idx / nqp;
    const int q = (int)(idx - e * nqp);
    const int64_t eg = s.tind ? (int64_t)s.tind[e] : e;
    double J[3][3], nn[3][3], inv[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double acc = 0.0;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
          const int32_t v = s.t[(int64_t)n * s.nel_total + eg];
          acc = acc + s.p[(int64_t)i * s.npts + v] * __ldg(s.mdphi + (n * 3 + j) * nqp + q);
        }
        J[i][j] = acc;
      }
    const double det = det3(J);
    if (det == 0.0) atomicExch(err, 1);
    cofactors3(J, nn);
    divide9(nn, det, inv);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int64_t o = (((int64_t)(i * 3 + j)) * s.nel + e) * nqp + q;
        if (DF) DF[o] = J[i][j];
        if (invDF) invDF[o] = inv[i][j];
      }
    if (detout) detout[idx] = det;
  }
}

__global__ void __launch_bounds__(128)
qp_reduce_kernel(const double *__restrict__ integrand, const double *__restrict__ dx, int64_t nel,
                 int nqp, int sequential, double *__restrict__ out) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nel;
       e += (int64_t)gridDim.x * blockDim.x) {
    const double *a = integrand + e * nqp, *d = dx + e * nqp;
    auto f = [&](int q) -> double { return a[q] * d[q]; };
    out[e] = sequential ? seq_sum(nqp, f) : pw_sum(nqp, f);
  }
}

static inline int nblk(int64_t n, int block) {
  int64_t g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > 148 * 32) g = 148 * 32;
  return (int)g;
}

}  // namespace skb

extern "C" int skb_tabulate(const skb_space_t *space, int b, double *grad, double *dx, double *x,
                            double *detabs, void *stream) {
  using namespace skb;
  if (!space) return SKB_EINVAL;
  const skb_space_t s = *space;
  cudaStream_t st = (cudaStream_t)stream;
  if (s.nel == 0) return SKB_OK;
  if (grad && (b < 0 || b >= s.nbs)) return SKB_EINVAL;
  if (s.mapping == SKB_MAP_AFFINE) {
    if (x && !s.X) return SKB_EINVAL;
    if (s.dim == 2)
      tabulate_affine_kernel<2><<<nblk(s.nel, 128), 128, 0, st>>>(s, b, grad, dx, x, detabs);
    else if (s.dim == 3)
      tabulate_affine_kernel<3><<<nblk(s.nel, 128), 128, 0, st>>>(s, b, grad, dx, x, detabs);
    else
      return SKB_EINVAL;
    count_launch();
    return (int)cudaGetLastError();
  }
  if (s.mapping == SKB_MAP_ISO_HEX1) {
    if (!s.mdphi || (x && !s.mphi)) return SKB_EINVAL;
    DeviceFlag flag(st);
    SKB_CUDA_TRY(flag.init());
    int *err = flag.p;
    tabulate_hex_kernel<<<nblk(s.nel * s.nqp, 128), 128, 0, st>>>(s, b, grad, dx, x, detabs, err);
    count_launch();
    int herr = 0;
    SKB_CUDA_TRY(flag.read(&herr));
    if (herr) return SKB_EZERODET;
    return (int)cudaGetLastError();
  }
  return SKB_EINVAL;
}

extern "C" int skb_mapping(const skb_space_t *space, double *DF, double *invDF, double *det,
                           void *stream) {
  using namespace skb;
  if (!space) return SKB_EINVAL;
  const skb_space_t s = *space;
  cudaStream_t st = (cudaStream_t)stream;
  if (s.nel == 0) return SKB_OK;
  if (s.mapping == SKB_MAP_AFFINE) {
    if (s.dim == 2)
      mapping_affine_kernel<2><<<nblk(s.nel, 128), 128, 0, st>>>(s, DF, invDF, det);
    else if (s.dim == 3)
      mapping_affine_kernel<3><<<nblk(s.nel, 128), 128, 0, st>>>(s, DF, invDF, det);
    else
      return SKB_EINVAL;
    count_launch();
    return (int)cudaGetLastError();
  }
  if (s.mapping == SKB_MAP_ISO_HEX1) {
    if (!s.mdphi) return SKB_EINVAL;
    DeviceFlag flag(st);
    SKB_CUDA_TRY(flag.init());
    int *err = flag.p;
    mapping_hex_kernel<<<nblk(s.nel * s.nqp, 128), 128, 0, st>>>(s, DF, invDF, det, err);
    count_launch();
    int herr = 0;
    SKB_CUDA_TRY(flag.read(&herr));
    if (herr) return SKB_EZERODET;
    return (int)cudaGetLastError();
  }
  return SKB_EINVAL;
}

extern "C" int skb_qp_reduce(const double *integrand, const double *dx, int64_t nel, int32_t nqp,
                             int sequential, double *out, void *stream) {
  using namespace skb;
  if (nel < 0 || nqp <= 0) return SKB_EINVAL;
  if (nel == 0) return SKB_OK;
  qp_reduce_kernel<<<nblk(nel, 128), 128, 0, (cudaStream_t)stream>>>(integrand, dx, nel, nqp, sequential, out);
  count_launch();
  return (int)cudaGetLastError();
}

extern "C" int64_t skb_launch_count(int reset) {
  long long v = skb::g_launches.load();
  if (reset) skb::g_launches.store(0);
  return (int64_t)v;
}

extern "C" void skb_debug_flags(int flags) { skb::g_debug.store(flags); }

extern "C" void skb_sm_reserve(int sms) { skb::g_sm_reserve.store(sms < 0 ? 0 : sms); }

extern "C" const char *skb_version(void) { return "skfem_b200 0.1 (sm_100a, fmad=off)"; }

// Boundary conditions on a device-resident CSR system: the
// step right after assembly, so that A and b never have to visit the host.
//
//   skb_csr_enforce    skfem.utils.enforce   (utils.py:327-400): rows D zeroed in
//                      place (entries stay in the pattern), diagonal set to diag
//   skb_csr_condense_* skfem.utils.condense  (utils.py:462-603): A[I][:, I] and
//                      b[I] - A[I][:, D] @ x[D] for sorted index sets; the
//                      matrix-vector part adds a_ij * x_j left to right in column
//                      order from 0.0 like scipy's csr_matvec, so it is bit-identical
//   skb_csr_spmv       y = A x, same order (the hand-off to iterative solvers)
#include "skb_common.cuh"

namespace skb {

__global__ void __launch_bounds__(128)
enforce_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
               double *__restrict__ data, const int32_t *__restrict__ D, int64_t nD, double diag,
               int *__restrict__ missing) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nD;
       k += (int64_t)gridDim.x * blockDim.x) {
    const int32_t r = D[k];
    bool found = false;
    for (int32_t s = indptr[r]; s < indptr[r + 1]; ++s) {
      const bool dg = indices[s] == r;
      data[s] = dg ? diag : 0.0;
      found |= dg;
    }
    if (!found) atomicExch(missing, 1);
  }
}

__global__ void __launch_bounds__(128)
condense_count_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                      const int32_t *__restrict__ I, int64_t nI,
                      const int32_t *__restrict__ colmap, int32_t *__restrict__ counts) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nI;
       k += (int64_t)gridDim.x * blockDim.x) {
    const int32_t r = I[k];
    int32_t c = 0;
    for (int32_t s = indptr[r]; s < indptr[r + 1]; ++s) c += colmap[indices[s]] >= 0;
    counts[k] = c;
  }
}

__global__ void __launch_bounds__(128)
condense_fill_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                     const double *__restrict__ data, const int32_t *__restrict__ I, int64_t nI,
                     const int32_t *__restrict__ colmap, const int32_t *__restrict__ new_indptr,
                     int32_t *__restrict__ new_indices, double *__restrict__ new_data,
                     const double *__restrict__ x, const double *__restrict__ b,
                     double *__restrict__ bout) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nI;
       k += (int64_t)gridDim.x * blockDim.x) {
    const int32_t r = I[k];
    int32_t o = new_indptr[k];
    double y = 0.0;  // (A[I][:, D] @ x[D])[k], scipy csr_matvec order
    for (int32_t s = indptr[r]; s < indptr[r + 1]; ++s) {
      const int32_t c = indices[s];
      const int32_t nc = colmap[c];
      if (nc >= 0) {
        new_indices[o] = nc;
        new_data[o] = data[s];
        ++o;
      } else if (bout) {
        y = y + data[s] * x[c];
      }
    }
    if (bout) bout[k] = b[r] - y;
  }
}

__global__ void __launch_bounds__(128)
spmv_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
            const double *__restrict__ data, const double *__restrict__ x, double *__restrict__ y,
            int64_t nrows) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrows;
       r += (int64_t)gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (int32_t s = indptr[r]; s < indptr[r + 1]; ++s) acc = acc + data[s] * x[indices[s]];
    y[r] = acc;
  }
}

static inline int bc_blocks(int64_t n) {
  int64_t g = (n + 127) / 128;
  if (g < 1) g = 1;
  if (g > 148 * 32) g = 148 * 32;
  return (int)g;
}

}  // namespace skb

extern "C" int skb_csr_enforce(const int32_t *indptr, const int32_t *indices, double *data,
                               const int32_t *D, int64_t nD, double diag, int32_t *missing_diag,
                               void *stream) {
  using namespace skb;
  if (nD < 0) return SKB_EINVAL;
  if (nD == 0) return SKB_OK;
  if (!indptr || !indices || !data || !D || !missing_diag) return SKB_EINVAL;
  enforce_kernel<<<bc_blocks(nD), 128, 0, (cudaStream_t)stream>>>(indptr, indices, data, D, nD,
                                                                  diag, missing_diag);
  count_launch();
  return (int)cudaGetLastError();
}

extern "C" int skb_csr_condense_count(const int32_t *indptr, const int32_t *indices,
                                      const int32_t *I, int64_t nI, const int32_t *colmap,
                                      int32_t *counts, void *stream) {
  using namespace skb;
  if (nI < 0) return SKB_EINVAL;
  if (nI == 0) return SKB_OK;
  if (!indptr || !indices || !I || !colmap || !counts) return SKB_EINVAL;
  condense_count_kernel<<<bc_blocks(nI), 128, 0, (cudaStream_t)stream>>>(indptr, indices, I, nI,
                                                                         colmap, counts);
  count_launch();
  return (int)cudaGetLastError();
}

extern "C" int skb_csr_condense_fill(const int32_t *indptr, const int32_t *indices,
                                     const double *data, const int32_t *I, int64_t nI,
                                     const int32_t *colmap, const int32_t *new_indptr,
                                     int32_t *new_indices, double *new_data, const double *x,
                                     const double *b, double *bout, void *stream) {
  using namespace skb;
  if (nI < 0) return SKB_EINVAL;
  if (nI == 0) return SKB_OK;
  if (!indptr || !indices || !data || !I || !colmap || !new_indptr) return SKB_EINVAL;
  if (bout && (!x || !b)) return SKB_EINVAL;
  condense_fill_kernel<<<bc_blocks(nI), 128, 0, (cudaStream_t)stream>>>(
      indptr, indices, data, I, nI, colmap, new_indptr, new_indices, new_data, x, b, bout);
  count_launch();
  return (int)cudaGetLastError();
}

extern "C" int skb_csr_spmv(const int32_t *indptr, const int32_t *indices, const double *data,
                            const double *x, double *y, int64_t nrows, void *stream) {
  using namespace skb;
  if (nrows < 0) return SKB_EINVAL;
  if (nrows == 0) return SKB_OK;
  if (!indptr || !indices || !data || !x || !y) return SKB_EINVAL;
  spmv_kernel<<<bc_blocks(nrows), 128, 0, (cudaStream_t)stream>>>(indptr, indices, data, x, y,
                                                                  nrows);
  count_launch();
  return (int)cudaGetLastError();
}

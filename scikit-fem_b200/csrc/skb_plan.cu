// Sparsity plan + deterministic numeric phase.
//
// Replaces COOData._assemble_scipy_csr (assembly/form/coo_data.py:27-36), i.e.
// scipy's coo_matrix.eliminate_zeros() followed by tocsr() (coo_tocsr,
// csr_sort_indices, csr_sum_duplicates), and the 1-tensor branch of
// COOData.toarray (coo_data.py:102-108, scipy coo_todense).
//
// The pattern is value dependent (SURVEY finding 1): a CSR slot exists iff at
// least one element-local contribution is != 0.0.  The plan is therefore built
// from the computed local data:
//   keys   : row*ncols+col of every surviving COO triplet (sentinel otherwise)
//   sort   : stable LSD radix sort (cub::DeviceRadixSort) of (key, coo index)
//   unique : head flags + exclusive scan -> slot ids, nnz
//   finalize: indptr / indices / segptr / perm
// The numeric phase is a segmented sum over the precomputed permutation in a
// fixed order (stable COO order: entry-major, element-minor) - no float
// atomics, bit-identical across runs.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include "skb_common.cuh"

namespace skb {

__global__ void make_keys_kernel(const int32_t *__restrict__ dofs_v,
                                 const int32_t *__restrict__ dofs_u, int nbv, int64_t nel,
                                 int64_t ncoo, uint64_t ncols, uint64_t sentinel,
                                 const double *__restrict__ local, int drop_zeros,
                                 uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < ncoo;
       k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ent = k / nel, e = k - ent * nel;
    const int j = (int)(ent / nbv), i = (int)(ent - (int64_t)j * nbv);
    uint64_t key = sentinel;
    const bool keep = !(drop_zeros && local && local[k] == 0.0);  // coo_data.py:35
    if (keep) {
      const uint64_t row = (uint64_t)dofs_v[(int64_t)i * nel + e];
      const uint64_t col = dofs_u ? (uint64_t)dofs_u[(int64_t)j * nel + e] : 0ull;
      key = row * ncols + col;
    }
    keys[k] = key;
    vals[k] = (uint32_t)k;
  }
}

// flag[k] = 1 iff k starts a new (row, col) group among surviving triplets
__global__ void head_flags_kernel(const uint64_t *__restrict__ keys, int64_t ncoo,
                                  uint64_t sentinel, uint32_t *__restrict__ flag) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < ncoo;
       k += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t key = keys[k];
    flag[k] = (key != sentinel && (k == 0 || keys[k - 1] != key)) ? 1u : 0u;
  }
}

// counts[0] = nnz, counts[1] = nkeep
__global__ void counts_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ slot,
                              int64_t ncoo, uint64_t sentinel, unsigned long long *counts) {
  // slot[] holds the inclusive scan of the head flags
  if (blockIdx.x == 0 && threadIdx.x == 0) counts[0] = ncoo ? slot[ncoo - 1] : 0;
  // nkeep = index of the first sentinel: binary search (keys are sorted)
  if (blockIdx.x == 0 && threadIdx.x == 1) {
    int64_t lo = 0, hi = ncoo;
    while (lo < hi) {
      int64_t mid = (lo + hi) >> 1;
      if (keys[mid] < sentinel) lo = mid + 1; else hi = mid;
    }
    counts[1] = (unsigned long long)lo;
  }
}

__global__ void finalize_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals,
                                const uint32_t *__restrict__ slot, int64_t nkeep, int64_t nnz,
                                int64_t nrows, uint64_t ncols, int32_t *__restrict__ indptr,
                                int32_t *__restrict__ indices, uint32_t *__restrict__ segptr,
                                uint32_t *__restrict__ perm) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k <= nkeep;
       k += (int64_t)gridDim.x * blockDim.x) {
    if (k == nkeep) {  // tail: close the last segment and trailing empty rows
      segptr[nnz] = (uint32_t)nkeep;
      int64_t last_row = nkeep ? (int64_t)(keys[nkeep - 1] / ncols) : -1;
      for (int64_t r = last_row + 1; r <= nrows; ++r) indptr[r] = (int32_t)nnz;
      continue;
    }
    perm[k] = vals[k];
    const uint64_t key = keys[k];
    const bool head = (k == 0) || (keys[k - 1] != key);
    if (!head) continue;
    const int64_t s = (int64_t)slot[k] - 1;  // inclusive scan -> slot id
    const int64_t row = (int64_t)(key / ncols);
    indices[s] = (int32_t)(key - (uint64_t)row * ncols);
    segptr[s] = (uint32_t)k;
    // rows (prev_row, row] start at slot s
    const int64_t prev_row = (k == 0) ? -1 : (int64_t)(keys[k - 1] / ncols);
    for (int64_t r = prev_row + 1; r <= row; ++r) indptr[r] = (int32_t)s;
  }
}

// One thread per CSR slot.  The slot's first four entries are fetched side by side (four
// independent perm loads, then four independent value loads) before they are added in plan
// order, so a slot costs three dependent memory round trips instead of 2 n + 1; the sum
// itself is sequential - the order scipy's csr_sum_duplicates uses.
template <class At>
__device__ __forceinline__ void reduce_slots(const uint32_t *__restrict__ perm,
                                             const uint32_t *__restrict__ segptr, int64_t nnz,
                                             double *__restrict__ data, At at) {
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < nnz;
       s += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t a = segptr[s], b = segptr[s + 1], n = b - a;
    const uint32_t p0 = perm[a];
    const uint32_t p1 = n > 1 ? perm[a + 1] : p0, p2 = n > 2 ? perm[a + 2] : p0,
                   p3 = n > 3 ? perm[a + 3] : p0;
    const double v0 = at(p0);
    const double v1 = n > 1 ? at(p1) : 0.0, v2 = n > 2 ? at(p2) : 0.0, v3 = n > 3 ? at(p3) : 0.0;
    double acc = v0;
    if (n > 1) acc = acc + v1;
    if (n > 2) acc = acc + v2;
    if (n > 3) acc = acc + v3;
    for (uint32_t k = a + 4; k < b; ++k) acc = acc + at(perm[k]);
    data[s] = acc;
  }
}

__global__ void csr_reduce_kernel(const double *__restrict__ local, const uint32_t *__restrict__ perm,
                                  const uint32_t *__restrict__ segptr, int64_t nnz,
                                  double *__restrict__ data) {
  reduce_slots(perm, segptr, nnz, data, [&](uint32_t k) { return __ldg(local + k); });
}

// the same sums over element-major local data (nel, Nbv, Nbu): COO entry k = (j Nbv + i) nel + e
// lives at e Nbu Nbv + i Nbu + j, so the entries one CSR row takes from one element are
// consecutive in memory instead of nel doubles apart
__global__ void csr_reduce_em_kernel(const double *__restrict__ local, const uint32_t nel,
                                     const uint32_t nbu, const uint32_t nbv,
                                     const uint32_t *__restrict__ perm,
                                     const uint32_t *__restrict__ segptr, int64_t nnz,
                                     double *__restrict__ data) {
  const uint32_t nb2 = nbu * nbv;
  reduce_slots(perm, segptr, nnz, data, [&](uint32_t k) {
    const uint32_t ji = k / nel, e = k - ji * nel, j = ji / nbv, i = ji - j * nbv;
    return __ldg(local + (size_t)e * nb2 + i * nbu + j);
  });
}

__global__ void vec_reduce_kernel(const double *__restrict__ local, const uint32_t *__restrict__ perm,
                                  const uint32_t *__restrict__ segptr,
                                  const int32_t *__restrict__ indptr, int64_t nrows,
                                  double *__restrict__ vec) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrows;
       r += (int64_t)gridDim.x * blockDim.x) {
    double acc = 0.0;  // coo_todense accumulates into zeros
    if (indptr[r + 1] > indptr[r]) {
      const int32_t s = indptr[r];
      for (uint32_t k = segptr[s]; k < segptr[s + 1]; ++k) acc = acc + __ldg(local + perm[k]);
    }
    vec[r] = acc;
  }
}

static inline int nblocks(int64_t n, int block) {
  int64_t g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > 148 * 32) g = 148 * 32;
  return (int)g;
}

static int key_bits(uint64_t max_key) {
  int b = 1;
  while (b < 64 && (max_key >> b) != 0) ++b;
  return b;
}

}  // namespace skb

extern "C" int64_t skb_plan_scratch_bytes(int64_t ncoo) {
  if (ncoo <= 0) return 256;
  size_t sort_bytes = 0, scan_bytes = 0;
  cub::DoubleBuffer<uint64_t> dk(nullptr, nullptr);
  cub::DoubleBuffer<uint32_t> dv(nullptr, nullptr);
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, dk, dv, (int64_t)ncoo, 0, 64, 0);
  cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, (uint32_t *)nullptr, (uint32_t *)nullptr,
                                (int64_t)ncoo, 0);
  size_t m = sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
  return (int64_t)(m + 256 + 2 * sizeof(unsigned long long));
}

extern "C" int skb_plan_symbolic(const int32_t *dofs_v, const int32_t *dofs_u, int32_t nbv,
                                 int32_t nbu, int64_t nel, int64_t nrows, int64_t ncols,
                                 const double *local_or_null, int drop_zeros, uint64_t *keys_a,
                                 uint64_t *keys_b, uint32_t *vals_a, uint32_t *vals_b,
                                 uint32_t *slot, void *tmp, int64_t tmp_bytes,
                                 int64_t *counts_host, void *stream) {
  using namespace skb;
  cudaStream_t st = (cudaStream_t)stream;
  if (!dofs_v || nbv <= 0 || nbu <= 0 || nel < 0 || nrows <= 0 || ncols <= 0 || !counts_host)
    return SKB_EINVAL;
  const int64_t ncoo = (int64_t)nbv * nbu * nel;
  if (ncoo >= (int64_t)0xffffffffLL) return SKB_ETOOBIG;  // perm/segptr are uint32
  counts_host[0] = counts_host[1] = 0;
  counts_host[2] = 0;
  if (ncoo == 0) return SKB_OK;
  if (tmp_bytes < skb_plan_scratch_bytes(ncoo)) return SKB_EINVAL;
  const uint64_t sentinel = (uint64_t)nrows * (uint64_t)ncols;
  make_keys_kernel<<<nblocks(ncoo, 256), 256, 0, st>>>(dofs_v, dofs_u, nbv, nel, ncoo,
                                                       (uint64_t)ncols, sentinel, local_or_null,
                                                       drop_zeros, keys_a, vals_a);
  SKB_CUDA_TRY(cudaGetLastError());
  count_launch(3);  // make_keys, head_flags, counts (CUB sort/scan passes not counted)
  unsigned long long *counts_dev = (unsigned long long *)tmp;
  void *cub_tmp = (char *)tmp + 256;
  size_t cub_bytes = (size_t)tmp_bytes - 256;
  cub::DoubleBuffer<uint64_t> dk(keys_a, keys_b);
  cub::DoubleBuffer<uint32_t> dv(vals_a, vals_b);
  SKB_CUDA_TRY(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, dk, dv, ncoo, 0,
                                               key_bits(sentinel), st));
  // tell the caller which buffer holds the sorted result (0 = a, 1 = b)
  counts_host[2] = (dk.Current() == keys_a) ? 0 : 1;
  if ((dk.Current() == keys_a) != (dv.Current() == vals_a)) return SKB_EINVAL;
  uint64_t *ks = dk.Current();
  head_flags_kernel<<<nblocks(ncoo, 256), 256, 0, st>>>(ks, ncoo, sentinel, slot);
  SKB_CUDA_TRY(cudaGetLastError());
  cub_bytes = (size_t)tmp_bytes - 256;
  SKB_CUDA_TRY(cub::DeviceScan::InclusiveSum(cub_tmp, cub_bytes, slot, slot, ncoo, st));
  counts_kernel<<<1, 32, 0, st>>>(ks, slot, ncoo, sentinel, counts_dev);
  SKB_CUDA_TRY(cudaGetLastError());
  unsigned long long h[2] = {0, 0};
  SKB_CUDA_TRY(cudaMemcpyAsync(h, counts_dev, sizeof(h), cudaMemcpyDeviceToHost, st));
  SKB_CUDA_TRY(cudaStreamSynchronize(st));
  counts_host[0] = (int64_t)h[0];
  counts_host[1] = (int64_t)h[1];
  return SKB_OK;
}

extern "C" int skb_plan_finalize(int64_t ncoo, int64_t nrows, int64_t ncols, int64_t nnz,
                                 int64_t nkeep, const uint64_t *keys_sorted,
                                 const uint32_t *vals_sorted, const uint32_t *slot,
                                 int32_t *indptr, int32_t *indices, uint32_t *segptr,
                                 uint32_t *perm, void *stream) {
  using namespace skb;
  cudaStream_t st = (cudaStream_t)stream;
  if (nrows <= 0 || ncols <= 0 || nnz < 0 || nkeep < 0 || nkeep > ncoo || !indptr || !segptr)
    return SKB_EINVAL;
  if (nnz >= (int64_t)0x7fffffffLL) return SKB_ETOOBIG;  // int32 indptr (scipy would use int64)
  finalize_kernel<<<nblocks(nkeep + 1, 256), 256, 0, st>>>(keys_sorted, vals_sorted, slot, nkeep,
                                                           nnz, nrows, (uint64_t)ncols, indptr,
                                                           indices, segptr, perm);
  count_launch();
  return (int)cudaGetLastError();
}

extern "C" int skb_csr_reduce(const double *local, const uint32_t *perm, const uint32_t *segptr,
                              int64_t nnz, double *data, void *stream) {
  using namespace skb;
  if (nnz < 0) return SKB_EINVAL;
  if (nnz == 0) return SKB_OK;
  csr_reduce_kernel<<<nblocks(nnz, 256), 256, 0, (cudaStream_t)stream>>>(local, perm, segptr, nnz,
                                                                         data);
  count_launch();
  return (int)cudaGetLastError();
}

extern "C" int skb_csr_reduce_em(const double *local_em, int64_t nel, int32_t nbu, int32_t nbv,
                                 const uint32_t *perm, const uint32_t *segptr, int64_t nnz,
                                 double *data, void *stream) {
  using namespace skb;
  if (nnz < 0 || nel < 0 || nbu <= 0 || nbv <= 0 || (int64_t)nbu * nbv * nel >= (1ll << 32))
    return SKB_EINVAL;
  if (nnz == 0) return SKB_OK;
  if (nel == 0) return SKB_EINVAL;           // slots without entries cannot exist
  csr_reduce_em_kernel<<<nblocks(nnz, 256), 256, 0, (cudaStream_t)stream>>>(
      local_em, (uint32_t)nel, (uint32_t)nbu, (uint32_t)nbv, perm, segptr, nnz, data);
  count_launch();
  return (int)cudaGetLastError();
}

extern "C" int skb_vec_reduce(const double *local, const uint32_t *perm, const uint32_t *segptr,
                              const int32_t *indptr, int64_t nrows, double *vec, void *stream) {
  using namespace skb;
  if (nrows < 0) return SKB_EINVAL;
  if (nrows == 0) return SKB_OK;
  vec_reduce_kernel<<<nblocks(nrows, 256), 256, 0, (cudaStream_t)stream>>>(local, perm, segptr,
                                                                           indptr, nrows, vec);
  count_launch();
  return (int)cudaGetLastError();
}

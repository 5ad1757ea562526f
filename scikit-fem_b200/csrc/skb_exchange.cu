// Interface-row exchange of the multi-GPU path: the two device-side
// steps around the NCCL all-to-all-v.  The reference's analogue is PETSc's MATIS -> mpiaij
// conversion, which adds the interface rows across ranks (skfem/assembly/form/coo_data.py:
// 151-170); here every rank packs the values of the CSR slots whose row a peer owns, one
// collective moves them, and the owner adds each received segment into its row block.
#include "skb_common.cuh"

namespace skb {

__global__ void __launch_bounds__(256)
interface_pack_kernel(const double *__restrict__ vals, const int64_t *__restrict__ slots,
                      int64_t n, double *__restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = vals[slots[i]];
}

// data[pos[i]] += recv[i]; the targets of one segment (one source rank) are distinct, and the
// caller launches the segments in source-rank order on one stream: deterministic, no atomics
__global__ void __launch_bounds__(256)
interface_add_kernel(double *__restrict__ data, const int64_t *__restrict__ pos,
                     const double *__restrict__ recv, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = pos[i];
    data[s] = data[s] + recv[i];
  }
}

static int grid_for(int64_t n) {
  int64_t g = (n + 255) / 256;
  return (int)(g > 148 * 8 ? 148 * 8 : g);
}

}  // namespace skb

extern "C" int skb_pack_interface(const double *vals, const int64_t *slots, int64_t n,
                                  double *out, void *stream) {
  using namespace skb;
  if (n < 0) return SKB_EINVAL;
  if (n == 0) return SKB_OK;
  if (!vals || !slots || !out) return SKB_EINVAL;
  interface_pack_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(vals, slots, n, out);
  count_launch();
  return (int)cudaGetLastError();
}

extern "C" int skb_unpack_add_interface(double *data, const int64_t *pos, const double *recv,
                                        int64_t n, void *stream) {
  using namespace skb;
  if (n < 0) return SKB_EINVAL;
  if (n == 0) return SKB_OK;
  if (!data || !pos || !recv) return SKB_EINVAL;
  interface_add_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(data, pos, recv, n);
  count_launch();
  return (int)cudaGetLastError();
}

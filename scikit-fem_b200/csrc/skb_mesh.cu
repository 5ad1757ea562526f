// Tensor-product meshes generated on the device.
//
// Replaces MeshTet.init_tensor / MeshHex.init_tensor (mesh/mesh_tet_1.py:326-393,
// mesh/mesh_hex_1.py:97-155) for large grids: vertex v = iy + npy*ix + npy*npx*iz at
// (x[ix], y[iy], z[iz]) (np.meshgrid + flatten('F')); cell c = iy + (npy-1)*ix +
// (npy-1)*(npx-1)*iz with its eight corners base + {0, sy, sx, sz, sy+sx, sy+sz, sx+sz,
// sy+sx+sz} (sy = 1, sx = npy, sz = npy*npx); element `ty * ncells + c` takes the corners
// tab[ty][0..nnodes) - the six Kuhn tetrahedra around the diagonal 0 -> 7 (type-major element
// order), or the hexahedron itself.  Integer work only: p is a copy of the inputs, so the
// arrays are bit for bit the host generator's.
#include "skb_common.cuh"

namespace skb {

struct MeshTab { int32_t corner[6][8]; };

__global__ void mesh_tensor_points_kernel(const double *__restrict__ x,
                                          const double *__restrict__ y,
                                          const double *__restrict__ z, int npx, int npy,
                                          int64_t npts, double *__restrict__ p) {
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < npts;
       v += (int64_t)gridDim.x * blockDim.x) {
    const int64_t iz = v / ((int64_t)npy * npx), r = v - iz * (int64_t)npy * npx;
    const int64_t ix = r / npy, iy = r - ix * npy;
    p[v] = x[ix];
    p[npts + v] = y[iy];
    p[2 * npts + v] = z[iz];
  }
}

__global__ void mesh_tensor_cells_kernel(int npx, int npy, int npz, int ntypes, int nnodes,
                                         MeshTab tab, int32_t *__restrict__ t) {
  const int64_t cy = npy - 1, cx = npx - 1, ncells = cy * cx * (int64_t)(npz - 1);
  const int64_t nel = ncells * ntypes;
  const int64_t off[8] = {0, 1, npy, (int64_t)npy * npx, 1 + npy, 1 + (int64_t)npy * npx,
                          npy + (int64_t)npy * npx, 1 + npy + (int64_t)npy * npx};
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < ncells;
       c += (int64_t)gridDim.x * blockDim.x) {
    const int64_t iz = c / (cy * cx), r = c - iz * cy * cx;
    const int64_t ix = r / cy, iy = r - ix * cy;
    const int64_t base = iy + (int64_t)npy * ix + (int64_t)npy * npx * iz;
    for (int ty = 0; ty < ntypes; ++ty)
      for (int n = 0; n < nnodes; ++n)
        t[(int64_t)n * nel + (int64_t)ty * ncells + c] = (int32_t)(base + off[tab.corner[ty][n]]);
  }
}

// element_dofs rows (assembly/dofs.py:264-334): every row is  add + mul * src[e]  with src a
// row of t / t2e / t2f (entity-major blocks: dof c of entity n is offset + c + nd * n) or the
// element index itself (interior DOFs, src == NULL).
struct DofRow { const int32_t *src; int32_t mul, add; };
__global__ void element_dofs_kernel(const DofRow *__restrict__ rows, int nrows, int64_t nel,
                                    int32_t *__restrict__ out) {
  const int64_t n = (int64_t)nrows * nel;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / nel, e = idx - r * nel;
    const DofRow d = rows[r];
    const int32_t v = d.src ? d.src[e] : (int32_t)e;
    out[idx] = d.add + d.mul * v;
  }
}

}  // namespace skb

// rows_dev: nrows device records {int64 src pointer (0 = element index), int32 mul, int32 add};
// out int32[nrows][nel].
extern "C" int skb_element_dofs(const void *rows_dev, int32_t nrows, int64_t nel, int32_t *out,
                                void *stream) {
  using namespace skb;
  if (!rows_dev || !out || nrows <= 0 || nel < 0) return SKB_EINVAL;
  if (nel == 0) return SKB_OK;
  int64_t g = ((int64_t)nrows * nel + 255) / 256;
  if (g > 148 * 32) g = 148 * 32;
  element_dofs_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const DofRow *>(rows_dev), nrows, nel, out);
  count_launch();
  return (int)cudaGetLastError();
}

// x, y, z: sorted coordinates on the device (npx, npy, npz doubles); corner_host: ntypes x
// nnodes corner indices (0..7); p: double[3][npx*npy*npz], t: int32[nnodes][ntypes*ncells].
extern "C" int skb_mesh_tensor(const double *x, const double *y, const double *z, int32_t npx,
                               int32_t npy, int32_t npz, int32_t ntypes, int32_t nnodes,
                               const int32_t *corner_host, double *p, int32_t *t, void *stream) {
  using namespace skb;
  if (!x || !y || !z || !p || !t || !corner_host || npx < 2 || npy < 2 || npz < 2 ||
      ntypes < 1 || ntypes > 6 || nnodes < 1 || nnodes > 8)
    return SKB_EINVAL;
  const int64_t npts = (int64_t)npx * npy * npz;
  const int64_t ncells = (int64_t)(npx - 1) * (npy - 1) * (npz - 1);
  if (npts >= (int64_t)0x7fffffffLL || ncells * ntypes >= (int64_t)0x7fffffffLL)
    return SKB_ETOOBIG;                                  // int32 connectivity
  MeshTab tab;
  for (int ty = 0; ty < 6; ++ty)
    for (int n = 0; n < 8; ++n) {
      const int v = (ty < ntypes && n < nnodes) ? corner_host[ty * nnodes + n] : 0;
      if (v < 0 || v > 7) return SKB_EINVAL;
      tab.corner[ty][n] = v;
    }
  cudaStream_t st = (cudaStream_t)stream;
  auto blocks = [](int64_t n) {
    int64_t g = (n + 255) / 256;
    return (int)(g < 1 ? 1 : (g > 148 * 32 ? 148 * 32 : g));
  };
  mesh_tensor_points_kernel<<<blocks(npts), 256, 0, st>>>(x, y, z, npx, npy, npts, p);
  SKB_CUDA_TRY(cudaGetLastError());
  mesh_tensor_cells_kernel<<<blocks(ncells), 256, 0, st>>>(npx, npy, npz, ntypes, nnodes, tab, t);
  count_launch(2);
  return (int)cudaGetLastError();
}

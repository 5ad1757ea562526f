// Sparsity plan without a global sort: row buckets + per-row sorts.
//
// Same contract as skb_plan_symbolic / skb_plan_finalize (skb_plan.cu), i.e. the structure
// of COOData._assemble_scipy_csr (assembly/form/coo_data.py:27-36: eliminate_zeros + tocsr):
// indptr / indices of the value-dependent pattern and, per CSR slot, the surviving COO
// entries in stable COO order (perm / segptr) - bit for bit the arrays the radix-sort path
// produces (tests/test_gpu_plan_rows.py), but for a fraction of its memory traffic: the
// radix sort moves 12 bytes per COO entry five times in each direction, this path touches
// every entry once.
//
// The (row, col, k) triplets of a finite element COO list are not arbitrary: row r only
// receives entries from the elements that contain DOF r.  So
//   count   one thread per incidence (i, e) - row basis function i of element e: the mask of
//           its surviving columns (local value != 0.0, coo_data.py:35), the row's incidence
//           and candidate counters (one packed 64-bit atomic)
//   scan    exclusive scans of both counters over the rows (own kernels below)
//   fill    incidence lists: inc[incstart[r] ...] = {i * nel + e, mask}, order arbitrary (atomic
//           cursor); the column DOFs are transposed to element-major so that a row reads the
//           columns of one incident element in one piece
//   sort    one warp per row: candidates (col, k) from the row's incidences and masks, bitonic
//           sort in registers (up to 128 surviving entries; shared memory up to 512; one CTA
//           per row up to 8192), unique columns; writes the row's stretch of perm directly -
//           it starts at the row's candidate offset, because perm is ordered by (row, col, k)
//           - and the row's unique columns / segment starts into two scratch arrays
//   emit    after the scan of the unique counts (= indptr): indices and segptr
// Every array written is a pure function of the inputs (the sort erases the atomics' order).
#include "skb_common.cuh"

namespace skb {

constexpr int ROWS_SHORT_CAP = 512;     // surviving entries per row handled by one warp
constexpr int ROWS_LONG_CAP = 8192;     // ... by one CTA; longer rows: radix-sort path
constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 4;                       // per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// ---- exclusive scan of f(in[i]), i < n, into out[0..n] (out[n] = total), two levels -------
struct LoadHi { __device__ uint32_t operator()(unsigned long long v) const { return (uint32_t)(v >> 32); } };
struct LoadLo { __device__ uint32_t operator()(unsigned long long v) const { return (uint32_t)v; } };
struct LoadId { __device__ uint32_t operator()(uint32_t v) const { return v; } };

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total, uint32_t *sh) {
  // sh: [32] warp sums
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= d) x += y;
  }
  if (lane == 31) sh[warp] = x;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < (int)(blockDim.x >> 5) ? sh[lane] : 0u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w += y;
    }
    sh[lane] = w;                                    // inclusive warp-sum scan
  }
  __syncthreads();
  const uint32_t base = warp ? sh[warp - 1] : 0u;
  *total = sh[(blockDim.x >> 5) - 1];
  return base + x - v;
}

template <class In, class F>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_tile_sums_kernel(const In *__restrict__ in, int64_t n, F f, uint32_t *__restrict__ sums) {
  __shared__ uint32_t sh[32];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k)
    if (base + k < n) s += f(in[base + k]);
  uint32_t total;
  block_exclusive_scan(s, &total, sh);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// one CTA: exclusive scan of the tile sums in place (ntiles <= SCAN_TILE)
__global__ void __launch_bounds__(SCAN_THREADS)
scan_sums_kernel(uint32_t *__restrict__ sums, int ntiles) {
  __shared__ uint32_t sh[32];
  uint32_t v[SCAN_ITEMS], s = 0;
  const int base = threadIdx.x * SCAN_ITEMS;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    v[k] = base + k < ntiles ? sums[base + k] : 0u;
    s += v[k];
  }
  uint32_t total;
  uint32_t run = block_exclusive_scan(s, &total, sh);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (base + k < ntiles) sums[base + k] = run;
    run += v[k];
  }
}

template <class In, class F, class Out>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_apply_kernel(const In *__restrict__ in, int64_t n, F f, const uint32_t *__restrict__ sums,
                  Out *__restrict__ out) {
  __shared__ uint32_t sh[32];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    v[k] = base + k < n ? f(in[base + k]) : 0u;
    s += v[k];
  }
  uint32_t total;
  uint32_t run = sums[blockIdx.x] + block_exclusive_scan(s, &total, sh);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (base + k < n) out[base + k] = (Out)run;
    run += v[k];
  }
  // the grand total goes behind the last element
  if (base <= n - 1 && n - 1 < base + SCAN_ITEMS) out[n] = (Out)run;
}

template <class In, class F, class Out>
static int exclusive_scan(const In *in, int64_t n, F f, uint32_t *sums, Out *out, cudaStream_t st) {
  const int64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (ntiles > SCAN_TILE) return SKB_ETOOBIG;
  scan_tile_sums_kernel<<<(unsigned)ntiles, SCAN_THREADS, 0, st>>>(in, n, f, sums);
  scan_sums_kernel<<<1, SCAN_THREADS, 0, st>>>(sums, (int)ntiles);
  scan_apply_kernel<<<(unsigned)ntiles, SCAN_THREADS, 0, st>>>(in, n, f, sums, out);
  count_launch(3);
  return (int)cudaGetLastError();
}

// ---- count ---------------------------------------------------------------------------------
__global__ void rows_count_kernel(const int32_t *__restrict__ dofs_v, int nbv, int nbu,
                                  int64_t nel, const double *__restrict__ local, int drop_zeros,
                                  uint32_t *__restrict__ mask,
                                  unsigned long long *__restrict__ rc) {
  const int64_t ninc = (int64_t)nbv * nel;
  const uint32_t full = nbu >= 32 ? 0xffffffffu : ((1u << nbu) - 1u);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < ninc;
       idx += (int64_t)gridDim.x * blockDim.x) {
    uint32_t m = full;
    if (drop_zeros == 2) {
      m = mask[idx] & full;                              // masks supplied by the caller
    } else if (drop_zeros && local) {
      const int64_t i = idx / nel, e = idx - i * nel;
      m = 0;
      for (int j = 0; j < nbu; ++j)                      // coo_data.py:35
        m |= (uint32_t)(local[((int64_t)j * nbv + i) * nel + e] != 0.0) << j;
    }
    mask[idx] = m;
    if (m) atomicAdd(&rc[dofs_v[idx]], (1ull << 32) | (unsigned long long)__popc(m));
  }
}

// ---- fill ----------------------------------------------------------------------------------
// incidence record: {i * nel + e, mask of the surviving columns}
__global__ void rows_fill_kernel(const int32_t *__restrict__ dofs_v, int64_t ninc,
                                 const uint32_t *__restrict__ mask,
                                 const uint32_t *__restrict__ incstart,
                                 uint32_t *__restrict__ cursor, uint2 *__restrict__ inc) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < ninc;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t m = mask[idx];
    if (!m) continue;
    const int32_t r = dofs_v[idx];
    inc[incstart[r] + atomicAdd(&cursor[r], 1u)] = make_uint2((uint32_t)idx, m);
  }
}

// column DOFs element-major, (nel, nbu): the columns of one element are one contiguous read
__global__ void rows_transpose_kernel(const int32_t *__restrict__ dofs_u, int nbu, int64_t nel,
                                      int32_t *__restrict__ out) {
  const int64_t n = (int64_t)nbu * nel;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n;
       o += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = o / nbu;
    const int j = (int)(o - e * nbu);
    out[o] = dofs_u[(int64_t)j * nel + e];
  }
}

// ---- sort ----------------------------------------------------------------------------------
// Rows too long for the register sort (rows_sort_warp_kernel below files them in two lists):
// bitonic sort in shared memory.  THREADS == 32: one warp per row, WARPS rows per CTA (up to
// CAP = 512 entries); else one CTA per row (up to CAP = 8192).  Longer rows raise *too_long.
template <int THREADS, int WARPS, int CAP>
__global__ void __launch_bounds__(THREADS * WARPS)
rows_sort_kernel(const int32_t *__restrict__ dofs_ut, int nbv, int nbu, int64_t nel,
                 const uint32_t *__restrict__ list, const int32_t *__restrict__ nlist,
                 const uint32_t *__restrict__ incstart, const uint32_t *__restrict__ candstart,
                 const uint2 *__restrict__ inc,
                 uint32_t *__restrict__ perm, uint32_t *__restrict__ ucol,
                 uint32_t *__restrict__ uoff, uint32_t *__restrict__ nuniq,
                 int32_t *__restrict__ too_long) {
  extern __shared__ unsigned long long rows_smem[];
  const int sub = THREADS == 32 ? (int)(threadIdx.x >> 5) : 0;
  const int tid = THREADS == 32 ? (int)(threadIdx.x & 31) : (int)threadIdx.x;
  unsigned long long *buf = rows_smem + (size_t)sub * CAP;
  __shared__ uint32_t s_off[WARPS][THREADS == 32 ? 1 : 33];
  auto sync = [&]() {
    if (THREADS == 32) __syncwarp(); else __syncthreads();
  };
  const int64_t l0 = THREADS == 32 ? (int64_t)blockIdx.x * WARPS + sub : (int64_t)blockIdx.x;
  const int64_t lstride = THREADS == 32 ? (int64_t)gridDim.x * WARPS : (int64_t)gridDim.x;
  const int64_t nl = *nlist;
  for (int64_t li = l0; li < nl; li += lstride) {
    const int64_t r = list[li];
    const uint32_t c0 = candstart[r], nc = candstart[r + 1] - c0;
    if (nc > (uint32_t)CAP) {
      if (tid == 0) { *too_long = 1; nuniq[r] = 0; }
      continue;
    }
    const uint32_t i0 = incstart[r], ni = incstart[r + 1] - i0;
    int n2 = 32;
    while (n2 < (int)nc) n2 <<= 1;
    // candidates -> buf: incidences in chunks of THREADS, offsets by a running scan
    uint32_t run = 0;
    for (uint32_t q0 = 0; q0 < ni; q0 += THREADS) {
      const uint32_t q = q0 + tid;
      uint32_t idx = 0, m = 0;
      if (q < ni) {
        const uint2 rec = inc[i0 + q];
        idx = rec.x;
        m = rec.y;
      }
      const uint32_t cnt = __popc(m);
      // exclusive scan of cnt over the chunk
      uint32_t x = cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
        if ((tid & 31) >= d) x += y;
      }
      uint32_t off = run + x - cnt, chunk_total;
      if (THREADS == 32) {
        chunk_total = __shfl_sync(0xffffffffu, x, 31);
      } else {
        const int w = tid >> 5;
        if ((tid & 31) == 31) s_off[0][w + 1] = x;
        if (tid == 0) s_off[0][0] = 0;
        __syncthreads();
        if (tid == 0)
          for (int k = 1; k <= THREADS / 32; ++k) s_off[0][k] += s_off[0][k - 1];
        __syncthreads();
        off += s_off[0][w];
        chunk_total = s_off[0][THREADS / 32];
        __syncthreads();
      }
      if (m) {
        const int64_t i = idx / nel, e = idx - i * nel;
        while (m) {
          const int j = __ffs(m) - 1;
          m &= m - 1;
          const uint32_t col = dofs_ut ? (uint32_t)dofs_ut[e * nbu + j] : 0u;
          const unsigned long long k = (unsigned long long)(((int64_t)j * nbv + i) * nel + e);
          buf[off++] = ((unsigned long long)col << 32) | k;
        }
      }
      run += chunk_total;
    }
    for (int x = (int)nc + tid; x < n2; x += THREADS) buf[x] = ~0ull;
    sync();
    // bitonic sort of buf[0..n2)
    for (int k2 = 2; k2 <= n2; k2 <<= 1)
      for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
        for (int x = tid; x < n2; x += THREADS) {
          const int y = x ^ j2;
          if (y > x) {
            const unsigned long long a = buf[x], b = buf[y];
            const bool up = (x & k2) == 0;
            if ((a > b) == up) { buf[x] = b; buf[y] = a; }
          }
        }
        sync();
      }
    // perm, unique columns and their segment starts (heads compacted by a running count)
    uint32_t urun = 0;
    for (uint32_t x0 = 0; x0 < nc; x0 += THREADS) {
      const uint32_t x = x0 + tid;
      bool head = false;
      uint32_t col = 0;
      if (x < nc) {
        const unsigned long long v = buf[x];
        col = (uint32_t)(v >> 32);
        perm[c0 + x] = (uint32_t)v;
        head = x == 0 || (uint32_t)(buf[x - 1] >> 32) != col;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, head);
      uint32_t pos = urun + __popc(bal & ((1u << (tid & 31)) - 1u));
      uint32_t chunk_total = __popc(bal);
      if (THREADS != 32) {
        const int w = tid >> 5;
        if ((tid & 31) == 0) s_off[0][w + 1] = chunk_total;
        if (tid == 0) s_off[0][0] = 0;
        __syncthreads();
        if (tid == 0)
          for (int k = 1; k <= THREADS / 32; ++k) s_off[0][k] += s_off[0][k - 1];
        __syncthreads();
        pos += s_off[0][w];
        chunk_total = s_off[0][THREADS / 32];
        __syncthreads();
      }
      if (head) {
        ucol[c0 + pos] = col;
        uoff[c0 + pos] = c0 + x;
      }
      urun += chunk_total;
    }
    if (tid == 0) nuniq[r] = urun;
    sync();
  }
}

// ---- sort, short rows: one warp per row, keys in registers ----------------------------------
// K keys per lane (element p = 32 s + lane), bitonic network with compile-time stages: partners
// 32 or more apart live in the same lane, closer ones come by shuffle.
template <int K>
__device__ __forceinline__ void warp_bitonic(unsigned long long (&key)[K], int lane) {
#pragma unroll
  for (int k2 = 2; k2 <= 32 * K; k2 <<= 1) {
#pragma unroll
    for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
      if (j2 >= 32) {
#pragma unroll
        for (int s = 0; s < K; ++s) {
          const int t = s ^ (j2 >> 5);
          if (t > s) {
            const bool up = ((32 * s) & k2) == 0;          // lane bits are below k2 here
            const unsigned long long mn = key[s] < key[t] ? key[s] : key[t];
            const unsigned long long mx = key[s] < key[t] ? key[t] : key[s];
            key[s] = up ? mn : mx;
            key[t] = up ? mx : mn;
          }
        }
      } else {
#pragma unroll
        for (int s = 0; s < K; ++s) {
          const unsigned long long pk = __shfl_xor_sync(0xffffffffu, key[s], j2);
          const bool up = (((32 * s) | lane) & k2) == 0;
          const bool lower = (lane & j2) == 0;
          // the lower partner keeps the minimum in an ascending run, the maximum otherwise
          const unsigned long long mn = key[s] < pk ? key[s] : pk;
          const unsigned long long mx = key[s] < pk ? pk : key[s];
          key[s] = (lower == up) ? mn : mx;
        }
      }
    }
  }
}

template <int K>
__device__ __forceinline__ void warp_row_sorted_out(const unsigned long long (&key)[K], int lane,
                                                    uint32_t nc, uint32_t c0, int64_t r,
                                                    uint32_t *__restrict__ perm,
                                                    uint32_t *__restrict__ ucol,
                                                    uint32_t *__restrict__ uoff,
                                                    uint32_t *__restrict__ nuniq) {
  uint32_t urun = 0, prev_last = 0xffffffffu;              // column of element 32 s - 1
#pragma unroll
  for (int s = 0; s < K; ++s) {
    const uint32_t x = 32u * s + lane;
    const uint32_t col = (uint32_t)(key[s] >> 32);
    uint32_t pcol = __shfl_up_sync(0xffffffffu, col, 1);
    if (lane == 0) pcol = prev_last;
    const bool in = x < nc;
    const bool head = in & ((x == 0) | (pcol != col));
    if (in) perm[c0 + x] = (uint32_t)key[s];
    const unsigned bal = __ballot_sync(0xffffffffu, head);
    if (head) {
      const uint32_t pos = urun + __popc(bal & ((1u << lane) - 1u));
      ucol[c0 + pos] = col;
      uoff[c0 + pos] = c0 + x;
    }
    urun += __popc(bal);
    prev_last = __shfl_sync(0xffffffffu, col, 31);
  }
  if (lane == 0) nuniq[r] = urun;
}

// Rows with up to 128 surviving entries are finished here; longer ones are filed in `mid`
// (up to 512: warp-wide shared-memory sort) or `lng` (CTA-wide).  cnt[0], cnt[1]: list lengths.
template <int WARPS>
__global__ void __launch_bounds__(32 * WARPS)
rows_sort_warp_kernel(const int32_t *__restrict__ dofs_ut, int nbv, int nbu, int64_t nel,
                      int64_t nrows, const uint32_t *__restrict__ incstart,
                      const uint32_t *__restrict__ candstart, const uint2 *__restrict__ inc,
                      uint32_t *__restrict__ perm, uint32_t *__restrict__ ucol,
                      uint32_t *__restrict__ uoff, uint32_t *__restrict__ nuniq,
                      uint32_t *__restrict__ mid, uint32_t *__restrict__ lng,
                      int32_t *__restrict__ cnt) {
  __shared__ unsigned long long s_key[WARPS][128];
  const int lane = threadIdx.x & 31, sub = threadIdx.x >> 5;
  unsigned long long *bk = s_key[sub];
  for (int64_t r = (int64_t)blockIdx.x * WARPS + sub; r < nrows; r += (int64_t)gridDim.x * WARPS) {
    const uint32_t c0 = candstart[r], nc = candstart[r + 1] - c0;
    if (nc == 0) {
      if (lane == 0) nuniq[r] = 0;
      continue;
    }
    if (nc > 128u) {
      if (lane == 0) {
        if (nc <= (uint32_t)ROWS_SHORT_CAP) mid[atomicAdd(&cnt[0], 1)] = (uint32_t)r;
        else lng[atomicAdd(&cnt[1], 1)] = (uint32_t)r;
      }
      continue;
    }
    // candidates: one lane per (incidence, column) pair, survivors compacted into bh / bl
    const uint32_t i0 = incstart[r], ni = incstart[r + 1] - i0;
    const uint32_t npair = ni * (uint32_t)nbu;
    uint32_t run = 0;
    for (uint32_t c = lane; c - lane < npair; c += 32) {
      bool keep = false;
      uint32_t col = 0, k = 0;
      if (c < npair) {
        const uint32_t q = c / (uint32_t)nbu, j = c - q * (uint32_t)nbu;
        const uint2 rec = inc[i0 + q];
        keep = (rec.y >> j) & 1u;
        if (keep) {
          const uint32_t i = rec.x / (uint32_t)nel;       // nbv * nel < 2^32 (checked by count)
          const uint32_t e = rec.x - i * (uint32_t)nel;
          col = dofs_ut ? (uint32_t)dofs_ut[(int64_t)e * nbu + j] : 0u;
          k = (uint32_t)(((int64_t)j * nbv + i) * nel + e);
        }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, keep);
      if (keep) bk[run + __popc(bal & ((1u << lane) - 1u))] = ((unsigned long long)col << 32) | k;
      run += __popc(bal);
    }
    __syncwarp();
    if (nc <= 64u) {
      unsigned long long key[2];
#pragma unroll
      for (int s = 0; s < 2; ++s) key[s] = 32u * s + lane < nc ? bk[32 * s + lane] : ~0ull;
      warp_bitonic<2>(key, lane);
      warp_row_sorted_out<2>(key, lane, nc, c0, r, perm, ucol, uoff, nuniq);
    } else {
      unsigned long long key[4];
#pragma unroll
      for (int s = 0; s < 4; ++s) key[s] = 32u * s + lane < nc ? bk[32 * s + lane] : ~0ull;
      warp_bitonic<4>(key, lane);
      warp_row_sorted_out<4>(key, lane, nc, c0, r, perm, ucol, uoff, nuniq);
    }
    __syncwarp();
  }
}

// ---- emit ----------------------------------------------------------------------------------
__global__ void rows_emit_kernel(int64_t nrows, const uint32_t *__restrict__ candstart,
                                 const int32_t *__restrict__ indptr,
                                 const uint32_t *__restrict__ ucol,
                                 const uint32_t *__restrict__ uoff,
                                 int32_t *__restrict__ indices, uint32_t *__restrict__ segptr,
                                 int64_t nnz, int64_t nkeep) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = w0; r < nrows; r += nw) {
    const int32_t s0 = indptr[r], nu = indptr[r + 1] - s0;
    const uint32_t c0 = candstart[r];
    for (int u = lane; u < nu; u += 32) {
      indices[s0 + u] = (int32_t)ucol[c0 + u];
      segptr[s0 + u] = uoff[c0 + u];
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) segptr[nnz] = (uint32_t)nkeep;
}

// ---- mesh entities through the same machinery ----------------------------------------------
// Mesh.build_entities (mesh/mesh.py:1065-1082: np.sort + np.unique(axis=1) of the vertex tuples
// of all local edges / facets) is the same problem as the plan: the unique sorted pairs
// (row = smallest vertex, col = other vertex) of an edge list are the CSR pattern of the
// "matrix" whose incidences are (local vertex i, element e) and whose surviving columns are
// the local vertices j adjacent to i with a larger global index.  Triangular facets: row = id
// of the edge of the two smallest vertices, col = largest vertex.  This kernel forms those
// masks: bit j of mask[i * nel + e] <=> j in adj[i] and tu[j][e] > vmax[i][e].
struct EntityAdj { uint32_t adj[32]; };
__global__ void entity_mask_kernel(const int32_t *__restrict__ tu, int nbu, int nbv, int64_t nel,
                                   const int32_t *__restrict__ vmax, EntityAdj a,
                                   uint32_t *__restrict__ mask) {
  const int64_t n = (int64_t)nbv * nel;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / nel, e = idx - i * nel;
    const int32_t vm = vmax[idx];
    uint32_t m = 0, cand = a.adj[i];
    while (cand) {
      const int j = __ffs(cand) - 1;
      cand &= cand - 1;
      if (j < nbu && tu[(int64_t)j * nel + e] > vm) m |= 1u << j;
    }
    mask[idx] = m;
  }
}

// slot[k] = CSR slot of COO entry k for every surviving entry (others keep their value)
__global__ void slot_of_entry_kernel(const uint32_t *__restrict__ segptr,
                                     const uint32_t *__restrict__ perm, int64_t nnz,
                                     int32_t *__restrict__ slot) {
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < nnz;
       s += (int64_t)gridDim.x * blockDim.x)
    for (uint32_t x = segptr[s]; x < segptr[s + 1]; ++x) slot[perm[x]] = (int32_t)s;
}

static inline int rows_blocks(int64_t n, int block) {
  int64_t g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > 148 * 32) g = 148 * 32;
  return (int)g;
}

}  // namespace skb

// Step 1: masks, row counters, scans.  Caller-provided scratch:
//   mask uint32[nbv*nel], rc uint64[nrows] (zeroed here), sums uint32[4096],
//   incstart / candstart uint32[nrows+1].
// counts_host[0] = nkeep (surviving COO entries), counts_host[1] = number of incidences with
// at least one surviving entry, after synchronising the stream.
extern "C" int skb_plan_rows_count(const int32_t *dofs_v, int32_t nbv, int32_t nbu, int64_t nel,
                                   int64_t nrows, const double *local_or_null, int drop_zeros,
                                   uint32_t *mask, unsigned long long *rc, uint32_t *sums,
                                   uint32_t *incstart, uint32_t *candstart, int64_t *counts_host,
                                   void *stream) {
  using namespace skb;
  cudaStream_t st = (cudaStream_t)stream;
  if (!dofs_v || nbv <= 0 || nbu <= 0 || nel < 0 || nrows <= 0 || !counts_host) return SKB_EINVAL;
  if (nbu > 32) return SKB_ETOOBIG;                      // one 32-bit column mask per incidence
  const int64_t ninc = (int64_t)nbv * nel;
  if (ninc >= (int64_t)0xffffffffLL || ninc * nbu >= (int64_t)0xffffffffLL) return SKB_ETOOBIG;
  counts_host[0] = counts_host[1] = 0;
  SKB_CUDA_TRY(cudaMemsetAsync(rc, 0, sizeof(unsigned long long) * (size_t)nrows, st));
  if (ninc) {
    rows_count_kernel<<<rows_blocks(ninc, 256), 256, 0, st>>>(dofs_v, nbv, nbu, nel,
                                                              local_or_null, drop_zeros, mask, rc);
    SKB_CUDA_TRY(cudaGetLastError());
    count_launch();
  }
  int rcode = exclusive_scan(rc, nrows, LoadHi(), sums, incstart, st);
  if (rcode != SKB_OK) return rcode;
  rcode = exclusive_scan(rc, nrows, LoadLo(), sums, candstart, st);
  if (rcode != SKB_OK) return rcode;
  uint32_t h[2] = {0, 0};
  SKB_CUDA_TRY(cudaMemcpyAsync(&h[0], candstart + nrows, 4, cudaMemcpyDeviceToHost, st));
  SKB_CUDA_TRY(cudaMemcpyAsync(&h[1], incstart + nrows, 4, cudaMemcpyDeviceToHost, st));
  SKB_CUDA_TRY(cudaStreamSynchronize(st));
  counts_host[0] = h[0];
  counts_host[1] = h[1];
  return SKB_OK;
}

// Step 2: incidence lists, per-row sorts, indptr.  Scratch: cursor uint32[2 * nrows],
// inc_words uint32[2 * ninc_kept], dofs_ut int32[nbu * nel] (unused if dofs_u is NULL), ucol /
// uoff uint32[nkeep], nuniq uint32[nrows], flag int32[3] (device).
// Writes perm uint32[nkeep], indptr int32[nrows+1]; *nnz_host after synchronising.
// Returns SKB_ETOOBIG if a row has more than 8192 surviving entries (the caller then uses the
// radix-sort path).
extern "C" int skb_plan_rows_sort(const int32_t *dofs_v, const int32_t *dofs_u, int32_t nbv,
                                  int32_t nbu, int64_t nel, int64_t nrows, const uint32_t *mask,
                                  const uint32_t *incstart, const uint32_t *candstart,
                                  uint32_t *cursor, uint32_t *inc_words, int32_t *dofs_ut,
                                  uint32_t *sums, uint32_t *perm,
                                  uint32_t *ucol, uint32_t *uoff, uint32_t *nuniq,
                                  int32_t *indptr, int32_t *flag, int64_t *nnz_host,
                                  void *stream) {
  using namespace skb;
  cudaStream_t st = (cudaStream_t)stream;
  if (!dofs_v || !nnz_host || nrows <= 0) return SKB_EINVAL;
  const int64_t ninc = (int64_t)nbv * nel;
  SKB_CUDA_TRY(cudaMemsetAsync(cursor, 0, 4 * (size_t)nrows, st));
  SKB_CUDA_TRY(cudaMemsetAsync(flag, 0, 12, st));
  uint2 *inc = reinterpret_cast<uint2 *>(inc_words);
  if (ninc) {
    rows_fill_kernel<<<rows_blocks(ninc, 256), 256, 0, st>>>(dofs_v, ninc, mask, incstart, cursor,
                                                             inc);
    SKB_CUDA_TRY(cudaGetLastError());
    if (dofs_u) {
      rows_transpose_kernel<<<rows_blocks((int64_t)nbu * nel, 256), 256, 0, st>>>(dofs_u, nbu, nel,
                                                                                   dofs_ut);
      SKB_CUDA_TRY(cudaGetLastError());
      count_launch();
    }
  }
  const int32_t *ut = dofs_u ? dofs_ut : nullptr;
  {
    // cursor[0, nrows) is free once the incidence lists are filled: it and cursor[nrows,
    // 2 nrows) hold the lists of the rows left to the shared-memory sorts (flag[1], flag[2]:
    // their lengths)
    uint32_t *mid = cursor, *lng = cursor + nrows;
    constexpr int W = 8;
    int64_t g = (nrows + W - 1) / W;
    if (g > 148 * 16) g = 148 * 16;
    rows_sort_warp_kernel<W><<<(unsigned)g, 32 * W, 0, st>>>(
        ut, nbv, nbu, nel, nrows, incstart, candstart, inc, perm, ucol, uoff, nuniq, mid, lng,
        flag + 1);
    SKB_CUDA_TRY(cudaGetLastError());
    auto ks = rows_sort_kernel<32, W, ROWS_SHORT_CAP>;
    const size_t sm = sizeof(unsigned long long) * W * ROWS_SHORT_CAP;
    SKB_CUDA_TRY(cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    ks<<<148 * 4, 32 * W, sm, st>>>(ut, nbv, nbu, nel, mid, flag + 1, incstart, candstart, inc,
                                    perm, ucol, uoff, nuniq, flag);
    SKB_CUDA_TRY(cudaGetLastError());
    auto kl = rows_sort_kernel<256, 1, ROWS_LONG_CAP>;
    const size_t sl = sizeof(unsigned long long) * ROWS_LONG_CAP;
    SKB_CUDA_TRY(cudaFuncSetAttribute(kl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sl));
    kl<<<148 * 3, 256, sl, st>>>(ut, nbv, nbu, nel, lng, flag + 2, incstart, candstart, inc, perm,
                                 ucol, uoff, nuniq, flag);
    SKB_CUDA_TRY(cudaGetLastError());
  }
  count_launch(4);
  int rcode = exclusive_scan(nuniq, nrows, LoadId(), sums, indptr, st);
  if (rcode != SKB_OK) return rcode;
  int32_t h[2] = {0, 0};
  SKB_CUDA_TRY(cudaMemcpyAsync(&h[0], indptr + nrows, 4, cudaMemcpyDeviceToHost, st));
  SKB_CUDA_TRY(cudaMemcpyAsync(&h[1], flag, 4, cudaMemcpyDeviceToHost, st));
  SKB_CUDA_TRY(cudaStreamSynchronize(st));
  if (h[1] || h[0] < 0) return SKB_ETOOBIG;              // int32 indptr (scipy would use int64)
  *nnz_host = h[0];
  return SKB_OK;
}

// Step 3: indices int32[nnz], segptr uint32[nnz+1].
extern "C" int skb_plan_rows_emit(int64_t nrows, int64_t nnz, int64_t nkeep,
                                  const uint32_t *candstart, const int32_t *indptr,
                                  const uint32_t *ucol, const uint32_t *uoff, int32_t *indices,
                                  uint32_t *segptr, void *stream) {
  using namespace skb;
  if (nrows <= 0 || nnz < 0 || !segptr) return SKB_EINVAL;
  int64_t g = (nrows * 32 + 255) / 256;
  if (g > 148 * 32) g = 148 * 32;
  if (g < 1) g = 1;
  rows_emit_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(nrows, candstart, indptr, ucol,
                                                                  uoff, indices, segptr, nnz, nkeep);
  count_launch();
  return (int)cudaGetLastError();
}

// Masks for mesh entities (see entity_mask_kernel): tu int32[nbu][nel] vertex ids, vmax
// int32[nbv][nel], adj_host uint32[nbv] (nbv, nbu <= 32); mask uint32[nbv * nel] out.  Feed it
// to skb_plan_rows_count with drop_zeros == 2 and local == NULL.
extern "C" int skb_entity_masks(const int32_t *tu, int32_t nbu, int32_t nbv, int64_t nel,
                                const int32_t *vmax, const uint32_t *adj_host, uint32_t *mask,
                                void *stream) {
  using namespace skb;
  if (!tu || !vmax || !adj_host || !mask || nbu <= 0 || nbu > 32 || nbv <= 0 || nbv > 32 ||
      nel < 0)
    return SKB_EINVAL;
  if (nel == 0) return SKB_OK;
  EntityAdj a;
  for (int i = 0; i < 32; ++i) a.adj[i] = i < nbv ? adj_host[i] : 0u;
  entity_mask_kernel<<<rows_blocks((int64_t)nbv * nel, 256), 256, 0, (cudaStream_t)stream>>>(
      tu, nbu, nbv, nel, vmax, a, mask);
  count_launch();
  return (int)cudaGetLastError();
}

// slot[perm[x]] = s for x in [segptr[s], segptr[s+1]): the CSR slot of every surviving COO entry
// (slot int32[ncoo], entries that did not survive are left untouched).
extern "C" int skb_plan_slot_of_entry(const uint32_t *segptr, const uint32_t *perm, int64_t nnz,
                                      int32_t *slot, void *stream) {
  using namespace skb;
  if (nnz < 0 || !segptr || !slot) return SKB_EINVAL;
  if (nnz == 0) return SKB_OK;
  slot_of_entry_kernel<<<rows_blocks(nnz, 256), 256, 0, (cudaStream_t)stream>>>(segptr, perm, nnz,
                                                                                slot);
  count_launch();
  return (int)cudaGetLastError();
}

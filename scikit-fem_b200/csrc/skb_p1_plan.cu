// Plan-time helper of the fused P1 path: bank-conflict-free ordering of the
// sliced-ELL staging indices.
//
// In P2 of p1tet_laplace_fused_kernel every lane of a 32-lane group adds the
// staged values in[ids[c*32 + lane]], c = 0..len-1.  The order in which a lane
// adds its terms is free (the reference's CSR sum order is unspecified, SURVEY
// A.9; it only has to be fixed at plan time), so the columns can be chosen to
// spread each LDS.64 over the shared-memory banks.  The hardware serves an
// LDS.64 in two half-warp passes; a pass needs as many wavefronts as the most
// loaded bank pair has distinct 8-byte words (measured: this model reproduces
// ncu's "L1 Wavefronts Shared" of the gather instructions, tools/
// sim_smem_conflicts.py).  For one half-group (16 lanes x len columns) this is
// an edge colouring of the bipartite multigraph lanes x banks: split every bank
// into ceil(deg/len) sub-banks of degree <= len, then colour properly with len
// colours (Koenig) using alternating-path recolouring; the column of a term is
// its colour, so a bank pair receives at most ceil(deg/len) words per column.
// Unused (lane, column) cells read one of 16 staged zeros (one per bank pair),
// the one in the least loaded bank pair of that column.
//
// One thread per half-group, everything in local memory; runs once per plan.
#include "skb_common.cuh"

namespace skb {

constexpr int SP_MAXLEN = 16;    // columns handled (longer groups keep their order)
constexpr int SP_MAXSUB = 48;    // sub-banks: 16 + 16*len/len
constexpr int SP_MAXE = 16 * SP_MAXLEN;
constexpr unsigned char SP_NONE = 0xFF;

__global__ void __launch_bounds__(64)
p1_plan_spread_kernel(uint16_t *__restrict__ rec16, const int64_t *__restrict__ grp_pos,
                      const int32_t *__restrict__ grp_len, int64_t ngroups, int zero_base) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= 2 * ngroups) return;
  const int64_t g = tid >> 1;
  const int half = (int)(tid & 1);
  const int len = grp_len[g];
  if (len <= 0 || len > SP_MAXLEN) return;
  uint16_t *ids = rec16 + grp_pos[g] + 16 * half;       // cell (c, l) at ids[c*32 + l]

  unsigned char eu[SP_MAXE], ev[SP_MAXE], ec[SP_MAXE];  // edge: lane, sub-bank, colour
  uint16_t ew[SP_MAXE];                                 // staged word index
  unsigned char lane_col[16][SP_MAXLEN], sub_col[SP_MAXSUB][SP_MAXLEN];
  unsigned char sub_bank[SP_MAXSUB];
  int deg[16], sub0[16], seen[16];
  // edges = non-padding cells; the bank pair of word w is w & 15
  int ne = 0;
#pragma unroll 1
  for (int b = 0; b < 16; ++b) { deg[b] = 0; seen[b] = 0; }
#pragma unroll 1
  for (int l = 0; l < 16; ++l)
    for (int c = 0; c < len; ++c) {
      const uint16_t w = ids[c * 32 + l];
      lane_col[l][c] = SP_NONE;
      if ((int)w < zero_base) {
        eu[ne] = (unsigned char)l;
        ew[ne] = w;
        ++deg[w & 15];
        ++ne;
      }
    }
  if (ne == 0 || ne >= 255) return;
  int nsub = 0;
#pragma unroll 1
  for (int b = 0; b < 16; ++b) {
    sub0[b] = nsub;
    const int k = (deg[b] + len - 1) / len;
    for (int i = 0; i < k; ++i) sub_bank[nsub + i] = (unsigned char)b;
    nsub += k;
  }
  if (nsub > SP_MAXSUB) return;
#pragma unroll 1
  for (int s = 0; s < nsub; ++s)
    for (int c = 0; c < len; ++c) sub_col[s][c] = SP_NONE;
  // edges of a bank go round-robin to its sub-banks (degree <= len each)
#pragma unroll 1
  for (int e = 0; e < ne; ++e) {
    const int b = ew[e] & 15;
    const int k = (deg[b] + len - 1) / len;
    ev[e] = (unsigned char)(sub0[b] + (seen[b]++ % k));
  }
  // proper edge colouring with `len` colours
#pragma unroll 1
  for (int e = 0; e < ne; ++e) {
    const int u = eu[e], v = ev[e];
    int a = -1, b = -1, both = -1;
    for (int c = 0; c < len; ++c) {
      const bool fu = lane_col[u][c] == SP_NONE, fv = sub_col[v][c] == SP_NONE;
      if (fu && fv && both < 0) both = c;
      if (fu && a < 0) a = c;
      if (fv && b < 0) b = c;
    }
    if (both < 0) {
      // a is free at the lane, b at the sub-bank: flip the a/b alternating path
      // that starts at the sub-bank; it cannot end at this lane (bipartite)
      int x = v, col = a;
      bool at_sub = true;
      unsigned char path[2 * SP_MAXSUB + 34];
      int np = 0;
      while (true) {
        const unsigned char e2 = at_sub ? sub_col[x][col] : lane_col[x][col];
        if (e2 == SP_NONE || np >= (int)sizeof(path)) break;
        path[np++] = e2;
        x = at_sub ? eu[e2] : ev[e2];
        at_sub = !at_sub;
        col = col == a ? b : a;
      }
      for (int i = 0; i < np; ++i) {
        const unsigned char e2 = path[i];
        lane_col[eu[e2]][ec[e2]] = SP_NONE;
        sub_col[ev[e2]][ec[e2]] = SP_NONE;
      }
      for (int i = 0; i < np; ++i) {
        const unsigned char e2 = path[i];
        ec[e2] = (unsigned char)(ec[e2] == a ? b : a);
        lane_col[eu[e2]][ec[e2]] = e2;
        sub_col[ev[e2]][ec[e2]] = e2;
      }
      both = a;
    }
    ec[e] = (unsigned char)both;
    lane_col[u][both] = (unsigned char)e;
    sub_col[v][both] = (unsigned char)e;
  }
  // write back: column c of lane l = its edge of colour c, else the staged zero
  // in the least loaded bank pair of that column
#pragma unroll 1
  for (int c = 0; c < len; ++c) {
    int load[16];
    for (int b = 0; b < 16; ++b) load[b] = 0;
    for (int s = 0; s < nsub; ++s)
      if (sub_col[s][c] != SP_NONE) ++load[sub_bank[s]];
    int best = 0;
    for (int b = 1; b < 16; ++b)
      if (load[b] < load[best]) best = b;
    for (int l = 0; l < 16; ++l) {
      const unsigned char e = lane_col[l][c];
      ids[c * 32 + l] = e == SP_NONE ? (uint16_t)(zero_base + best) : ew[e];
    }
  }
}

}  // namespace skb

extern "C" int skb_p1_plan_spread(uint16_t *rec16, const int64_t *grp_pos,
                                  const int32_t *grp_len, int64_t ngroups, int32_t zero_base,
                                  void *stream) {
  using namespace skb;
  if (ngroups < 0 || zero_base <= 0 || (zero_base & 15)) return SKB_EINVAL;
  if (ngroups == 0) return SKB_OK;
  if (!rec16 || !grp_pos || !grp_len) return SKB_EINVAL;
  const int64_t nthreads = 2 * ngroups;
  p1_plan_spread_kernel<<<(unsigned)((nthreads + 63) / 64), 64, 0, (cudaStream_t)stream>>>(
      rec16, grp_pos, grp_len, ngroups, zero_base);
  count_launch();
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Tile-local vertex renumbering against the bank conflicts of P1's coordinate
// gathers.  Thread e of a warp reads sx[tl[e].a] (a = 0..3): an LDS.64 served in
// two half-warp passes, each as long as its most loaded bank pair (id mod 16).
// Numbering the tile's vertices by ascending global id leaves 3.2 wavefronts per
// gather on the 100^3 mesh (ncu 3.3); a greedy colouring - vertices taken in order
// of first appearance, each getting the bank pair least used so far by the (half-
// warp, slot) sets it belongs to, ids = colour + 16 * rank within the colour -
// brings the model to 2.0, the minimum (tools/sim_smem_conflicts.py).
// One thread per tile, in place on the record: rewrites tl and permutes verts;
// unused ids (the vertex section is sized 16 * (ceil(nv/16) + 1) by the plan
// builder) keep pointing at a valid vertex.
// ---------------------------------------------------------------------------
namespace skb {

constexpr int RN_MAXV = 1024;   // vertices per tile handled (else the numbering is kept)
constexpr int RN_MAXS = 16;     // (half-warp, slot) sets per vertex tracked

__global__ void __launch_bounds__(32)
p1_plan_renumber_kernel(unsigned char *__restrict__ rec, const uint64_t *__restrict__ rec_start,
                        int ntiles, int T) {
  const int tile = blockIdx.x * blockDim.x + threadIdx.x;
  if (tile >= ntiles) return;
  unsigned char *r = rec + rec_start[tile];
  uint32_t *hdr = reinterpret_cast<uint32_t *>(r);
  const int cap_ids = (int)hdr[0];                       // size of the vertex section
  uint16_t *tl = reinterpret_cast<uint16_t *>(r + 32);   // [T][4]
  int32_t *verts = reinterpret_cast<int32_t *>(r + hdr[2]);
  const int nsets = (T / 16) * 4;
  if (cap_ids > RN_MAXV || nsets > 255) return;
  uint16_t newid[RN_MAXV];          // old id -> new id (0xFFFF = not seen yet)
  uint16_t order[RN_MAXV];          // old ids in order of first appearance
  unsigned char vset[RN_MAXV][RN_MAXS], vcnt[RN_MAXV];
  unsigned char mask[256][16];      // per set: vertices per bank pair so far
  int nv = 0, nseen = 0;
  for (int i = 0; i < cap_ids; ++i) { newid[i] = 0xFFFF; vcnt[i] = 0; }
  for (int s = 0; s < nsets; ++s)
    for (int c = 0; c < 16; ++c) mask[s][c] = 0;
  for (int e = 0; e < T; ++e) {
    if (tl[e * 4] == 0xFFFF) continue;                   // padding element of the last tile
    for (int a = 0; a < 4; ++a) {
      const int v = tl[e * 4 + a];
      if (v >= cap_ids) return;                          // malformed: keep everything
      if (newid[v] == 0xFFFF) { newid[v] = 0xFFFE; order[nseen++] = (uint16_t)v; }
      if (v + 1 > nv) nv = v + 1;
      const unsigned char sid = (unsigned char)((e >> 4) * 4 + a);
      bool have = false;
      for (int k = 0; k < vcnt[v]; ++k) have |= vset[v][k] == sid;
      if (!have && vcnt[v] < RN_MAXS) vset[v][vcnt[v]++] = sid;
    }
  }
  if (nseen == 0) return;
  const int per_colour = (nv + 15) / 16 + 1;             // ids < 16 * per_colour <= cap_ids
  if (16 * per_colour > cap_ids) return;
  int used[16];
  for (int c = 0; c < 16; ++c) used[c] = 0;
  for (int i = 0; i < nseen; ++i) {
    const int v = order[i];
    int best = -1, best_cost = 0x7fffffff;
    for (int c = 0; c < 16; ++c) {
      if (used[c] >= per_colour) continue;
      int cost = 0;
      for (int k = 0; k < vcnt[v]; ++k) cost += mask[vset[v][k]][c];
      cost = cost * 1024 + used[c];
      if (cost < best_cost) { best_cost = cost; best = c; }
    }
    newid[v] = (uint16_t)(best + 16 * used[best]);
    ++used[best];
    for (int k = 0; k < vcnt[v]; ++k) ++mask[vset[v][k]][best];
  }
  // permute the vertex list (through `order`/registers: no second buffer in the record)
  int32_t keep = verts[order[0]];
  // old global ids are needed after they are overwritten: stash them in local memory
  int32_t oldv[RN_MAXV];
  for (int i = 0; i < nv; ++i) oldv[i] = verts[i];
  for (int i = 0; i < cap_ids; ++i) verts[i] = keep;     // holes point at a valid vertex
  for (int i = 0; i < nseen; ++i) verts[newid[order[i]]] = oldv[order[i]];
  for (int e = 0; e < T; ++e) {
    if (tl[e * 4] == 0xFFFF) continue;
    for (int a = 0; a < 4; ++a) tl[e * 4 + a] = newid[tl[e * 4 + a]];
  }
}

}  // namespace skb

extern "C" int skb_p1_plan_renumber(void *rec, const uint64_t *rec_start, int32_t ntiles,
                                    int32_t tile_elems, void *stream) {
  using namespace skb;
  if (ntiles < 0 || tile_elems <= 0 || (tile_elems & 15)) return SKB_EINVAL;
  if (ntiles == 0) return SKB_OK;
  if (!rec || !rec_start) return SKB_EINVAL;
  p1_plan_renumber_kernel<<<(ntiles + 31) / 32, 32, 0, (cudaStream_t)stream>>>(
      (unsigned char *)rec, rec_start, ntiles, tile_elems);
  count_launch();
  return (int)cudaGetLastError();
}

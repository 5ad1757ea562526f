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

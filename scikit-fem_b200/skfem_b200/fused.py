"""Plan of the fused P1 Laplace path (csrc/skb_p1_fused.cu).

Built once per (basis, CSR pattern) on the device, reused by every warm
re-assembly.  Input: the connectivity ``t``, the vertex coordinates (only to
order elements along a Morton curve - any order is valid, a spatially compact
one keeps tiles' slot sets small) and the CSR pattern ``indptr/indices``
produced by the generic plan (skb_plan_*), which already encodes the
value-dependent zero elimination of the reference
(skfem/assembly/form/coo_data.py:35).

Output (all device arrays):
  tt                 (ntiles*T, 4) int32   tile-ordered connectivity, -1 padded
  tile_slot_start    (ntiles+1,)  first tile slot of each tile
  tile_contrib_start (ntiles+1,)  first contributor of each tile
  slot_ptr           per tile (nslots+1) uint16 offsets into its contributors
  contrib            uint16 staging indices  k(a,b)*T + e_local
  meta               per tile slot: CSR slot, or 0x80000000|scratch position
  sptr, gslot        per shared CSR slot: its scratch range and CSR slot

The preprocessing itself uses torch sort / unique / searchsorted (cold path,
plumbing); the warm path runs only this package's kernels.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _torch():
    import torch
    return torch


def _spread10(v):
    """Insert two zero bits between the 10 low bits of v (Morton interleave)."""
    v = v & 0x3FF
    v = (v | (v << 16)) & 0x30000FF
    v = (v | (v << 8)) & 0x300F00F
    v = (v | (v << 4)) & 0x30C30C3
    v = (v | (v << 2)) & 0x9249249
    return v


class P1FusedPlan:
    pass


def applicable(basis, form):
    from .element import ElementTetP1
    if form.native is None or form.native[1] != _lib.FORM_LAPLACE:
        return False
    if not isinstance(basis.elem, ElementTetP1) or not basis._affine:
        return False
    W = basis.W
    return bool(np.all(W == W[0]))


def build(basis, plan, T=1024, threads=256):
    torch = _torch()
    d = basis._dev()
    dev = d["device"]
    t = d["t"] if d["tind"] is None else d["t"][:, d["tind"].long()]
    p = d["p"]
    nel = int(t.shape[1])
    nnz = plan.nnz
    N = int(plan.shape[1])
    fp = P1FusedPlan()
    fp.T, fp.threads, fp.nel, fp.nnz = T, threads, nel, nnz
    ntiles = (nel + T - 1) // T
    fp.ntiles = ntiles
    i64 = torch.int64
    tl = t.long()
    # 1. Morton order of element centroids
    cent = p[:, tl].sum(dim=1)                      # (3, nel), 4x centroid
    lo = cent.min(dim=1, keepdim=True).values
    hi = cent.max(dim=1, keepdim=True).values
    q = ((cent - lo) / torch.clamp(hi - lo, min=1e-300) * 1023.0).clamp(0, 1023).to(i64)
    code = _spread10(q[0]) | (_spread10(q[1]) << 1) | (_spread10(q[2]) << 2)
    order = torch.argsort(code, stable=True)
    del cent, q, code
    tt = tl[:, order].t().contiguous()              # (nel, 4) int64, tile order
    e_idx = torch.arange(nel, device=dev, dtype=i64)
    tile_of = e_idx // T
    e_loc = e_idx - tile_of * T
    tile_ids = torch.arange(ntiles + 1, device=dev, dtype=i64)
    # 2. tile-local vertex numbering
    nv = int(p.shape[1])
    vkey = (tile_of[:, None] * nv + tt).reshape(-1)
    uv, vinv = torch.unique(vkey, sorted=True, return_inverse=True)
    uv_tile = uv // nv
    tile_vert_start = torch.searchsorted(uv_tile, tile_ids)
    fp.vcap = int((tile_vert_start[1:] - tile_vert_start[:-1]).max())
    if fp.vcap > 0xFFFF - 1:
        raise RuntimeError("fused plan: tile touches too many vertices")
    loc = (vinv - tile_vert_start[tile_of].repeat_interleave(4)).reshape(nel, 4)
    pad = ntiles * T - nel
    if pad:
        loc = torch.cat([loc, torch.full((pad, 4), 0xFFFF, dtype=i64, device=dev)])
    fp.tl = loc.to(torch.int16).contiguous()
    fp.tile_verts = (uv - uv_tile * nv).to(torch.int32).contiguous()
    fp.tile_vert_start = tile_vert_start.to(torch.int32).contiguous()
    del vkey, uv, vinv, loc, uv_tile
    # 3. CSR slot of every local entry (a, b)
    counts = (plan.indptr[1:] - plan.indptr[:-1]).long()
    row_of_slot = torch.repeat_interleave(torch.arange(N, device=dev, dtype=i64), counts)
    csr_key = row_of_slot * N + plan.indices.long()  # ascending (canonical CSR)
    del row_of_slot
    keys2, sids = [], []
    for a in range(4):
        for b in range(4):
            key = tt[:, a] * N + tt[:, b]            # row = test dof (v), col = trial dof (u)
            pos = torch.searchsorted(csr_key, key).clamp(max=max(nnz - 1, 0))
            ok = csr_key[pos] == key
            lo_, hi_ = (a, b) if a <= b else (b, a)
            k = lo_ * 4 - (lo_ * (lo_ - 1)) // 2 + (hi_ - lo_)
            keys2.append((tile_of * nnz + pos)[ok])
            sids.append((k * T + e_loc)[ok])
    key2 = torch.cat(keys2)
    sid = torch.cat(sids)
    del keys2, sids, csr_key
    key2, perm = torch.sort(key2, stable=True)
    sid = sid[perm]
    del perm
    # 4. tile slots = unique (tile, csr slot) pairs, contributions grouped per slot
    uniq, sinv, cnt = torch.unique_consecutive(key2, return_inverse=True, return_counts=True)
    del key2
    nts = int(uniq.shape[0])
    ncontrib = int(sid.shape[0])
    if int(cnt.max()) > 255:
        raise RuntimeError("fused plan: more than 255 contributions to one slot in a tile")
    ts_tile = uniq // nnz
    ts_gslot = uniq - ts_tile * nnz
    cs = torch.cumsum(cnt, 0) - cnt                  # first contribution of each slot
    kth = torch.arange(ncontrib, device=dev, dtype=i64) - cs[sinv]
    # within a tile, order slots by decreasing contribution count (sliced ELL)
    order3 = torch.argsort(ts_tile * 256 + (255 - cnt), stable=True)
    newpos = torch.empty(nts, dtype=i64, device=dev)
    newpos[order3] = torch.arange(nts, device=dev, dtype=i64)
    ts_tile, ts_gslot, cnt = ts_tile[order3], ts_gslot[order3], cnt[order3]
    tile_slot_start = torch.searchsorted(ts_tile, tile_ids)
    nslots_tile = tile_slot_start[1:] - tile_slot_start[:-1]
    ngroups_tile = (nslots_tile + 31) // 32
    tile_group_start = torch.cat([torch.zeros(1, dtype=i64, device=dev),
                                  torch.cumsum(ngroups_tile, 0)])
    ngroups = int(tile_group_start[-1])
    ts_idx = torch.arange(nts, device=dev, dtype=i64)
    j_in_tile = ts_idx - tile_slot_start[ts_tile]
    grp_of_slot = tile_group_start[ts_tile] + j_in_tile // 32
    lane_of_slot = j_in_tile % 32
    grp_len = torch.zeros(ngroups, dtype=i64, device=dev)
    grp_len.scatter_reduce_(0, grp_of_slot, cnt, reduce="amax", include_self=True)
    # per-tile index regions, each starting at a multiple of 8 indices (16 B)
    gsz_ids = grp_len * 32
    gcum = torch.cat([torch.zeros(1, dtype=i64, device=dev), torch.cumsum(gsz_ids, 0)])
    tile_ids_n = gcum[tile_group_start[1:]] - gcum[tile_group_start[:-1]]   # multiple of 32
    tile_contrib_start = torch.cat([torch.zeros(1, dtype=i64, device=dev),
                                    torch.cumsum(tile_ids_n, 0)])
    grp_tile = torch.repeat_interleave(torch.arange(ntiles, device=dev, dtype=i64),
                                       ngroups_tile)
    grp_base = gcum[:-1] - gcum[tile_group_start[:-1]][grp_tile]            # tile-relative
    ncontrib_sell = int(gcum[-1])
    s_new = newpos[sinv]
    g_of = grp_of_slot[s_new]
    cpos = tile_contrib_start[grp_tile[g_of]] + grp_base[g_of] + kth * 32 + lane_of_slot[s_new]
    zero_idx = 10 * T                                 # staged 0.0: padding adds nothing
    contrib = torch.full((max(ncontrib_sell, 8),), zero_idx, dtype=i64, device=dev)
    contrib[cpos] = sid
    fp.tile_slot_start = tile_slot_start.to(torch.int32).contiguous()
    fp.tile_group_start = tile_group_start.to(torch.int32).contiguous()
    fp.tile_contrib_start = tile_contrib_start.to(torch.int32).contiguous()
    fp.grp_base = grp_base.to(torch.int32).contiguous()
    fp.grp_len = grp_len.to(torch.int16).contiguous()
    fp.contrib = contrib.to(torch.int16).contiguous()
    max_ids = int(tile_ids_n.max()) if ntiles else 0
    fp.aux_bytes = ((max(fp.vcap * 32, max_ids * 2, 16) + 15) // 16) * 16
    del sid, contrib, cpos, s_new, kth, sinv
    # 5. slots touched by one tile go straight to csr_data, the others through scratch
    order2 = torch.argsort(ts_gslot * ntiles + ts_tile)  # by csr slot, tiles ascending
    g_sorted = ts_gslot[order2]
    ug, gcnt = torch.unique_consecutive(g_sorted, return_counts=True)
    if int(ug.shape[0]) != nnz:
        raise RuntimeError("fused plan: CSR pattern has slots no element contributes to")
    shared = gcnt > 1
    gsz = gcnt * shared
    gstart = torch.cumsum(gsz, 0) - gsz
    gfirst = torch.cumsum(gcnt, 0) - gcnt
    grp = torch.repeat_interleave(torch.arange(nnz, device=dev, dtype=i64), gcnt)
    rank = ts_idx - gfirst[grp]
    spos = gstart[grp] + rank
    meta_sorted = torch.where(shared[grp], spos | 0x80000000, g_sorted)
    meta = torch.empty(nts, dtype=i64, device=dev)
    meta[order2] = meta_sorted
    fp.meta = meta.to(torch.int32).contiguous()
    sh = torch.nonzero(shared).flatten()
    fp.nshared = int(sh.shape[0])
    fp.nscratch = int(gsz.sum())
    fp.gslot = sh.to(torch.int32).contiguous()
    fp.sptr = torch.cat([gstart[sh], torch.tensor([fp.nscratch], device=dev, dtype=i64)]
                        ).to(torch.int32).contiguous()
    fp.scratch = torch.empty(max(fp.nscratch, 1), dtype=torch.float64, device=dev)
    fp.nts, fp.ncontrib, fp.ncontrib_sell, fp.ngroups = nts, ncontrib, ncontrib_sell, ngroups
    fp.nverts_tiles = int(fp.tile_verts.shape[0])
    fp.w = float(basis.W[0])
    fp.nqp = int(basis.nqp)
    fp.p = p
    smem = 8 * (10 * T + 2) + fp.aux_bytes
    if smem > 226 * 1024:
        raise RuntimeError("fused plan: tile does not fit in shared memory")
    return fp


def run(fp, data, stream):
    """Warm numeric phase: two kernel launches, nothing else."""
    lib = _lib.lib()
    code = lib.skb_p1tet_laplace_fused(
        fp.p.data_ptr(), fp.p.shape[1], fp.tl.data_ptr(), fp.ntiles, fp.T, fp.threads,
        fp.tile_vert_start.data_ptr(), fp.tile_verts.data_ptr(), fp.aux_bytes,
        fp.tile_slot_start.data_ptr(), fp.tile_group_start.data_ptr(),
        fp.tile_contrib_start.data_ptr(), fp.grp_base.data_ptr(), fp.grp_len.data_ptr(),
        fp.contrib.data_ptr(), fp.meta.data_ptr(), C.c_double(fp.w), fp.nqp, data.data_ptr(),
        fp.scratch.data_ptr(), stream)
    _lib.check(code, "skb_p1tet_laplace_fused")
    code = lib.skb_p1_combine(fp.scratch.data_ptr(), fp.sptr.data_ptr(), fp.gslot.data_ptr(),
                              fp.nshared, data.data_ptr(), stream)
    _lib.check(code, "skb_p1_combine")


def stats(fp):
    """Bytes the fused step moves (for DESIGN.md / the roofline discussion)."""
    b = {
        "tl": fp.ntiles * fp.T * 8, "tile_verts": fp.nverts_tiles * 4,
        "p_gather_min": fp.nverts_tiles * 24, "contrib": fp.ncontrib_sell * 2,
        "meta": fp.nts * 4, "groups": fp.ngroups * 6,
        "direct_out": (fp.nnz - fp.nshared) * 8, "scratch_w": fp.nscratch * 8,
        "scratch_r": fp.nscratch * 8, "sptr_gslot": fp.nshared * 8, "combine_out": fp.nshared * 8,
    }
    b["total"] = sum(b.values())
    b["per_element"] = b["total"] / max(fp.nel, 1)
    b["tile_slots_per_csr_slot"] = fp.nts / max(fp.nnz, 1)
    b["vcap"] = fp.vcap
    b["sell_padding"] = fp.ncontrib_sell / max(fp.ncontrib, 1)
    return b

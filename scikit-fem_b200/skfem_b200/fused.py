"""Plan of the fused P1 Laplace path (csrc/skb_p1_fused.cu).

Built once per (basis, CSR pattern) on the device, reused by every warm
re-assembly.  Input: the connectivity ``t``, the vertex coordinates (only to
order elements along a Morton curve - any order is valid, a spatially compact
one keeps the tiles' slot sets small) and the CSR pattern ``indptr/indices``
produced by the generic plan (skb_plan_*), which already encodes the
value-dependent zero elimination of the reference
(skfem/assembly/form/coo_data.py:35).

Elements are cut into tiles of ``T``.  Every tile gets one contiguous,
16-byte aligned *record* (fetched by a single TMA bulk copy in the kernel):

    header  8 x uint32: nverts, ngroups, off_verts, off_grp, off_meta, off_ids,
            off_fsel, off_meta2
    tl      T x 4 uint16  tile-local vertex ids (0xFFFF = padding element)
    verts   nverts int32  global vertex ids of the tile
    grp     per group of 32 lanes: uint32 (offset/32 into ids) | len << 16
    meta    per lane: CSR slot, or 0x80000000 | scratch position
            (0xFFFFFFFF for lanes that write nothing)
    meta2   per lane: the mirror CSR slot (col,row) that receives the same sum
            (0xFFFFFFFF on the diagonal / for shared slots)
    fsel    per lane, one byte: log2 of the number of adjacent lanes whose
            partial sums the leader lane combines (long lists are split)
    ids     sliced-ELL uint16 staging indices k(a,b)*T + e_local; entry c of
            lane l of a group sits at base + 32 c + l; short lists are padded
            with 10*T, the index of a staged 0.0

A *tile slot* is a canonical (row <= col) CSR slot touched by the tile; the
Laplace local matrix is bitwise symmetric, so the mirror slot gets the same
sum.  Slots touched by one tile only are written straight to ``csr_data``; the
others go to ``scratch`` (grouped by CSR slot, tiles ascending) and are added
by ``skb_p1_combine``.

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


def _kd_order(corner, T, order=None, seg_cnt=None):
    """Element order whose consecutive chunks of T are compact boxes: a balanced k-d tree
    over the elements' bounding-box corners, every split placed after a whole number of tiles
    (left child = floor(k/2) full tiles), sorted along the segment's longest axis with the
    other two axes as tie-breakers.  All levels are processed at once with segmented sorts
    (two stable argsorts per level, ~log2(ntiles) levels).  Compared with cutting a Morton
    curve every T elements this leaves fewer CSR slots shared between tiles (C2: 49 % instead
    of 65 % of the canonical slots, 29 % fewer per-tile partials; tools/store_sectors.py).
    ``order`` / ``seg_cnt`` (optional): start from this element order already cut into
    segments of these sizes and refine every segment separately (nested tilings)."""
    torch = _torch()
    dev = corner.device
    i64 = torch.int64
    nel = int(corner.shape[1])
    lo = corner.min(dim=1, keepdim=True).values
    hi = corner.max(dim=1, keepdim=True).values
    scale = float(2 ** 20 - 1)
    # one scale for all axes: a split along the (physically) longest side of a box keeps the
    # tiles compact also when the domain itself is a thin slab (multi-GPU parts)
    span = torch.clamp((hi - lo).max(), min=1e-300)
    q = ((corner - lo) / span * scale).clamp(0, scale).to(i64)
    if order is None:
        order = torch.arange(nel, device=dev, dtype=i64)
        seg_cnt = torch.tensor([nel], device=dev, dtype=i64)
    pos = torch.arange(nel, device=dev, dtype=i64)
    seg_k = (seg_cnt + T - 1) // T
    while int(seg_k.max()) > 1:
        nseg = int(seg_cnt.shape[0])
        seg_of = torch.repeat_interleave(torch.arange(nseg, device=dev, dtype=i64), seg_cnt)
        qq = q[:, order]
        idx3 = seg_of.expand(3, -1)
        mn = torch.full((3, nseg), 2 ** 20, device=dev, dtype=i64).scatter_reduce_(
            1, idx3, qq, reduce="amin", include_self=True)
        mx = torch.full((3, nseg), -1, device=dev, dtype=i64).scatter_reduce_(
            1, idx3, qq, reduce="amax", include_self=True)
        axs = torch.argsort(mn - mx, dim=0, stable=True)        # longest extent first
        k0 = torch.gather(qq, 0, axs[0][seg_of][None])[0]
        k1 = torch.gather(qq, 0, axs[1][seg_of][None])[0]
        k2 = torch.gather(qq, 0, axs[2][seg_of][None])[0]
        key = (k0 << 42) | (k1 << 21) | k2
        key = torch.where((seg_k == 1)[seg_of], pos, key)        # finished tiles keep their order
        p1 = torch.argsort(key, stable=True)
        perm = p1[torch.argsort(seg_of[p1], stable=True)]
        order = order[perm]
        split = seg_k > 1
        kl = torch.where(split, seg_k // 2, seg_k)
        kr = seg_k - kl
        cl = torch.where(split, kl * T, seg_cnt)
        cnt2 = torch.stack([cl, seg_cnt - cl], dim=1).reshape(-1)
        k2n = torch.stack([kl, kr], dim=1).reshape(-1)
        keep = k2n > 0
        seg_cnt, seg_k = cnt2[keep], k2n[keep]
    return order


class P1FusedPlan:
    pass


def applicable(basis, form, version=1):
    """laplace on ElementTetP1 with an equal-weight rule; the second-generation kernel also
    takes mass (u * v) with the 4-point rule."""
    from .element import ElementTetP1
    if form.native is None or form.native[0] != "bilinear" or form.native[3] != "scalar":
        return False
    if not getattr(basis, "_native_ok", True):
        return False
    if not isinstance(basis.elem, ElementTetP1) or not basis._affine:
        return False
    W = basis.W
    if form.native[1] == _lib.FORM_LAPLACE:
        return bool(np.all(W == W[0]))
    if form.native[1] == _lib.FORM_MASS:
        return version == 2 and int(basis.nqp) == 4
    return False


def build(basis, plan, T=512, threads=480, ring=4, slot_map=None, spread=True, renumber=True,
          tiling="morton"):
    """``slot_map`` (optional int64 tensor, CSR slot -> output index): targets
    written by the kernels are remapped through it (multi-GPU direct write).
    ``tiling``: "morton" cuts a Morton curve over the element centroids every T elements;
    "kd" builds compact boxes with a balanced k-d tree (``_kd_order``)."""
    if tiling not in ("morton", "kd"):
        raise ValueError("fused plan: unknown tiling '{}'".format(tiling))
    torch = _torch()
    d = basis._dev()
    dev = d["device"]
    t = d["t"] if d["tind"] is None else d["t"][:, d["tind"].long()]
    p = d["p"]
    nel = int(t.shape[1])
    nnz = plan.nnz
    N = int(plan.shape[1])
    fp = P1FusedPlan()
    fp.T, fp.threads, fp.ring, fp.nel, fp.nnz = T, threads, ring, nel, nnz
    fp.tiling = tiling
    ntiles = (nel + T - 1) // T
    fp.ntiles = ntiles
    i64 = torch.int64

    def arange(n):
        return torch.arange(n, device=dev, dtype=i64)

    def excl(x):
        return torch.cumsum(x, 0) - x

    tl = t.long()
    # 1. element order: consecutive chunks of T elements are the tiles
    if tiling == "kd":
        order = _kd_order(p[:, tl].min(dim=1).values, T)
    else:                                           # Morton order of element centroids
        cent = p[:, tl].sum(dim=1)                  # (3, nel), 4x centroid
        lo = cent.min(dim=1, keepdim=True).values
        hi = cent.max(dim=1, keepdim=True).values
        q = ((cent - lo) / torch.clamp(hi - lo, min=1e-300) * 1023.0).clamp(0, 1023).to(i64)
        code = _spread10(q[0]) | (_spread10(q[1]) << 1) | (_spread10(q[2]) << 2)
        order = torch.argsort(code, stable=True)
        del cent, q, code
    tt = tl[:, order].t().contiguous()              # (nel, 4) int64, tile order
    e_idx = arange(nel)
    tile_of = e_idx // T
    e_loc = e_idx - tile_of * T
    tile_ids = arange(ntiles + 1)
    # 2. tile-local vertex numbering
    nv = int(p.shape[1])
    vkey = (tile_of[:, None] * nv + tt).reshape(-1)
    uv, vinv = torch.unique(vkey, sorted=True, return_inverse=True)
    uv_tile = uv // nv
    tile_vert_start = torch.searchsorted(uv_tile, tile_ids)
    nverts_tile = tile_vert_start[1:] - tile_vert_start[:-1]
    # with renumbering (csrc/skb_p1_plan.cu) ids are bank pair + 16 * rank: the vertex
    # section of a tile gets 16 * (ceil(nv / 16) + 1) entries, unused ones stay valid
    nverts_sec = 16 * ((nverts_tile + 15) // 16 + 1) if renumber else nverts_tile
    fp.vcap = (int(nverts_sec.max()) + 1) // 2 * 2    # even: keeps shared sections 16 B aligned
    if fp.vcap >= 0xFFFF:
        raise RuntimeError("fused plan: tile touches too many vertices")
    loc = (vinv - tile_vert_start[tile_of].repeat_interleave(4)).reshape(nel, 4)
    vert_gid = uv - uv_tile * nv
    vert_tile = uv_tile
    vert_loc = arange(int(uv.shape[0])) - tile_vert_start[uv_tile]
    del vkey, uv, vinv
    # 3. CSR slot of every local entry (a, b)
    counts = (plan.indptr[1:] - plan.indptr[:-1]).long()
    row_of_slot = torch.repeat_interleave(arange(N), counts)
    cols = plan.indices.long()
    csr_key = row_of_slot * N + cols                 # ascending (canonical CSR)
    # the Laplace local matrix is bitwise symmetric, hence so are the zero
    # mask and the pattern: slot (r,c) and its mirror (c,r) receive identical
    # terms.  Only canonical slots (row <= col) are reduced; the kernel writes
    # the sum to the mirror slot too.
    mirror = torch.searchsorted(csr_key, cols * N + row_of_slot).clamp(max=max(nnz - 1, 0))
    if not bool((csr_key[mirror] == cols * N + row_of_slot).all()):
        raise FusedPlanTooBig("fused plan: CSR pattern is not structurally symmetric")
    n_canonical = int((row_of_slot <= cols).sum())
    del row_of_slot, cols
    keys2, sids = [], []
    for a in range(4):
        for b in range(a, 4):
            ra, rb = tt[:, a], tt[:, b]
            key = torch.minimum(ra, rb) * N + torch.maximum(ra, rb)
            pos = torch.searchsorted(csr_key, key).clamp(max=max(nnz - 1, 0))
            ok = csr_key[pos] == key
            k = a * 4 - (a * (a - 1)) // 2 + (b - a)
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
    if int(cnt.max()) > 0xFFFF:
        raise RuntimeError("fused plan: too many contributions to one slot in a tile")
    ts_tile = uniq // nnz
    ts_gslot = uniq - ts_tile * nnz
    kth = arange(ncontrib) - excl(cnt)[sinv]
    # Long lists (diagonal slots collect ~24 terms) are split over F = 2 or 4
    # adjacent lanes whose partial sums the kernel combines with shuffles in a
    # fixed tree; this bounds the serial add chain of a lane to ~8 terms.
    # Within a tile, slots are ordered F-major (keeps every F-block aligned to
    # F lanes), then by decreasing chunk length (sliced ELL).
    F = torch.where(cnt <= 8, 1, torch.where(cnt <= 16, 2, 4))
    chunk = (cnt + F - 1) // F
    fclass = torch.where(F == 4, 0, torch.where(F == 2, 1, 2))
    order3 = torch.argsort((ts_tile * 4 + fclass) * 65536 + (65535 - chunk), stable=True)
    newpos = torch.empty(nts, dtype=i64, device=dev)
    newpos[order3] = arange(nts)
    ts_tile, ts_gslot, cnt = ts_tile[order3], ts_gslot[order3], cnt[order3]
    F, chunk = F[order3], chunk[order3]
    tile_slot_start = torch.searchsorted(ts_tile, tile_ids)
    nslots_tile = tile_slot_start[1:] - tile_slot_start[:-1]
    Fcum = excl(F)
    lane0 = Fcum - Fcum[tile_slot_start[:-1]][ts_tile]          # leader lane slot within tile
    nlanes_tile = torch.zeros(ntiles, dtype=i64, device=dev).scatter_add_(0, ts_tile, F)
    ngroups_tile = (nlanes_tile + 31) // 32
    tile_group_start = torch.cat([torch.zeros(1, dtype=i64, device=dev),
                                  torch.cumsum(ngroups_tile, 0)])
    ngroups = int(tile_group_start[-1])
    ts_idx = arange(nts)
    j_in_tile = lane0                                            # leader position (meta, fsel)
    grp_of_slot = tile_group_start[ts_tile] + lane0 // 32        # F-blocks never straddle groups
    lane_of_slot = lane0 % 32
    grp_len = torch.zeros(ngroups, dtype=i64, device=dev)
    grp_len.scatter_reduce_(0, grp_of_slot, chunk, reduce="amax", include_self=True)
    grp_len = (grp_len + 1) // 2 * 2                 # the kernel's P2 loop is unrolled by 2
    grp_tile = torch.repeat_interleave(arange(ntiles), ngroups_tile)
    gcum = torch.cat([torch.zeros(1, dtype=i64, device=dev), torch.cumsum(grp_len * 32, 0)])
    nids_tile = gcum[tile_group_start[1:]] - gcum[tile_group_start[:-1]]   # multiples of 32
    grp_base = gcum[:-1] - gcum[tile_group_start[:-1]][grp_tile]          # tile-relative
    if int(grp_base.max()) // 32 > 0xFFFF:
        raise RuntimeError("fused plan: tile index list too long")
    ncontrib_sell = int(gcum[-1])
    # 5. slots touched by one tile go straight to csr_data, the others through scratch
    order2 = torch.argsort(ts_gslot * ntiles + ts_tile)  # by csr slot, tiles ascending
    g_sorted = ts_gslot[order2]
    ug, gcnt = torch.unique_consecutive(g_sorted, return_counts=True)
    if int(ug.shape[0]) != n_canonical:
        raise RuntimeError("fused plan: CSR pattern has slots no element contributes to")
    shared = gcnt > 1
    gsz = gcnt * shared
    gstart = excl(gsz)
    gfirst = excl(gcnt)
    grp = torch.repeat_interleave(arange(int(ug.shape[0])), gcnt)
    spos = gstart[grp] + (ts_idx - gfirst[grp])
    NONE = 0xFFFFFFFF

    def tgt(slots):                      # where a CSR slot's value is written
        return slots if slot_map is None else slot_map[slots]
    meta = torch.empty(nts, dtype=i64, device=dev)
    meta[order2] = torch.where(shared[grp], spos | 0x80000000, tgt(g_sorted))
    # mirror target: written by the tile only for exclusive off-diagonal slots
    # (shared ones are mirrored by skb_p1_combine)
    mir_sorted = mirror[g_sorted]
    meta2 = torch.empty(nts, dtype=i64, device=dev)
    meta2[order2] = torch.where(shared[grp] | (mir_sorted == g_sorted),
                                torch.full_like(g_sorted, NONE), tgt(mir_sorted))
    sh = torch.nonzero(shared).flatten()
    fp.nshared = int(sh.shape[0])
    fp.nscratch = int(gsz.sum())
    # diagonal slots mirror onto themselves: keep gslot2 == gslot after mapping
    fp.gslot = tgt(ug[sh]).to(torch.int32).contiguous()
    fp.gslot2 = tgt(mirror[ug[sh]]).to(torch.int32).contiguous()
    fp.sptr = torch.cat([gstart[sh], torch.tensor([fp.nscratch], device=dev, dtype=i64)]
                        ).to(torch.int32).contiguous()
    fp.scratch = torch.empty(max(fp.nscratch, 1), dtype=torch.float64, device=dev)
    del order2, g_sorted, grp, spos
    # 6. pack the per-tile records
    HDR = 32
    off_verts = HDR + 8 * T
    off_grp = off_verts + 4 * ((nverts_sec + 3) // 4 * 4)
    off_meta = off_grp + 16 * ((ngroups_tile + 3) // 4)
    off_meta2 = off_meta + 128 * ngroups_tile
    off_fsel = off_meta2 + 128 * ngroups_tile        # one byte per lane: log2(F) of the leader
    off_ids = off_fsel + 32 * ngroups_tile
    size = off_ids + 2 * nids_tile                   # multiple of 16
    rec_start = torch.cat([torch.zeros(1, dtype=i64, device=dev), torch.cumsum(size, 0)])
    total = int(rec_start[-1])
    fp.rec_cap = int(size.max())
    buf32 = torch.zeros(total // 4, dtype=torch.int32, device=dev)
    buf16 = buf32.view(torch.int16)
    rs = rec_start[:-1]
    hdr = torch.stack([nverts_sec, ngroups_tile, torch.full_like(rs, off_verts), off_grp,
                       off_meta, off_ids, off_fsel, off_meta2], dim=1)
    buf32[(rs // 4)[:, None] + arange(8)[None, :]] = hdr.to(torch.int32)
    # tl (padding elements of the last tile: 0xFFFF)
    pad = ntiles * T - nel
    if pad:
        last = (int(rs[-1]) + HDR) // 2 + 4 * (T - pad)
        buf16[last:last + 4 * pad] = -1
    buf16[((rs[tile_of] + HDR) // 2 + 4 * e_loc)[:, None] + arange(4)[None, :]] = \
        loc.to(torch.int16)
    # verts
    if renumber:    # every id of the (larger) section points at a valid vertex of the tile
        sec_tile = torch.repeat_interleave(arange(ntiles), nverts_sec)
        sec_pos = arange(int(nverts_sec.sum())) - excl(nverts_sec)[sec_tile]
        buf32[(rs[sec_tile] + off_verts) // 4 + sec_pos] = \
            vert_gid[tile_vert_start[:-1]][sec_tile].to(torch.int32)
        del sec_tile, sec_pos
    buf32[(rs[vert_tile] + off_verts) // 4 + vert_loc] = vert_gid.to(torch.int32)
    # grp: offset/32 | len << 16
    g_local = arange(ngroups) - tile_group_start[grp_tile]
    gword = (grp_base // 32) | (grp_len << 16)
    buf32[(rs[grp_tile] + off_grp[grp_tile]) // 4 + g_local] = gword.to(torch.int32)
    # meta: unused lanes -> 0xFFFFFFFF, then the real ones
    lanes = arange(32)
    buf32[((rs[grp_tile] + off_meta[grp_tile]) // 4 + g_local * 32)[:, None] + lanes[None, :]] = -1
    buf32[(rs[ts_tile] + off_meta[ts_tile]) // 4 + j_in_tile] = meta.to(torch.int32)
    buf32[((rs[grp_tile] + off_meta2[grp_tile]) // 4 + g_local * 32)[:, None] + lanes[None, :]] = -1
    buf32[(rs[ts_tile] + off_meta2[ts_tile]) // 4 + j_in_tile] = meta2.to(torch.int32)
    # fsel: 0 -> the lane's own sum, 1 -> pair sum, 2 -> sum of four lanes (zero elsewhere)
    buf8 = buf32.view(torch.uint8)
    fsel = torch.where(F == 4, 2, torch.where(F == 2, 1, 0))
    buf8[rs[ts_tile] + off_fsel[ts_tile] + j_in_tile] = fsel.to(torch.uint8)
    # ids: padding -> index of the staged zero, then the real contributions
    zero_idx = 10 * T
    tile_ids_start = excl(nids_tile)
    id_tile = torch.repeat_interleave(arange(ntiles), nids_tile)
    id_pos = (rs[id_tile] + off_ids[id_tile]) // 2 + (arange(ncontrib_sell) - tile_ids_start[id_tile])
    buf16[id_pos] = torch.tensor(zero_idx, dtype=i64, device=dev).to(torch.int16)
    del id_tile, id_pos
    s_new = newpos[sinv]
    g_of = grp_of_slot[s_new]
    t_of = grp_tile[g_of]
    sub = kth // chunk[s_new]                        # which lane of the slot's F-block
    cpos = ((rs[t_of] + off_ids[t_of]) // 2 + grp_base[g_of] + (kth - sub * chunk[s_new]) * 32
            + lane_of_slot[s_new] + sub)
    buf16[cpos] = sid.to(torch.int16)
    # reorder every lane's terms over the columns so that each LDS.64 of P2 is
    # spread over the shared-memory banks (csrc/skb_p1_plan.cu)
    if spread and dev.type == "cuda" and ngroups:
        grp_pos = ((rs[grp_tile] + off_ids[grp_tile]) // 2 + grp_base).contiguous()
        glen32 = grp_len.to(torch.int32).contiguous()
        code = _lib.lib().skb_p1_plan_spread(
            buf16.data_ptr(), grp_pos.data_ptr(), glen32.data_ptr(), ngroups, zero_idx,
            C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(code, "skb_p1_plan_spread")
        torch.cuda.current_stream().synchronize()     # grp_pos / glen32 die here
    fp.rec = buf32
    fp.rec_start = rec_start.contiguous()            # int64 == uint64 for the kernel
    if renumber and dev.type == "cuda":
        code = _lib.lib().skb_p1_plan_renumber(
            buf32.data_ptr(), fp.rec_start.data_ptr(), ntiles, T,
            C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(code, "skb_p1_plan_renumber")
    fp.nts, fp.ncontrib, fp.ncontrib_sell, fp.ngroups = nts, ncontrib, ncontrib_sell, ngroups
    fp.nverts_tiles = int(vert_gid.shape[0])
    fp.rec_bytes = total
    fp.w = float(basis.W[0])
    fp.nqp = int(basis.nqp)
    fp.p = p
    # coordinates all 0 or within [2^-60, 2^60]: the kernel's exact division
    # needs no per-element exponent checks (csrc/skb_p1_fused.cu, P1Args::tame)
    ap = p.abs()
    fp.tame = int(bool(((ap == 0) | ((ap >= 2.0 ** -60) & (ap <= 2.0 ** 60))).all()))
    fp.smem = int(_lib.lib().skb_p1_fused_smem_bytes(T, ring, fp.rec_cap, fp.vcap))
    if fp.smem > 227 * 1024:
        raise FusedPlanTooBig("fused plan: tile does not fit in shared memory "
                              "({} B); use a smaller tile".format(fp.smem))
    return fp


class FusedPlanTooBig(RuntimeError):
    pass


def build_auto(basis, plan, T=512, threads=480, ring=4, slot_map=None, spread=True,
               renumber=True, tiling="morton"):
    """Build with the requested tile, halving it while the tile's record ring
    and coordinates do not fit in shared memory (irregular meshes whose tiles
    touch many vertices).  Returns None if even the smallest tile is too big:
    the caller then stays on the generic path."""
    options = [(T, threads)] + [c for c in ((512, 256), (256, 256), (256, 128), (128, 96))
                                if c[0] < T]
    for tile, thr in options:
        try:
            return build(basis, plan, T=tile, threads=thr, ring=ring, slot_map=slot_map,
                         spread=spread, renumber=renumber, tiling=tiling)
        except FusedPlanTooBig:
            continue
    return None


def run(fp, data, stream, fast=False, l2_persist=False):
    """Warm numeric phase: two kernel launches, nothing else.  ``fast``: fused
    multiply-add arithmetic in the element kernel (see csrc/skb_p1_fused.cu, FAST)."""
    lib = _lib.lib()
    persist = l2_persist and fp.nscratch > 0
    if persist:      # keep the tile partials in L2 between the two kernels
        _lib.check(lib.skb_l2_window(fp.scratch.data_ptr(), 8 * fp.nscratch, stream),
                   "skb_l2_window")
    code = lib.skb_p1tet_laplace_fused(
        fp.p.data_ptr(), fp.p.shape[1], fp.rec.data_ptr(), fp.rec_start.data_ptr(), fp.ntiles,
        fp.T, fp.threads, fp.ring, fp.rec_cap, fp.vcap, 2 if fast else fp.tame, C.c_double(fp.w),
        fp.nqp,
        data.data_ptr(), fp.scratch.data_ptr(), stream)
    _lib.check(code, "skb_p1tet_laplace_fused")
    code = lib.skb_p1_combine(fp.scratch.data_ptr(), fp.sptr.data_ptr(), fp.gslot.data_ptr(),
                              fp.gslot2.data_ptr(), fp.nshared, data.data_ptr(), stream)
    _lib.check(code, "skb_p1_combine")
    if persist:
        _lib.check(lib.skb_l2_window(None, 0, stream), "skb_l2_window")


def stats(fp):
    """Bytes the fused step moves (for DESIGN.md / the roofline discussion)."""
    b = {
        "records": fp.rec_bytes, "p_gather_min": fp.nverts_tiles * 24,
        "csr_out": fp.nnz * 8, "scratch_w": fp.nscratch * 8,
        "scratch_r": fp.nscratch * 8, "sptr_gslot": fp.nshared * 12,
    }
    b["total"] = sum(b.values())
    b["per_element"] = b["total"] / max(fp.nel, 1)
    b["tile_slots_per_csr_slot"] = fp.nts / max(fp.nnz, 1)
    b["vcap"], b["rec_cap"], b["smem"] = fp.vcap, fp.rec_cap, fp.smem
    b["sell_padding"] = fp.ncontrib_sell / max(fp.ncontrib, 1)
    b["shared_slots"], b["tiling"] = fp.nshared, getattr(fp, "tiling", "morton")
    return b

"""Plan of the fused P1 Laplace path, second generation (csrc/skb_p1_fused2.cu).

Built once per (connectivity, CSR pattern) on the device, reused by every warm
re-assembly - also after the vertex coordinates changed (``run(..., p=new_p)``): the
kernel re-derives the zero mask of every local matrix and raises a flag when it differs
from the plan's, i.e. when the reference's value-dependent pattern
(skfem/assembly/form/coo_data.py:35, ``eliminate_zeros``) would no longer be this one.

Layout
------
Elements are ordered by a balanced k-d tree into *super-tiles* (compact boxes of ``S``
tiles) of *tiles* (``T`` elements).  Every tile has one contiguous 16-byte aligned
record (one TMA bulk copy in the kernel):

    header  8 x uint32: nverts, ngroups, off_verts, off_grp, off_lane, off_ids, nelems, 0
    tl      T x 4 uint16  tile-local vertex ids in the low 10 bits (0xFFFF x 4 = padding
            element); bits 10-15 of .x hold bits 0-5, of .y bits 6-9 of the element's
            *kept mask*: bit k set <=> local entry k (row-major upper triangle) is nonzero
            in the cold assembly that defined the CSR pattern
    verts   nverts int32  global vertex ids of the tile
    grp     per group of 32 lanes: uint32 (offset / 32 words into ids) | rows << 16 |
            (group contains split lists) << 31
    lane    per lane uint16: pool index (13 bits) | fsel << 13 | first-touch << 15;
            0xFFFF for lanes that own no slot (members of a split list, padding)
    ids     sliced-ELL staging positions 8 * (k(a,b)*T + e_local) (byte offsets), two per
            uint32: row r of lane l of a group at base + 32 r + l holds columns 2r (low half)
            and 2r + 1; short lists are padded with 8 * (10*T + b), the offset of a staged 0.0;
            a lane's sum is (sum of its even columns) + (sum of its odd columns)

A *tile slot* is a canonical (row <= col) CSR slot touched by the tile; the Laplace local
matrix is bitwise symmetric, so the mirror slot gets the same sum.  Every super-tile owns
a *pool* of accumulators in shared memory, one per canonical slot it touches, numbered in
CSR order.  After the last tile of a super-tile the pool is flushed through the
super-tile's *flush table* ``fl`` (uint32 pairs): slots touched by this super-tile only go
straight to ``csr_data`` (and the mirror slot), the others to ``scratch`` (grouped by CSR
slot, super-tiles ascending) and are added by ``skb_p1_combine2``.

The preprocessing uses torch sort / unique / searchsorted (cold path, plumbing); the warm
path runs only this package's kernels.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .fused import FusedPlanTooBig, _kd_order  # noqa: F401  (same exceptions / tiling)

NONE = 0xFFFFFFFF
L2_WINDOW_MAX_BYTES = 40 << 20
POOL_MAX = 8192          # 13-bit pool index


def _torch():
    import torch
    return torch


def _kd_order_nested(corner, T, S):
    """k-d order whose consecutive chunks of S*T elements are compact boxes (super-tiles)
    that are themselves cut into compact boxes of T elements."""
    torch = _torch()
    if S <= 1:
        return _kd_order(corner, T)
    outer = _kd_order(corner, S * T)
    nel = int(corner.shape[1])
    nst = (nel + S * T - 1) // (S * T)
    cnt = torch.full((nst,), S * T, dtype=torch.int64, device=corner.device)
    cnt[-1] = nel - (nst - 1) * S * T
    return _kd_order(corner, T, order=outer, seg_cnt=cnt)


class P1FusedPlan2:
    version = 2


def applicable(basis, form):
    from .fused import applicable as _a
    return _a(basis, form, version=2)


def _pick_S(n_canonical, nel, T, pool_cap):
    """Tiles per super-tile from the slot density (canonical slots per element) and a
    surface allowance; verified (and halved if necessary) by the caller."""
    r = max(n_canonical / max(nel, 1), 0.25)
    S = 1
    while S < 64 and 1.6 * r * (2 * S) * T <= pool_cap:
        S *= 2
    return S


def build(basis, plan, T=256, ring=3, pool_cap=2048, S=None, slot_map=None, spread=True,
          renumber=True, ctas_per_sm=0, defer_finalize=False, form_id=None):
    """``slot_map`` (optional int64 tensor, CSR slot -> output index): targets written by
    the kernels are remapped through it (multi-GPU direct write).  ``pool_cap``: most
    accumulators a super-tile may need (shared memory: 8 B each); ``S``: tiles per
    super-tile (default: the largest power of two whose pools fit).  ``defer_finalize``
    (tests): leave the ids in the column-major layout the plan passes work on; the caller
    applies the passes and calls :func:`finalize`."""
    torch = _torch()
    d = basis._dev()
    dev = d["device"]
    t = d["t"] if d["tind"] is None else d["t"][:, d["tind"].long()]
    p = d["p"]
    nel = int(t.shape[1])
    nnz = plan.nnz
    N = int(plan.shape[1])
    pool_cap = min(int(pool_cap), POOL_MAX) // 2 * 2
    i64 = torch.int64

    def arange(n):
        return torch.arange(n, device=dev, dtype=i64)

    def excl(x):
        return torch.cumsum(x, 0) - x

    tl = t.long()
    # CSR keys; mirror slots (the pattern of a symmetric form is structurally symmetric)
    counts = (plan.indptr[1:] - plan.indptr[:-1]).long()
    row_of_slot = torch.repeat_interleave(arange(N), counts)
    cols = plan.indices.long()
    csr_key = row_of_slot * N + cols                 # ascending (canonical CSR)
    mirror = torch.searchsorted(csr_key, cols * N + row_of_slot).clamp(max=max(nnz - 1, 0))
    if not bool((csr_key[mirror] == cols * N + row_of_slot).all()):
        raise FusedPlanTooBig("fused plan: CSR pattern is not structurally symmetric")
    n_canonical = int((row_of_slot <= cols).sum())
    del row_of_slot, cols, counts
    corner = p[:, tl].min(dim=1).values
    if S is None:
        S = _pick_S(n_canonical, nel, T, pool_cap)
    while True:
        fp = _build_with(basis, plan, T, ring, pool_cap, S, slot_map, spread, renumber, tl, p,
                         corner, csr_key, mirror, n_canonical, arange, excl, dev, defer_finalize,
                         form_id)
        if fp is not None:
            break
        if S == 1:
            raise FusedPlanTooBig("fused plan: a single tile needs more than {} accumulators"
                                  .format(pool_cap))
        S //= 2
    # the kernel is latency bound: one more resident CTA per SM is worth more than a larger
    # super-tile (fewer shared slots) - halve S while that buys a CTA
    while fp.S > 1 and _ctas_per_sm(fp.smem) < 3:
        fp2 = _build_with(basis, plan, T, ring, pool_cap, fp.S // 2, slot_map, spread, renumber,
                          tl, p, corner, csr_key, mirror, n_canonical, arange, excl, dev,
                          defer_finalize, form_id)
        if fp2 is None or _ctas_per_sm(fp2.smem) <= _ctas_per_sm(fp.smem):
            break
        fp = fp2
    fp.ctas_per_sm = int(ctas_per_sm)
    return fp


def _ctas_per_sm(smem):
    return int((227 * 1024) // (smem + 1024))


def _build_with(basis, plan, T, ring, pool_cap, S, slot_map, spread, renumber, tl, p, corner,
                csr_key, mirror, n_canonical, arange, excl, dev, defer_finalize, form_id=None):
    torch = _torch()
    i64 = torch.int64
    nel = int(tl.shape[1])
    nnz = plan.nnz
    N = int(plan.shape[1])
    ntiles = (nel + T - 1) // T
    nst = (ntiles + S - 1) // S
    # 1. element order: consecutive chunks of T elements are the tiles, of S tiles the super-tiles
    order = _kd_order_nested(corner, T, S)
    tt = tl[:, order].t().contiguous()              # (nel, 4) int64, tile order
    e_idx = arange(nel)
    tile_of = e_idx // T
    e_loc = e_idx - tile_of * T
    tile_ids = arange(ntiles + 1)
    # 2. CSR slot of every local entry (a, b); entries whose slot is absent (exact zeros of
    # every element sharing it) are left out of the lists
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
    del keys2, sids
    key2, perm = torch.sort(key2, stable=True)
    sid = sid[perm]
    del perm
    # 3. tile slots = unique (tile, csr slot) pairs, contributions grouped per slot
    uniq, sinv, cnt = torch.unique_consecutive(key2, return_inverse=True, return_counts=True)
    del key2
    nts = int(uniq.shape[0])
    ncontrib = int(sid.shape[0])
    ts_tile = uniq // nnz
    ts_gslot = uniq - ts_tile * nnz
    # 4. pools: one accumulator per (super-tile, csr slot), numbered in CSR order
    ts_st = ts_tile // S
    pkey = ts_st * nnz + ts_gslot
    pu, pinv = torch.unique(pkey, sorted=True, return_inverse=True)
    pu_st = pu // nnz
    pu_gslot = pu - pu_st * nnz
    st_ids = arange(nst + 1)
    pu_start = torch.searchsorted(pu_st, st_ids)
    npool = pu_start[1:] - pu_start[:-1]
    pool_need = int(npool.max()) if nst else 0
    if pool_need + (pool_need & 1) > pool_cap:
        return None
    pool_idx = pinv - pu_start[ts_st]                # per tile slot
    # flush tables start at even entries (16-byte aligned TMA source): pad odd pools by one
    npad = npool + (npool & 1)
    st_fl0 = torch.cat([torch.zeros(1, dtype=i64, device=dev), torch.cumsum(npad, 0)])
    fl_pos = st_fl0[pu_st] + (arange(int(pu.shape[0])) - pu_start[pu_st])
    # first touch: the smallest tile of every (super-tile, slot) group stores, later ones add
    first_tile = torch.full((int(pu.shape[0]),), ntiles, dtype=i64, device=dev)
    first_tile.scatter_reduce_(0, pinv, ts_tile, reduce="amin", include_self=True)
    first = first_tile[pinv] == ts_tile
    del pkey, first_tile
    # 5. slots touched by one super-tile go straight to csr_data, the others through scratch
    # (pu is sorted by (st, gslot); regroup by gslot, super-tiles ascending)
    o2 = torch.argsort(pu_gslot * nst + pu_st)
    g_sorted = pu_gslot[o2]
    ug, gcnt = torch.unique_consecutive(g_sorted, return_counts=True)
    if int(ug.shape[0]) != n_canonical:
        raise RuntimeError("fused plan: CSR pattern has slots no element contributes to")
    shared = gcnt > 1
    gsz = gcnt * shared
    gstart = excl(gsz)
    gfirst = excl(gcnt)
    grp_of = torch.repeat_interleave(arange(int(ug.shape[0])), gcnt)
    spos = gstart[grp_of] + (arange(int(pu.shape[0])) - gfirst[grp_of])

    def tgt(slots):                      # where a CSR slot's value is written
        return slots if slot_map is None else slot_map[slots]
    npu = int(st_fl0[-1])
    fl_m = torch.full((npu + 2,), NONE, dtype=i64, device=dev)       # padding entries: no target
    fl_m[fl_pos[o2]] = torch.where(shared[grp_of], spos | 0x80000000, tgt(g_sorted))
    mir_sorted = mirror[g_sorted]
    fl_m2 = torch.full((npu + 2,), NONE, dtype=i64, device=dev)
    fl_m2[fl_pos[o2]] = torch.where(shared[grp_of] | (mir_sorted == g_sorted),
                                    torch.full_like(g_sorted, NONE), tgt(mir_sorted))
    fp = P1FusedPlan2()
    fp.T, fp.ring, fp.nel, fp.nnz, fp.S = T, ring, nel, nnz, S
    fp.ntiles, fp.nst = ntiles, nst
    fp.pool_cap = max(pool_need + (pool_need & 1), 2)
    fp.fl = _i32(torch.stack([fl_m, fl_m2], dim=1)).contiguous()   # (npu, 2) uint32 bit patterns
    fp.st_fl0 = st_fl0.contiguous()
    fp.st_tile0 = torch.clamp(st_ids * S, max=ntiles).to(torch.int32).contiguous()
    sh = torch.nonzero(shared).flatten()
    fp.nshared = int(sh.shape[0])
    fp.nscratch = int(gsz.sum())
    fp.gslot = tgt(ug[sh]).to(torch.int32).contiguous()
    fp.gslot2 = tgt(mirror[ug[sh]]).to(torch.int32).contiguous()
    fp.sptr = torch.cat([gstart[sh], torch.tensor([fp.nscratch], device=dev, dtype=i64)]
                        ).to(torch.int32).contiguous()
    fp.scratch = torch.empty(max(fp.nscratch, 1), dtype=torch.float64, device=dev)
    fp.npool_total = npu
    del o2, g_sorted, grp_of, spos, fl_m, fl_m2, mir_sorted, pu, pu_st, pu_gslot, fl_pos
    # 6. lanes: long lists (vertex diagonals collect ~24 terms) are split over F = 2 or 4
    # adjacent lanes combined by a fixed shuffle tree; within a tile F-major, then by
    # decreasing chunk length (sliced ELL)
    if int(cnt.max()) > 0xFFFF:
        raise RuntimeError("fused plan: too many contributions to one slot in a tile")
    kth = arange(ncontrib) - excl(cnt)[sinv]
    F = torch.where(cnt <= 8, 1, torch.where(cnt <= 16, 2, 4))
    chunk = (cnt + F - 1) // F
    fclass = torch.where(F == 4, 0, torch.where(F == 2, 1, 2))
    order3 = torch.argsort((ts_tile * 4 + fclass) * 65536 + (65535 - chunk), stable=True)
    newpos = torch.empty(nts, dtype=i64, device=dev)
    newpos[order3] = arange(nts)
    ts_tile, cnt = ts_tile[order3], cnt[order3]
    pool_idx, first = pool_idx[order3], first[order3]
    F, chunk = F[order3], chunk[order3]
    tile_slot_start = torch.searchsorted(ts_tile, tile_ids)
    Fcum = excl(F)
    lane0 = Fcum - Fcum[tile_slot_start[:-1]][ts_tile]          # leader lane within the tile
    nlanes_tile = torch.zeros(ntiles, dtype=i64, device=dev).scatter_add_(0, ts_tile, F)
    ngroups_tile = (nlanes_tile + 31) // 32
    tile_group_start = torch.cat([torch.zeros(1, dtype=i64, device=dev),
                                  torch.cumsum(ngroups_tile, 0)])
    ngroups = int(tile_group_start[-1])
    grp_of_slot = tile_group_start[ts_tile] + lane0 // 32        # F-blocks never straddle groups
    lane_of_slot = lane0 % 32
    grp_len = torch.zeros(ngroups, dtype=i64, device=dev)
    grp_len.scatter_reduce_(0, grp_of_slot, chunk, reduce="amax", include_self=True)
    grp_len = (grp_len + 1) // 2 * 2                 # two columns per 32-bit word
    grp_tile = torch.repeat_interleave(arange(ntiles), ngroups_tile)
    gcum = torch.cat([torch.zeros(1, dtype=i64, device=dev), torch.cumsum(grp_len * 32, 0)])
    nids_tile = gcum[tile_group_start[1:]] - gcum[tile_group_start[:-1]]   # multiples of 64
    grp_base = gcum[:-1] - gcum[tile_group_start[:-1]][grp_tile]          # tile-relative, uint16 units
    if ngroups and int(grp_base.max()) // 64 > 0xFFFF:
        raise RuntimeError("fused plan: tile index list too long")
    ncontrib_sell = int(gcum[-1])
    # 7. tile-local vertex numbering
    nv = int(p.shape[1])
    vkey = (tile_of[:, None] * nv + tt).reshape(-1)
    uv, vinv = torch.unique(vkey, sorted=True, return_inverse=True)
    uv_tile = uv // nv
    tile_vert_start = torch.searchsorted(uv_tile, tile_ids)
    nverts_tile = tile_vert_start[1:] - tile_vert_start[:-1]
    nverts_sec = 16 * ((nverts_tile + 15) // 16 + 1) if renumber else nverts_tile
    fp.vcap = (int(nverts_sec.max()) + 1) // 2 * 2
    if fp.vcap > 1024:
        raise FusedPlanTooBig("fused plan: tile touches too many vertices")
    loc = (vinv - tile_vert_start[tile_of].repeat_interleave(4)).reshape(nel, 4)
    vert_gid = uv - uv_tile * nv
    vert_tile = uv_tile
    vert_loc = arange(int(uv.shape[0])) - tile_vert_start[uv_tile]
    del vkey, uv, vinv
    # 8. pack the per-tile records
    HDR = 32
    off_verts = HDR + 8 * T
    off_grp = off_verts + 4 * ((nverts_sec + 3) // 4 * 4)
    off_lane = off_grp + 16 * ((ngroups_tile + 3) // 4)
    off_ids = off_lane + 64 * ngroups_tile           # 32 x uint16 per group
    size = off_ids + 2 * nids_tile                   # multiple of 16
    rec_start = torch.cat([torch.zeros(1, dtype=i64, device=dev), torch.cumsum(size, 0)])
    total = int(rec_start[-1])
    fp.rec_cap = int(size.max())
    buf32 = torch.zeros(total // 4, dtype=torch.int32, device=dev)
    buf16 = buf32.view(torch.int16)
    rs = rec_start[:-1]
    nel_tile = torch.clamp(nel - arange(ntiles) * T, max=T)
    hdr = torch.stack([nverts_sec, ngroups_tile, torch.full_like(rs, off_verts), off_grp,
                       off_lane, off_ids, nel_tile, torch.zeros_like(rs)], dim=1)
    buf32[(rs // 4)[:, None] + arange(8)[None, :]] = hdr.to(torch.int32)
    pad = ntiles * T - nel
    if pad:
        last = (int(rs[-1]) + HDR) // 2 + 4 * (T - pad)
        buf16[last:last + 4 * pad] = -1
    tl_pos = ((rs[tile_of] + HDR) // 2 + 4 * e_loc)[:, None] + arange(4)[None, :]
    buf16[tl_pos] = loc.to(torch.int16)
    if renumber:    # every id of the (larger) section points at a valid vertex of the tile
        sec_tile = torch.repeat_interleave(arange(ntiles), nverts_sec)
        sec_pos = arange(int(nverts_sec.sum())) - excl(nverts_sec)[sec_tile]
        buf32[(rs[sec_tile] + off_verts) // 4 + sec_pos] = \
            vert_gid[tile_vert_start[:-1]][sec_tile].to(torch.int32)
        del sec_tile, sec_pos
    buf32[(rs[vert_tile] + off_verts) // 4 + vert_loc] = vert_gid.to(torch.int32)
    g_local = arange(ngroups) - tile_group_start[grp_tile]
    # bit 31: the group holds split lists (F > 1), i.e. the kernel must run its shuffle tree
    grp_split = torch.zeros(ngroups, dtype=i64, device=dev)
    grp_split.scatter_reduce_(0, grp_of_slot, (F > 1).long(), reduce="amax", include_self=True)
    gword = (grp_base // 64) | ((grp_len // 2) << 16) | (grp_split << 31)
    buf32[(rs[grp_tile] + off_grp[grp_tile]) // 4 + g_local] = _i32(gword)
    # lane words: unused lanes 0xFFFF, leaders pool | fsel << 13 | first << 15
    lanes = arange(32)
    buf16[((rs[grp_tile] + off_lane[grp_tile]) // 2 + g_local * 32)[:, None] + lanes[None, :]] = -1
    fsel = torch.where(F == 4, 2, torch.where(F == 2, 1, 0))
    lword = pool_idx | (fsel << 13) | (first.long() << 15)
    buf16[(rs[ts_tile] + off_lane[ts_tile]) // 2 + lane0] = _i16(lword)
    # ids in column-major sliced-ELL order first (cell (c, l) of a group at base + 32 c + l):
    # the layout the bank-spreading pass works on
    zero_idx = 10 * T
    ell = torch.full((max(ncontrib_sell, 1),), zero_idx, dtype=torch.int16, device=dev)
    s_new = newpos[sinv]
    g_of = grp_of_slot[s_new]
    sub = kth // chunk[s_new]                        # which lane of the slot's F-block
    cpos = gcum[:-1][g_of] + (kth - sub * chunk[s_new]) * 32 + lane_of_slot[s_new] + sub
    ell[cpos] = sid.to(torch.int16)
    del s_new, g_of, sub, cpos
    if spread and dev.type == "cuda" and ngroups:
        grp_pos = gcum[:-1].contiguous()
        glen32 = grp_len.to(torch.int32).contiguous()
        code = _lib.lib().skb_p1_plan_spread(
            ell.data_ptr(), grp_pos.data_ptr(), glen32.data_ptr(), ngroups, zero_idx,
            C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(code, "skb_p1_plan_spread")
        torch.cuda.current_stream().synchronize()
    fp._ell, fp._gcum, fp._grp_len = ell, gcum, grp_len      # kept for the CPU plan-pass tests
    fp._ids_base16 = ((rs[grp_tile] + off_ids[grp_tile]) // 2 + grp_base) if ngroups else None
    fp.rec = buf32
    fp.rec_start = rec_start.contiguous()            # int64 == uint64 for the kernel
    fp._tl_pos, fp._order = tl_pos, order
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
    fp.ctas_per_sm = 0
    fp.flag = torch.zeros(1, dtype=torch.int32, device=dev)
    fp.form_id = _lib.FORM_LAPLACE if form_id is None else int(form_id)
    if fp.form_id == _lib.FORM_MASS:
        # phi[4][4] (basis function x quadrature point) then W[4], host doubles
        fp.tab = (C.c_double * 20)(*np.concatenate(
            [np.asarray(basis._phi, dtype=np.float64).reshape(-1)[:16],
             np.asarray(basis.W, dtype=np.float64)[:4]]))
        fp.mode = 4
    else:
        fp.tab = None
        fp.mode = arithmetic_mode(p, fp.w, fp.nqp)
    fp.smem = int(_lib.lib().skb_p1_fused2_smem_bytes(T, ring, fp.rec_cap, fp.vcap, fp.pool_cap))
    if fp.smem > 227 * 1024:
        raise FusedPlanTooBig("fused plan: tile does not fit in shared memory "
                              "({} B); use a smaller tile".format(fp.smem))
    if not defer_finalize:
        finalize(fp)
    return fp


def finalize(fp, nz=None):
    """Last steps of the record packing, after the in-place plan passes: the ELL ids go into
    the records two per 32-bit word, and the zero mask of every element's local matrix -
    ``nz`` (int tensor, original element order, bit k <=> entry k nonzero) or, by default,
    measured by one mask pass of the kernel itself on the plan's coordinates - goes into the
    spare bits of ``tl``."""
    torch = _torch()
    buf16 = fp.rec.view(torch.int16)
    dev = fp.rec.device
    if fp._ids_base16 is not None:
        ell, gcum = fp._ell, fp._gcum
        n = int(gcum[-1])
        j = torch.arange(n, device=dev, dtype=torch.int64)
        g = torch.searchsorted(gcum, j, right=True) - 1
        o = j - gcum[g]
        c, lane = o // 32, o % 32
        # stored as byte offsets into the staging array (index * 8 < 65536 for T <= 512)
        buf16[fp._ids_base16[g] + ((c // 2) * 32 + lane) * 2 + (c & 1)] = \
            _i16((ell[:n].long() & 0xFFFF) * 8)
        del j, g, o, c, lane
    fp._ell = fp._gcum = fp._grp_len = fp._ids_base16 = None
    if nz is None:
        out = torch.zeros(fp.ntiles * fp.T, dtype=torch.int16, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _launch(fp, fp.p, fp.scratch, stream, fp.mode, out.data_ptr())
        kept = (out[:fp.nel].long() & 0x3FF)
        fp.flag.zero_()
    else:
        kept = nz.to(dev).long()[fp._order] & 0x3FF
    bits = torch.stack([(kept & 0x3F) << 10, ((kept >> 6) & 0xF) << 10], dim=1)
    pos = fp._tl_pos[:, :2]
    cur = buf16[pos].long() & 0xFFFF
    buf16[pos] = _i16((cur & 0x3FF) | bits)
    fp._tl_pos = fp._order = None


def _i32(x):
    """int64 in [0, 2^32) -> the int32 with the same bit pattern."""
    return ((x + 0x80000000) % 0x100000000 - 0x80000000).to(_torch().int32)


def _i16(x):
    """int64 in [0, 65535] -> the int16 with the same bit pattern."""
    return ((x + 0x8000) % 0x10000 - 0x8000).to(_torch().int16)


def arithmetic_mode(p, w, nqp):
    """Kernel arithmetic variant (csrc/skb_p1_fused2.cu, MODE) the coordinates allow."""
    ap = p.abs()
    nz = ap != 0

    def within(r):
        return bool((~nz | ((ap >= 2.0 ** -r) & (ap <= 2.0 ** r))).all())
    if nqp != 4 or not within(60):
        return 0
    if within(28) and 2.0 ** -20 <= w <= 1.0:
        return 2
    return 1


def build_auto(basis, plan, T=256, ring=3, pool_cap=2048, S=None, slot_map=None, spread=True,
               renumber=True, ctas_per_sm=0, form_id=None):
    """Build with the requested tile, halving it while it does not fit in shared memory
    (irregular meshes whose tiles touch many vertices).  Returns None if even the smallest
    tile is too big: the caller then stays on the generic path."""
    for tile in [T] + [c for c in (256, 128) if c < T]:
        try:
            return build(basis, plan, T=tile, ring=ring, pool_cap=pool_cap, S=S,
                         slot_map=slot_map, spread=spread, renumber=renumber,
                         ctas_per_sm=ctas_per_sm, form_id=form_id)
        except FusedPlanTooBig:
            continue
    return None


def _launch(fp, p, data, stream, mode, nz_out=None):
    if getattr(fp, "form_id", _lib.FORM_LAPLACE) == _lib.FORM_MASS:
        code = _lib.lib().skb_p1tet_mass_fused2(
            fp.tab, p.data_ptr(), p.shape[1], fp.rec.data_ptr(), fp.rec_start.data_ptr(),
            fp.st_fl0.data_ptr(), fp.fl.data_ptr(), fp.nst, fp.ntiles, fp.S, fp.T, fp.ring,
            fp.rec_cap, fp.vcap, fp.pool_cap, fp.ctas_per_sm & 0xff, data.data_ptr(),
            fp.scratch.data_ptr(), fp.flag.data_ptr(), nz_out, stream)
        _lib.check(code, "skb_p1tet_mass_fused2")
        return
    code = _lib.lib().skb_p1tet_laplace_fused2(
        p.data_ptr(), p.shape[1], fp.rec.data_ptr(), fp.rec_start.data_ptr(),
        fp.st_fl0.data_ptr(), fp.fl.data_ptr(), fp.nst, fp.ntiles, fp.S, fp.T, fp.ring,
        fp.rec_cap, fp.vcap, fp.pool_cap, fp.ctas_per_sm, mode, C.c_double(fp.w), fp.nqp,
        data.data_ptr(), fp.scratch.data_ptr(), fp.flag.data_ptr(), nz_out, stream)
    _lib.check(code, "skb_p1tet_laplace_fused2")


def run(fp, data, stream, fast=False, p=None, l2_persist=True):
    """Warm numeric phase: two kernel launches, nothing else.  ``p``: vertex coordinates to
    assemble with (default: the ones the plan was built from; same shape, same device; the
    caller vouches that they stay within the range ``fp.mode`` was chosen for, see
    :func:`arithmetic_mode`)."""
    lib = _lib.lib()
    mass = getattr(fp, "form_id", _lib.FORM_LAPLACE) == _lib.FORM_MASS
    pp = fp.p if p is None else p
    # the partials of the shared slots are written once by the fused kernel and read once by the
    # combine kernel right after it: an L2 persisting window keeps that round trip out of HBM
    # (measured -1.2 % of the step; the same window on the vertex coordinates costs +1.2 %)
    # ... as long as the partials are a small part of the 126 MB L2: a window over hundreds of
    # MB (the 47 M-element parts of BASELINE configs[4] on 2 GPUs: 236 MB) sets most of the L2
    # aside and starves the streamed records - measured 2.86 instead of 1.47 ms per step
    persist = l2_persist and 0 < 8 * fp.nscratch <= L2_WINDOW_MAX_BYTES
    if persist:
        _lib.check(lib.skb_l2_window(fp.scratch.data_ptr(), 8 * fp.nscratch, stream),
                   "skb_l2_window")
    _launch(fp, pp, data, stream, 3 if (fast and not mass) else fp.mode)
    code = lib.skb_p1_combine2(fp.scratch.data_ptr(), fp.sptr.data_ptr(), fp.gslot.data_ptr(),
                               fp.gslot2.data_ptr(), fp.nshared, data.data_ptr(), stream)
    _lib.check(code, "skb_p1_combine2")
    if persist:
        _lib.check(lib.skb_l2_window(None, 0, stream), "skb_l2_window")


def pattern_changed(fp, reset=True):
    """True if a warm run since the last check saw an element whose zero mask differs from
    the plan's (synchronises the stream)."""
    changed = bool(int(fp.flag.item()) & 1)
    if changed and reset:
        fp.flag.zero_()
    return changed


def stats(fp):
    """Bytes the fused step moves (for DESIGN.md / the roofline discussion)."""
    b = {
        "records": fp.rec_bytes, "p_gather_min": fp.nverts_tiles * 24,
        "csr_out": fp.nnz * 8, "flush_table": fp.npool_total * 8,
        "scratch_w": fp.nscratch * 8, "scratch_r": fp.nscratch * 8,
        "sptr_gslot": fp.nshared * 12,
    }
    b["total"] = sum(b.values())
    b["per_element"] = b["total"] / max(fp.nel, 1)
    b["tile_slots_per_csr_slot"] = fp.nts / max(fp.nnz, 1)
    b["vcap"], b["rec_cap"], b["smem"], b["pool_cap"] = fp.vcap, fp.rec_cap, fp.smem, fp.pool_cap
    b["sell_padding"] = fp.ncontrib_sell / max(fp.ncontrib, 1)
    b["shared_slots"], b["super_tile_tiles"], b["mode"] = fp.nshared, fp.S, fp.mode
    b["tile"], b["ring"] = fp.T, fp.ring
    return b

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


def build(basis, plan, T=1024):
    torch = _torch()
    d = basis._dev()
    dev = d["device"]
    t = d["t"] if d["tind"] is None else d["t"][:, d["tind"].long()]
    p = d["p"]
    nel = int(t.shape[1])
    nnz = plan.nnz
    N = int(plan.shape[1])
    fp = P1FusedPlan()
    fp.T, fp.nel, fp.nnz = T, nel, nnz
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
    pad = ntiles * T - nel
    if pad:
        tt_pad = torch.cat([tt, torch.full((pad, 4), -1, dtype=i64, device=dev)])
    else:
        tt_pad = tt
    fp.tt = tt_pad.to(torch.int32).contiguous()
    # 2. CSR slot of every local entry (a, b)
    counts = (plan.indptr[1:] - plan.indptr[:-1]).long()
    row_of_slot = torch.repeat_interleave(torch.arange(N, device=dev, dtype=i64), counts)
    csr_key = row_of_slot * N + plan.indices.long()  # ascending (canonical CSR)
    del row_of_slot
    e_idx = torch.arange(nel, device=dev, dtype=i64)
    tile_of = e_idx // T
    e_loc = e_idx - tile_of * T
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
    del keys2, sids
    key2, perm = torch.sort(key2, stable=True)
    sid = sid[perm]
    del perm
    # 3. tile slots = unique (tile, csr slot) pairs
    uniq, cnt = torch.unique_consecutive(key2, return_counts=True)
    del key2
    nts = int(uniq.shape[0])
    ts_tile = uniq // nnz
    ts_gslot = uniq - ts_tile * nnz
    cs = torch.cumsum(cnt, 0) - cnt                  # first contributor of each tile slot
    ncontrib = int(sid.shape[0])
    tile_ids = torch.arange(ntiles + 1, device=dev, dtype=i64)
    tile_slot_start = torch.searchsorted(ts_tile, tile_ids)
    cs_ext = torch.cat([cs, torch.tensor([ncontrib], device=dev, dtype=i64)])
    tile_contrib_start = cs_ext[tile_slot_start]
    slot_ptr = torch.empty(nts + ntiles, dtype=i64, device=dev)
    ts_idx = torch.arange(nts, device=dev, dtype=i64)
    slot_ptr[ts_idx + ts_tile] = cs - tile_contrib_start[ts_tile]
    slot_ptr[tile_slot_start[1:] + tile_ids[:-1]] = (tile_contrib_start[1:]
                                                     - tile_contrib_start[:-1])
    assert int((tile_contrib_start[1:] - tile_contrib_start[:-1]).max()) <= 65535
    fp.tile_slot_start = tile_slot_start.to(torch.int32).contiguous()
    fp.tile_contrib_start = tile_contrib_start.to(torch.int32).contiguous()
    fp.slot_ptr = slot_ptr.to(torch.int16).contiguous()
    fp.contrib = sid.to(torch.int16).contiguous()
    del sid, slot_ptr
    # 4. slots touched by one tile go straight to csr_data, the others through scratch
    order2 = torch.argsort(ts_gslot, stable=True)    # groups by csr slot, tiles ascending
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
    fp.nts, fp.ncontrib = nts, ncontrib
    fp.w = float(basis.W[0])
    fp.nqp = int(basis.nqp)
    fp.p = p
    return fp


def run(fp, data, stream):
    """Warm numeric phase: two kernel launches, nothing else."""
    lib = _lib.lib()
    code = lib.skb_p1tet_laplace_fused(
        fp.p.data_ptr(), fp.p.shape[1], fp.tt.data_ptr(), fp.ntiles, fp.T,
        fp.tile_slot_start.data_ptr(), fp.tile_contrib_start.data_ptr(),
        fp.slot_ptr.data_ptr(), fp.contrib.data_ptr(), fp.meta.data_ptr(),
        C.c_double(fp.w), fp.nqp, data.data_ptr(), fp.scratch.data_ptr(), stream)
    _lib.check(code, "skb_p1tet_laplace_fused")
    code = lib.skb_p1_combine(fp.scratch.data_ptr(), fp.sptr.data_ptr(), fp.gslot.data_ptr(),
                              fp.nshared, data.data_ptr(), stream)
    _lib.check(code, "skb_p1_combine")


def stats(fp):
    """Bytes the fused step moves (for DESIGN.md / the roofline discussion)."""
    b = {
        "tt": fp.ntiles * fp.T * 16, "contrib": fp.ncontrib * 2,
        "slot_ptr": (fp.nts + fp.ntiles) * 2, "meta": fp.nts * 4,
        "direct_out": (fp.nnz - fp.nshared) * 8, "scratch_w": fp.nscratch * 8,
        "scratch_r": fp.nscratch * 8, "sptr_gslot": fp.nshared * 8, "combine_out": fp.nshared * 8,
    }
    b["total"] = sum(b.values())
    b["per_element"] = b["total"] / max(fp.nel, 1)
    b["tile_slots_per_csr_slot"] = fp.nts / max(fp.nnz, 1)
    return b

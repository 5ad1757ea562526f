"""CPU check of the second-generation fused-path plan (skfem_b200/fused2.py).

The plan builder is torch-only, so it runs on CPU tensors; this test decodes the per-tile
records and the per-super-tile flush tables exactly as csrc/skb_p1_fused2.cu does and emulates
the kernel in numpy

    P1     local 4x4 Laplace matrix per element from tl / verts / p, the 10 symmetric entries
           staged at vals[k*T + e]; the element's zero mask against the mask bits of tl
    P2     one lane per tile slot: sum of the staged values listed in its sliced-ELL rows (two
           indices per 32-bit word), split lists combined by the shuffle tree, result stored
           (first touch) or added to the super-tile's pool
    flush  after the last tile of a super-tile: pool -> csr_data / mirror / scratch
    skb_p1_combine2   per shared slot, the partials in super-tile order

and compares the assembled values with the oracle's CSR.  The two device-only plan passes
(csrc/skb_p1_plan.cu) are applied from the same source compiled by g++."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from cases import mesh_of  # noqa: F401  (conftest puts tests/ and the package on sys.path)

NONE = 0xFFFFFFFF


def _ensure_lib():
    import os
    from skfem_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()


def _plan_on_cpu(m):
    from oracle import skfem_oracle as O
    b = O.cell_basis(m, O.element("tet_p1"))
    idx, data, shape = O.bilinear_coo(O.laplace, b)
    A = O.coo_to_csr(idx, data, shape)
    A.sort_indices()
    nel = m.t.shape[1]
    loc = data.reshape(4, 4, nel)                     # [j][i][e], bitwise symmetric
    nz = np.zeros(nel, dtype=np.int64)
    k = 0
    for a in range(4):
        for c in range(a, 4):
            nz |= (loc[a, c] != 0).astype(np.int64) << k
            k += 1
    plan = SimpleNamespace(indptr=torch.from_numpy(A.indptr.astype(np.int32)),
                           indices=torch.from_numpy(A.indices.astype(np.int32)),
                           nnz=int(A.nnz), shape=A.shape)
    dev = {"device": torch.device("cpu"), "t": torch.from_numpy(m.t.astype(np.int32)),
           "tind": None, "p": torch.from_numpy(np.ascontiguousarray(m.p))}
    basis = SimpleNamespace(_dev=lambda: dev, W=b.W, nqp=b.W.shape[0])
    return basis, plan, A, torch.from_numpy(nz)


def _local_p1_laplace(X, w, nqp):
    A = X[:, 1:] - X[:, :1]
    inv = np.linalg.inv(A)
    g = np.vstack([-inv.sum(axis=0), inv])
    return (g @ g.T) * abs(np.linalg.det(A)) * w * nqp


def _emulate(fp, p):
    T = fp.T
    rec = fp.rec.numpy()
    rec16, rs = rec.view(np.uint16), fp.rec_start.numpy()
    fl = fp.fl.numpy().view(np.uint32)
    st_tile0, st_fl0 = fp.st_tile0.numpy(), fp.st_fl0.numpy()
    csr = np.full(fp.nnz, np.nan)
    scratch = np.full(max(fp.nscratch, 1), np.nan)
    sym = {}
    k = 0
    for a in range(4):
        for b in range(a, 4):
            sym[(a, b)] = k
            k += 1
    nelems = 0
    for st in range(fp.nst):
        pool = np.full(fp.pool_cap, np.nan)
        for tile in range(st_tile0[st], st_tile0[st + 1]):
            base = int(rs[tile])
            nverts, ngroups, off_verts, off_grp, off_lane, off_ids, nel_t, _ = (
                int(v) & 0xFFFFFFFF for v in rec[base // 4: base // 4 + 8])
            assert nverts <= fp.vcap and int(rs[tile + 1]) - base <= fp.rec_cap
            tl = rec16[(base + 32) // 2: (base + 32) // 2 + 4 * T].reshape(T, 4)
            verts = rec[(base + off_verts) // 4: (base + off_verts) // 4 + nverts]
            # ---- P1 ----
            vals = np.zeros(10 * T + 16)
            for e in range(T):
                if tl[e, 0] == 0xFFFF:
                    assert e >= nel_t
                    continue
                nelems += 1
                ids = (tl[e] & 0x3FF).astype(np.int64)
                keep = int(tl[e, 0] >> 10) | (int(tl[e, 1] >> 10) << 6)
                assert (tl[e, 2] >> 10) == 0 and (tl[e, 3] >> 10) == 0
                loc = _local_p1_laplace(p[:, verts[ids]], fp.w, fp.nqp)
                tol = 1e-9 * np.abs(loc).max()
                for (a, b), kk in sym.items():
                    vals[kk * T + e] = loc[a, b]
                    assert bool(keep >> kk & 1) == (abs(loc[a, b]) > tol)
            # ---- P2 ----
            grp = rec[(base + off_grp) // 4: (base + off_grp) // 4 + ngroups].view(np.uint32)
            lanew = rec16[(base + off_lane) // 2: (base + off_lane) // 2 + 32 * ngroups]
            ids0 = (base + off_ids) // 4
            for g in range(ngroups):
                rows, off = int(grp[g] >> 16) & 0x7FFF, int(grp[g] & 0xFFFF)
                split = bool(grp[g] >> 31)
                w2 = rec[ids0 + off * 32: ids0 + (off + rows) * 32].view(np.uint32).reshape(rows, 32)
                s0, s1 = np.zeros(32), np.zeros(32)
                for r in range(rows):
                    lo, hi = (w2[r] & 0xFFFF).astype(np.int64), (w2[r] >> 16).astype(np.int64)
                    assert (lo % 8 == 0).all() and (hi % 8 == 0).all()   # byte offsets
                    assert lo.max() < 8 * (10 * T + 16) and hi.max() < 8 * (10 * T + 16)
                    s0, s1 = s0 + vals[lo // 8], s1 + vals[hi // 8]
                acc = s0 + s1
                t1 = acc + np.r_[acc[1:], acc[-1:]]
                t2 = t1 + np.r_[t1[2:], t1[-2:]]
                for lane in range(32):
                    lw = int(lanew[g * 32 + lane])
                    if lw == 0xFFFF:
                        continue
                    fs = (lw >> 13) & 3
                    assert split or fs == 0
                    res = acc[lane] if fs == 0 else (t1[lane] if fs == 1 else t2[lane])
                    pi = lw & 0x1FFF
                    assert pi < st_fl0[st + 1] - st_fl0[st]
                    if lw & 0x8000:
                        assert np.isnan(pool[pi])
                        pool[pi] = res
                    else:
                        assert not np.isnan(pool[pi])
                        pool[pi] = pool[pi] + res
        # ---- flush ----
        assert st_fl0[st] % 2 == 0 and st_fl0[st + 1] - st_fl0[st] <= fp.pool_cap
        for i in range(st_fl0[st + 1] - st_fl0[st]):
            m, m2 = int(fl[st_fl0[st] + i, 0]), int(fl[st_fl0[st] + i, 1])
            if m == NONE:                              # padding entry of an odd pool
                assert m2 == NONE and np.isnan(pool[i])
                continue
            assert not np.isnan(pool[i])
            if m & 0x80000000:
                assert np.isnan(scratch[m & 0x7FFFFFFF])
                scratch[m & 0x7FFFFFFF] = pool[i]
            else:
                assert np.isnan(csr[m])
                csr[m] = pool[i]
            if m2 != NONE:
                assert np.isnan(csr[m2])
                csr[m2] = pool[i]
    assert nelems == fp.nel
    # ---- skb_p1_combine2 ----
    sptr, gslot, gslot2 = (x.numpy().view(np.uint32) for x in (fp.sptr, fp.gslot, fp.gslot2))
    for k in range(fp.nshared):
        acc = 0.0
        for i in range(int(sptr[k]), int(sptr[k + 1])):
            acc = acc + scratch[i]
        assert np.isnan(csr[gslot[k]])
        csr[gslot[k]] = acc
        if gslot2[k] != gslot[k]:
            assert np.isnan(csr[gslot2[k]])
            csr[gslot2[k]] = acc
    assert not np.isnan(scratch[:fp.nscratch]).any()
    return csr


def _mesh(kind):
    import skfem_b200 as fem
    from oracle import skfem_oracle as O
    if kind == "ball":
        ball = fem.MeshTet.init_ball(2)
        return mesh_of(dict(p=ball.p, t=ball.t), "tet")
    m = O.mesh_tet_tensor(np.linspace(0, 1, 7), np.linspace(0, 1, 6), np.linspace(0, 1, 5))
    if kind == "morphed":    # unstructured geometry: no exact zeros, full 15-point pattern
        q = m.p.copy()
        q[0] = m.p[0] + 0.03 * np.sin(7 * m.p[1])
        q[1] = m.p[1] + 0.02 * m.p[2] ** 2
        m = mesh_of(dict(p=q, t=m.t), "tet")
    return m


@pytest.mark.parametrize("kind,T,S", [("kuhn", 128, 1), ("kuhn", 128, 4), ("morphed", 256, 2),
                                      ("morphed", 128, None), ("ball", 128, None),
                                      ("ball", 128, 2)])
def test_fused2_plan_reproduces_the_csr(kind, T, S):
    from skfem_b200 import fused2
    _ensure_lib()
    m = _mesh(kind)
    basis, plan, A, nz = _plan_on_cpu(m)
    fp = fused2.build(basis, plan, T=T, S=S, pool_cap=4096, defer_finalize=True)
    fused2.finalize(fp, nz=nz)
    assert fp.ntiles == -(-m.t.shape[1] // T) and fp.rec_cap % 16 == 0
    assert (fp.rec_start.numpy() % 16 == 0).all()
    assert fp.nst == -(-fp.ntiles // fp.S)
    csr = _emulate(fp, m.p)
    assert not np.isnan(csr).any()                    # every CSR slot written exactly once
    np.testing.assert_allclose(csr, A.data, rtol=1e-11, atol=1e-12 * np.abs(A.data).max())
    st = fused2.stats(fp)
    assert st["sell_padding"] >= 1.0 and st["tile_slots_per_csr_slot"] >= 0.5


def test_fused2_pool_capacity_limits_the_super_tile():
    """A pool too small for the requested super-tile halves S; one too small for a single
    tile is reported as FusedPlanTooBig (the caller then tries a smaller tile)."""
    from skfem_b200 import fused2
    _ensure_lib()
    m = _mesh("morphed")
    basis, plan, A, nz = _plan_on_cpu(m)
    big = fused2.build(basis, plan, T=128, S=4, pool_cap=4096, defer_finalize=True)
    small = fused2.build(basis, plan, T=128, S=4, pool_cap=big.pool_cap - 2, defer_finalize=True)
    assert big.S == 4 and small.S < 4 and small.pool_cap <= big.pool_cap - 2
    fused2.finalize(small, nz=nz)
    csr = _emulate(small, m.p)
    np.testing.assert_allclose(csr, A.data, rtol=1e-11, atol=1e-12 * np.abs(A.data).max())
    with pytest.raises(fused2.FusedPlanTooBig):
        fused2.build(basis, plan, T=128, S=1, pool_cap=16, defer_finalize=True)


@pytest.mark.parametrize("kind", ["kuhn", "ball"])
def test_fused2_plan_passes_keep_the_plan_correct(kind):
    """Bank spreading runs on the column-major ELL array, vertex renumbering on the records,
    both from the shipped source compiled for the host; the plan must stay a correct plan."""
    import host_plan_passes
    from skfem_b200 import fused2
    _ensure_lib()
    m = _mesh(kind)
    basis, plan, A, nz = _plan_on_cpu(m)
    T = 128
    fp = fused2.build(basis, plan, T=T, pool_cap=4096, defer_finalize=True)
    before_rec, before_ell = fp.rec.numpy().copy(), fp._ell.numpy().copy()
    host_plan_passes.apply2(fp, T)
    assert (before_rec != fp.rec.numpy()).any() and (before_ell != fp._ell.numpy()).any()
    fused2.finalize(fp, nz=nz)
    csr = _emulate(fp, m.p)
    assert not np.isnan(csr).any()
    np.testing.assert_allclose(csr, A.data, rtol=1e-11, atol=1e-12 * np.abs(A.data).max())


def test_arithmetic_mode_follows_the_coordinate_range():
    from skfem_b200 import fused2
    p = torch.tensor([[0.0, 1.0, 0.5], [0.25, 2.0, 0.0], [1e-3, 3.0, 7.0]], dtype=torch.float64)
    w = 0.25 / 6
    assert fused2.arithmetic_mode(p, w, 4) == 2
    assert fused2.arithmetic_mode(p * 2.0 ** 40, w, 4) == 1      # outside [2^-28, 2^28]
    assert fused2.arithmetic_mode(p * 2.0 ** -100, w, 4) == 0    # outside the exact_div range
    assert fused2.arithmetic_mode(p, w, 5) == 0                  # not the 4-point rule
    assert fused2.arithmetic_mode(p, 2.0, 4) == 1                # weight outside [2^-20, 1]

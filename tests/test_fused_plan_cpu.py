"""CPU check of the fused-path plan (skfem_b200/fused.py): the plan builder is
torch-only, so it runs on CPU tensors; this test decodes the per-tile records
exactly as csrc/skb_p1_fused.cu does and emulates the kernel's two phases in numpy

    P1   local 4x4 Laplace matrix per element from tl / verts / p, the 10 symmetric
         entries staged at vals[k*T + e]
    P2   one lane per tile slot: sum of the staged values listed in its sliced-ELL
         column, split lists combined by the shuffle tree, result written to the
         CSR slot / its mirror / the scratch position named by meta / meta2
    skb_p1_combine   per shared slot, the partials in tile order

and compares the assembled values with the oracle's CSR.  The two device-only plan
passes - bank spreading and vertex renumbering, csrc/skb_p1_plan.cu - permute this
layout in place; their kernels are scalar per-thread code, so tests/host_plan_passes.py
compiles the same source with g++ and the last test here checks that a plan is still
a correct plan after them (and that they did what they are for)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from cases import mesh_of  # noqa: F401  (conftest puts tests/ and the package on sys.path)

NONE = 0xFFFFFFFF


def _plan_on_cpu(m):
    from oracle import skfem_oracle as O
    b = O.cell_basis(m, O.element("tet_p1"))
    A = O.assemble_bilinear(O.laplace, b)
    A.sort_indices()
    plan = SimpleNamespace(indptr=torch.from_numpy(A.indptr.astype(np.int32)),
                           indices=torch.from_numpy(A.indices.astype(np.int32)),
                           nnz=int(A.nnz), shape=A.shape)
    dev = {"device": torch.device("cpu"), "t": torch.from_numpy(m.t.astype(np.int32)),
           "tind": None, "p": torch.from_numpy(np.ascontiguousarray(m.p))}
    basis = SimpleNamespace(_dev=lambda: dev, W=b.W, nqp=b.W.shape[0])
    return basis, plan, A


def _local_p1_laplace(X, w, nqp):
    """4x4 local matrix of one tetrahedron with vertex coordinates X (3, 4)."""
    A = X[:, 1:] - X[:, :1]
    inv = np.linalg.inv(A)
    g = np.vstack([-inv.sum(axis=0), inv])            # gradients of the 4 hat functions
    return (g @ g.T) * abs(np.linalg.det(A)) * w * nqp


def _emulate(fp, p, T, host_combine=False):
    rec = fp.rec.numpy()
    rec16, rec8 = rec.view(np.uint16), rec.view(np.uint8)
    rs = fp.rec_start.numpy()
    csr = np.full(fp.nnz, np.nan)
    scratch = np.full(max(fp.nscratch, 1), np.nan)
    sym = {}
    k = 0
    for a in range(4):
        for b in range(a, 4):
            sym[(a, b)] = k
            k += 1
    for tile in range(fp.ntiles):
        base = int(rs[tile])
        nverts, ngroups, off_verts, off_grp, off_meta, off_ids, off_fsel, off_meta2 = (
            int(v) & 0xFFFFFFFF for v in rec[base // 4: base // 4 + 8])
        tl = rec16[(base + 32) // 2: (base + 32) // 2 + 4 * T].reshape(T, 4)
        verts = rec[(base + off_verts) // 4: (base + off_verts) // 4 + nverts]
        assert nverts <= fp.vcap
        # ---- P1 ----
        vals = np.zeros(10 * T + 16)
        for e in range(T):
            if tl[e, 0] == 0xFFFF:
                continue
            loc = _local_p1_laplace(p[:, verts[tl[e].astype(np.int64)]], fp.w, fp.nqp)
            for (a, b), kk in sym.items():
                vals[kk * T + e] = loc[a, b]
        # ---- P2 ----
        grp = rec[(base + off_grp) // 4: (base + off_grp) // 4 + ngroups].view(np.uint32)
        meta = rec[(base + off_meta) // 4: (base + off_meta) // 4 + 32 * ngroups].view(np.uint32)
        meta2 = rec[(base + off_meta2) // 4: (base + off_meta2) // 4 + 32 * ngroups].view(np.uint32)
        fsel = rec8[base + off_fsel: base + off_fsel + 32 * ngroups]
        ids0 = (base + off_ids) // 2
        for g in range(ngroups):
            ln, off = int(grp[g] >> 16), int(grp[g] & 0xFFFF)
            cols = rec16[ids0 + off * 32: ids0 + (off + ln) * 32].reshape(ln, 32).astype(np.int64)
            assert ln % 2 == 0 and cols.max() < 10 * T + 16
            acc = vals[cols].sum(axis=0)
            down1 = np.r_[acc[1:], acc[-1:]]
            t1 = acc + down1
            t2 = t1 + np.r_[t1[2:], t1[-2:]]
            for lane in range(32):
                fs = int(fsel[g * 32 + lane])
                res = acc[lane] if fs == 0 else (t1[lane] if fs == 1 else t2[lane])
                m, m2 = int(meta[g * 32 + lane]), int(meta2[g * 32 + lane])
                if m != NONE:
                    if m & 0x80000000:
                        assert np.isnan(scratch[m & 0x7FFFFFFF])
                        scratch[m & 0x7FFFFFFF] = res
                    else:
                        assert np.isnan(csr[m])
                        csr[m] = res
                if m2 != NONE:
                    assert np.isnan(csr[m2])
                    csr[m2] = res
    # ---- skb_p1_combine ----
    if host_combine:         # the real p1_combine_kernel source, compiled for the host
        import host_plan_passes
        csr_k = host_plan_passes.combine(fp, scratch.copy(), csr.copy())
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
    if host_combine:
        assert np.array_equal(csr_k, csr)             # bit for bit: same order of additions
    return csr


@pytest.mark.parametrize("T,morph,tiling", [(128, False, "morton"), (256, True, "morton"),
                                            (512, True, "morton"), (128, False, "kd"),
                                            (256, True, "kd"), (512, True, "kd")])
def test_fused_plan_records_reproduce_the_csr(T, morph, tiling):
    import os
    from oracle import skfem_oracle as O
    from skfem_b200 import _lib, fused
    if not os.path.exists(_lib.LIB_PATH):     # the plan builder asks the library for its
        import __graft_entry__ as g           # shared-memory footprint (host-only call)
        g.build()
    m = O.mesh_tet_tensor(np.linspace(0, 1, 7), np.linspace(0, 1, 6), np.linspace(0, 1, 5))
    if morph:                # unstructured geometry: no exact zeros, full 15-point pattern
        q = m.p.copy()
        q[0] = m.p[0] + 0.03 * np.sin(7 * m.p[1])
        q[1] = m.p[1] + 0.02 * m.p[2] ** 2
        m = mesh_of(dict(p=q, t=m.t), "tet")
    basis, plan, A = _plan_on_cpu(m)
    fp = fused.build(basis, plan, T=T, tiling=tiling)
    assert fp.ntiles == -(-m.t.shape[1] // T) and fp.rec_cap % 16 == 0
    assert (fp.rec_start.numpy() % 16 == 0).all()
    csr = _emulate(fp, m.p, T)
    assert not np.isnan(csr).any()                    # every CSR slot written exactly once
    np.testing.assert_allclose(csr, A.data, rtol=1e-11, atol=1e-12 * np.abs(A.data).max())
    st = fused.stats(fp)
    assert st["sell_padding"] >= 1.0 and st["tile_slots_per_csr_slot"] >= 0.5


@pytest.mark.parametrize("tiling", ["morton", "kd"])
def test_fused_plan_on_an_unstructured_ball(tiling):
    """Same emulation on MeshTet.init_ball(2): curved, unstructured, vertex valences up to
    ~30 tets, so the F = 2 / 4 split lists and their shuffle-tree combination are exercised."""
    import os
    import skfem_b200 as fem
    from skfem_b200 import _lib, fused
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    ball = fem.MeshTet.init_ball(2)
    m = mesh_of(dict(p=ball.p, t=ball.t), "tet")
    basis, plan, A = _plan_on_cpu(m)
    fp = fused.build(basis, plan, T=128, tiling=tiling)
    csr = _emulate(fp, m.p, 128)
    assert not np.isnan(csr).any()
    np.testing.assert_allclose(csr, A.data, rtol=1e-11, atol=1e-12 * np.abs(A.data).max())


def test_kd_tiling_is_a_permutation_and_shares_fewer_slots():
    """fused._kd_order: every element exactly once, tiles = consecutive chunks of T, and on a
    Kuhn grid fewer canonical CSR slots are shared between tiles than with the Morton cut."""
    import os
    from oracle import skfem_oracle as O
    from skfem_b200 import _lib, fused
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    g1 = np.linspace(0, 1, 13)
    m = O.mesh_tet_tensor(g1, g1, g1)
    corner = torch.from_numpy(m.p[:, m.t].min(axis=1))
    for T in (128, 512, 500):
        order = fused._kd_order(corner, T).numpy()
        assert np.array_equal(np.sort(order), np.arange(m.t.shape[1]))
    basis, plan, _ = _plan_on_cpu(m)
    shared = {tl: fused.build(basis, plan, T=512, tiling=tl).nshared for tl in ("morton", "kd")}
    assert shared["kd"] < shared["morton"]
    with pytest.raises(ValueError):
        fused.build(basis, plan, T=512, tiling="hilbert")


def _smem_passes(fp, T):
    """Shared-memory passes (wavefronts) of P2's value gathers and of P1's coordinate
    gathers under the calibrated model of tools/sim_smem_conflicts.py: an 8-byte access of
    16 lanes costs max over the 16 bank pairs of the number of distinct words in the pair."""
    rec = fp.rec.numpy()
    rec16 = rec.view(np.uint16)
    rs = fp.rec_start.numpy()

    def cost(words):                       # words: (..., 16) -> passes per half-warp access
        words = words.reshape(-1, 16).astype(np.int64)
        out = 0
        for row in words:
            pairs = {}
            for w in set(row.tolist()):
                pairs[w & 15] = pairs.get(w & 15, 0) + 1
            out += max(pairs.values())
        return out
    p1 = p2 = 0
    for tile in range(fp.ntiles):
        base = int(rs[tile])
        hdr = [int(v) & 0xFFFFFFFF for v in rec[base // 4: base // 4 + 8]]
        ngroups, off_grp, off_ids = hdr[1], hdr[3], hdr[5]
        tl = rec16[(base + 32) // 2: (base + 32) // 2 + 4 * T].reshape(T, 4)
        live = tl[:, 0] != 0xFFFF
        for slot in range(4):
            col = np.where(live, tl[:, slot], 0)
            p1 += cost(col.reshape(-1, 16))
        grp = rec[(base + off_grp) // 4: (base + off_grp) // 4 + ngroups].view(np.uint32)
        for g in grp:
            ln, off = int(g >> 16), int(g & 0xFFFF)
            ids = rec16[(base + off_ids) // 2 + off * 32: (base + off_ids) // 2 + (off + ln) * 32]
            p2 += cost(ids.reshape(ln, 2, 16))
    return p1, p2


@pytest.mark.parametrize("tiling", ["morton", "kd"])
@pytest.mark.parametrize("shape", ["kuhn", "ball"])
def test_plan_passes_keep_the_plan_correct_and_cut_bank_conflicts(shape, tiling):
    import os
    import skfem_b200 as fem
    import host_plan_passes
    from oracle import skfem_oracle as O
    from skfem_b200 import _lib, fused
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    if shape == "kuhn":
        g1 = np.linspace(0, 1, 8)
        m = O.mesh_tet_tensor(g1, g1, g1)
    else:
        ball = fem.MeshTet.init_ball(2)
        m = mesh_of(dict(p=ball.p, t=ball.t), "tet")
    T = 256
    basis, plan, A = _plan_on_cpu(m)
    fp = fused.build(basis, plan, T=T, tiling=tiling)
    p1_before, p2_before = _smem_passes(fp, T)
    before = fp.rec.numpy().copy()
    host_plan_passes.apply(fp, T)
    assert (before != fp.rec.numpy()).any()
    csr = _emulate(fp, m.p, T, host_combine=True)
    assert not np.isnan(csr).any()
    np.testing.assert_allclose(csr, A.data, rtol=1e-11, atol=1e-12 * np.abs(A.data).max())
    p1_after, p2_after = _smem_passes(fp, T)
    assert p1_after < p1_before and p2_after < p2_before, (p1_before, p1_after,
                                                            p2_before, p2_after)
    print(shape, tiling, "P1 gathers", p1_before, "->", p1_after,
          "P2 gathers", p2_before, "->", p2_after)

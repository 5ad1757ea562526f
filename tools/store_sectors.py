"""How scattered are P2's global stores?  (CPU-only plan analysis.)

Builds the fused-path plan on CPU tensors (the builder is torch-only), decodes meta / meta2 of
every 32-lane group and counts the distinct 32-byte sectors each of the kernel's three store
instructions (csr_data[m], scratch[m & ~bit31], csr_data[m2]) touches.  The lower bound is
ceil(active lanes / 4).  Usage: python tools/store_sectors.py [cells_per_side] [--morph] [--kd]
"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scikit-fem_b200"))
from oracle import skfem_oracle as O          # noqa: E402  (analysis tool, not product)
from skfem_b200 import fused                  # noqa: E402

NONE = 0xFFFFFFFF


def plan_on_cpu(m, T, **kw):
    b = O.cell_basis(m, O.element("tet_p1"))
    A = O.assemble_bilinear(O.laplace, b)
    A.sort_indices()
    plan = SimpleNamespace(indptr=torch.from_numpy(A.indptr.astype(np.int32)),
                           indices=torch.from_numpy(A.indices.astype(np.int32)),
                           nnz=int(A.nnz), shape=A.shape)
    dev = {"device": torch.device("cpu"), "t": torch.from_numpy(m.t.astype(np.int32)),
           "tind": None, "p": torch.from_numpy(np.ascontiguousarray(m.p))}
    basis = SimpleNamespace(_dev=lambda: dev, W=b.W, nqp=b.W.shape[0])
    return fused.build(basis, plan, T=T, **kw)


def sectors(fp):
    rec = fp.rec.numpy()
    rs = fp.rec_start.numpy()
    tot = {"csr": [0, 0, 0], "scratch": [0, 0, 0], "mirror": [0, 0, 0]}   # instr, sectors, bound
    for tile in range(fp.ntiles):
        base = int(rs[tile])
        hdr = [int(v) & 0xFFFFFFFF for v in rec[base // 4: base // 4 + 8]]
        ngroups, off_meta, off_meta2 = hdr[1], hdr[4], hdr[7]
        meta = rec[(base + off_meta) // 4: (base + off_meta) // 4 + 32 * ngroups].view(np.uint32)
        meta2 = rec[(base + off_meta2) // 4: (base + off_meta2) // 4 + 32 * ngroups].view(np.uint32)
        for g in range(ngroups):
            m = meta[g * 32:(g + 1) * 32].astype(np.int64)
            m2 = meta2[g * 32:(g + 1) * 32].astype(np.int64)
            live = m != NONE
            sc = live & ((m & 0x80000000) != 0)
            for name, idx in (("csr", m[live & ~sc]), ("scratch", m[sc] & 0x7FFFFFFF),
                              ("mirror", m2[m2 != NONE])):
                if idx.size:
                    t = tot[name]
                    t[0] += 1
                    t[1] += np.unique(idx // 4).size
                    t[2] += -(-idx.size // 4)
    return tot


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 20
    g = np.linspace(0, 1, n + 1)
    m = O.mesh_tet_tensor(g, g, g)
    if "--morph" in sys.argv:
        q = m.p.copy()
        q[0] = m.p[0] + 0.03 * np.sin(7 * m.p[1])
        q[1] = m.p[1] + 0.02 * m.p[2] ** 2
        m = SimpleNamespace(p=np.ascontiguousarray(q), t=np.ascontiguousarray(m.t), refdom="tet")
    fp = plan_on_cpu(m, 512, tiling="kd" if "--kd" in sys.argv else "morton")
    st = fused.stats(fp)
    tot = sectors(fp)
    print("tiling %s: tiles %d, ELL padding %.3f, record bytes per tile %.0f, vcap %d, smem %d, "
          "tile slots %d, bytes per element %.1f"
          % (st["tiling"], fp.ntiles, st["sell_padding"], st["records"] / fp.ntiles, st["vcap"],
             st["smem"], fp.nts, st["per_element"]))
    ncanon = (fp.nnz + int(m.p.shape[1])) // 2          # canonical slots: row <= col
    print("canonical CSR slots %d, shared between tiles %d (%.1f %%), partials %d"
          % (ncanon, fp.nshared, 100.0 * fp.nshared / ncanon, fp.nscratch))
    for name, (ni, ns, nb) in tot.items():
        print("%-8s %7d store instr, %6.2f sectors each (bound %5.2f), %8.1f sectors per tile"
              % (name, ni, ns / max(ni, 1), nb / max(ni, 1), ns / fp.ntiles))
    print("total sectors per tile: %.1f (bound %.1f)"
          % (sum(v[1] for v in tot.values()) / fp.ntiles,
             sum(v[2] for v in tot.values()) / fp.ntiles))


if __name__ == "__main__":
    main()

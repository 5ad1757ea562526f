"""Offline model of the fused kernel's shared-memory bank conflicts.

Builds the fused plan on the CPU (the plan builder is torch-only; the CSR
pattern comes from the oracle) for an init_tensor mesh, decodes the per-tile
records exactly as the kernel does and counts shared-memory wavefronts of

  * P2: the LDS.64 gathers ``in[ids[c*32 + lane]]`` of every group / column,
  * P1: the LDS.64 coordinate gathers ``sx[tl[e].a]`` of every warp of 32
        consecutive elements.

Model: an LDS.64 is served in two half-warp passes; a 64-bit access to word
index p occupies bank pair p mod 16; lanes of a half-warp reading the same word
are broadcast; wavefronts of a pass = max over bank pairs of the number of
distinct words.  Calibrated against ncu's "L1 Wavefronts Shared" per instruction
on the 100^3 mesh (round-1 layout): model 5.48 / 3.23 wavefronts per P2 / P1
gather, ncu 5.4 / 3.3.
Used to evaluate plan-time layout choices without GPU time.

    python tools/sim_smem_conflicts.py --cells 24 [--tile 512]
"""
import argparse
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scikit-fem_b200"))


def cpu_plan(cells):
    from oracle import skfem_oracle as O
    x = np.linspace(0, 1, cells + 1)
    m = O.mesh_tet_tensor(x, x, x)
    b = O.cell_basis(m, O.element("tet_p1"))
    A = O.assemble_bilinear(O.laplace, b)
    A.sort_indices()
    plan = SimpleNamespace(indptr=torch.from_numpy(A.indptr.astype(np.int32)),
                           indices=torch.from_numpy(A.indices.astype(np.int32)),
                           nnz=int(A.nnz), shape=A.shape)
    dev = {"device": torch.device("cpu"), "t": torch.from_numpy(m.t.astype(np.int32)),
           "tind": None, "p": torch.from_numpy(np.ascontiguousarray(m.p))}
    basis = SimpleNamespace(_dev=lambda: dev, W=b.W, nqp=b.W.shape[0])
    return basis, plan


def wavefronts(words):
    """words: (n, 32) int array of 8-byte word indices -> wavefronts per row."""
    total = np.zeros(words.shape[0], dtype=np.int64)
    for half in (words[:, :16], words[:, 16:]):      # LDS.64: two half-warp passes
        w = np.sort(half, axis=1)
        first = np.ones_like(w, dtype=bool)
        first[:, 1:] = w[:, 1:] != w[:, :-1]
        out = np.zeros(words.shape[0], dtype=np.int64)
        for b in range(16):
            out = np.maximum(out, ((w % 16 == b) & first).sum(axis=1))
        total += out
    return total


def analyse(fp, T, max_tiles=200):
    rec = fp.rec.numpy()
    rec16 = rec.view(np.uint16)
    rs = fp.rec_start.numpy()
    p2_w, p2_n, p1_w, p1_n = 0, 0, 0, 0
    for tile in range(min(fp.ntiles, max_tiles)):
        base = rs[tile]
        hdr = rec[base // 4: base // 4 + 8]
        nverts, ngroups, off_verts, off_grp, off_meta, off_ids = (int(v) for v in hdr[:6])
        tl = rec16[(base + 32) // 2: (base + 32) // 2 + 4 * T].reshape(T, 4).astype(np.int64)
        grp = rec[(base + off_grp) // 4: (base + off_grp) // 4 + ngroups].astype(np.int64)
        ids0 = (base + off_ids) // 2
        for g in range(ngroups):
            ln, off = int(grp[g] >> 16) & 0xffff, int(grp[g] & 0xffff)
            cols = rec16[ids0 + off * 32: ids0 + (off + ln) * 32].reshape(ln, 32).astype(np.int64)
            p2_w += int(wavefronts(cols).sum())
            p2_n += ln
        valid = tl[:, 0] != 0xFFFF
        for w0 in range(0, T, 32):
            if not valid[w0:w0 + 32].all():
                continue
            for a in range(4):
                p1_w += 3 * int(wavefronts(tl[w0:w0 + 32, a][None, :])[0])
                p1_n += 3
    return p2_w / max(p2_n, 1), p1_w / max(p1_n, 1), p2_n, p1_n


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=24)
    ap.add_argument("--tile", type=int, default=512)
    ap.add_argument("--tiles", type=int, default=120)
    ap.add_argument("--gpu", action="store_true",
                    help="take the plan the product builds on the GPU instead of the CPU build")
    args = ap.parse_args()
    from skfem_b200 import fused
    if args.gpu:
        import skfem_b200 as fem
        from skfem_b200.form import set_options
        from skfem_b200.models.poisson import laplace
        set_options(fused_tile=args.tile)
        x = np.linspace(0, 1, args.cells + 1)
        gb = fem.Basis(fem.MeshTet.init_tensor(x, x, x), fem.ElementTetP1())
        laplace.assemble_device(gb)
        laplace.assemble_device(gb)
        fp = [v for v in gb._plans.values() if isinstance(v, fused.P1FusedPlan)][0]
        fp.rec, fp.rec_start = fp.rec.cpu(), fp.rec_start.cpu()
    else:
        basis, plan = cpu_plan(args.cells)
        fp = fused.build(basis, plan, T=args.tile)
    p2, p1, n2, n1 = analyse(fp, args.tile, args.tiles)
    print("tiles {}  P2 gathers: {:.2f} wavefronts/LDS.64 over {} instr (ideal 2)   "
          "P1 coordinate gathers: {:.2f} over {} instr (ideal <= 2)".format(
              fp.ntiles, p2, n2, p1, n1))

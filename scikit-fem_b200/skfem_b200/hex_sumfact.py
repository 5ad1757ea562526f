"""Host tables for the sum-factorised ElementHex2 kernel (csrc/skb_hex_sf.cu).

The 27 basis functions of ElementHex2 (skfem/element/element_hex/element_hex2.py:1213-1260)
are products l_a(x) l_b(y) l_c(z) of the 1-D quadratic Lagrange functions on the nodes
{0, 1/2, 1}, the 8 functions of the geometry element ElementHex1 (element_hex1.py:23-69) are
products of {1 - x, x}, and the default rule of a hexahedron (quadrature.py:63-77) is the
tensor product of one Gauss-Legendre rule.  The kernel therefore contracts one axis at a time;
what it needs from the host are the 1-D tables below - everything is derived from the element's
``doflocs`` and the basis' own quadrature points, nothing is assumed about their ordering.

    pp[type][3 i + j][q]   products of the 1-D functions at the 1-D points,
                           type bit 0: derivative on the i side, bit 1: on the j side
    g[s][q]                1 - x_q (s = 0), x_q (s = 1)
    bnode[k]               a + 3 b + 9 c of basis function k  (node indices 0, 1, 2 = 0, 1/2, 1)
    vtx[4 a + 2 b + c]     local vertex of the geometry element at the corner (a, b, c)
    qstride[k]             stride of axis k's 1-D index in the basis' quadrature point index
"""
import ctypes as C

import numpy as np

from .element import ElementHex1, ElementHex2, _lagrange2

NQ = 7            # the kernel is instantiated for the default rule: 7 Gauss points per axis


def _tensor_grid(X):
    """(x1d, strides) when the points X (3, nq^3) are a full tensor grid of one 1-D point set
    indexed as q = sum_k i_k * stride_k, else None."""
    if X.ndim != 2 or X.shape[0] != 3:
        return None
    x1d = np.unique(X[0])
    n = x1d.shape[0]
    if n ** 3 != X.shape[1]:
        return None
    strides = []
    q = np.arange(X.shape[1])
    for k in range(3):
        diff = np.nonzero(X[k] != X[k][0])[0]
        if diff.size == 0:
            return None
        s = int(diff[0])
        if not np.array_equal(X[k], x1d[(q // s) % n]):
            return None
        strides.append(s)
    if sorted(strides) != [1, n, n * n]:
        return None
    return x1d, strides


def tables(basis):
    """dict of contiguous numpy arrays for ``skb_local_hex_sumfact`` or None when the basis is
    not (ElementHex2 on a trilinear hexahedral mesh at a 7^3 tensor rule)."""
    cached = getattr(basis, "_hex_sf", False)
    if cached is not False:
        return cached
    out = None
    try:
        ok = (type(basis.elem) is ElementHex2 and basis.ncomp == 1 and not basis._affine
              and type(basis.mesh.elem()) is ElementHex1)
    except AttributeError:
        ok = False
    grid = _tensor_grid(basis.X) if ok else None
    if grid is not None and grid[0].shape[0] == NQ:
        x1d, strides = grid
        nodes = (0., .5, 1.)
        l = np.array([_lagrange2(n, x1d)[0] for n in nodes])       # (3, nq)
        dl = np.array([_lagrange2(n, x1d)[1] for n in nodes])
        pp = np.empty((4, 9, NQ))
        for ty in range(4):
            u = dl if ty & 1 else l
            v = dl if ty & 2 else l
            for i in range(3):
                for j in range(3):
                    pp[ty, 3 * i + j] = u[i] * v[j]
        bnode = np.zeros(32, dtype=np.uint8)
        locs = np.asarray(basis.elem.doflocs)
        code = np.rint(2 * locs).astype(np.int64)                  # 0, 1, 2
        corners = np.rint(np.asarray(basis.mesh.elem().doflocs)).astype(np.int64)
        vtx = np.zeros(8, dtype=np.uint8)
        vtx[4 * corners[:, 0] + 2 * corners[:, 1] + corners[:, 2]] = np.arange(8)
        if (np.array_equal(code / 2., locs) and len(set(map(tuple, code))) == 27
                and len(set(map(tuple, corners))) == 8):
            bnode[:27] = code[:, 0] + 3 * code[:, 1] + 9 * code[:, 2]
            out = {"nq": NQ, "pp": np.ascontiguousarray(pp),
                   "g": np.ascontiguousarray(np.array([1. - x1d, x1d])),
                   "bnode": bnode, "vtx": vtx,
                   "qstride": np.asarray(strides, dtype=np.int32), "x1d": x1d}
    basis._hex_sf = out
    return out


def launch(lib, space, form_id, tab, out_ptr, stream, element_major=False):
    """Call the C entry; returns its status code."""
    as_p = lambda a, ty: a.ctypes.data_as(C.POINTER(ty))
    return lib.skb_local_hex_sumfact(
        C.byref(space), int(form_id), int(tab["nq"]), as_p(tab["qstride"], C.c_int32),
        as_p(tab["pp"], C.c_double), as_p(tab["g"], C.c_double),
        as_p(tab["bnode"], C.c_uint8), as_p(tab["vtx"], C.c_uint8), int(bool(element_major)),
        out_ptr, stream)

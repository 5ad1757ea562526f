"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by the product).

A flat numpy/scipy restatement of the scikit-fem (v12.0.1) assembly hot path

    CellBasis(mesh, elem) -> BilinearForm/LinearForm._assemble -> COO -> CSR

It executes the *same numpy/scipy primitives in the same order* as the
reference (einsum contractions, ufunc chains, ``np.sum(axis=1)``,
``coo_matrix.eliminate_zeros().tocsr()``), so the results are bit-identical
to the reference on the same machine.  Parity status: PINNED -- the functions
below are checked bitwise against outputs of the real reference imported from
/root/reference (tools/gen_golden.py -> tests/golden/*.npz, tests/test_oracle_*).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.

Every function cites the reference file:line it follows (paths relative to
/root/reference/skfem/).
"""
from __future__ import annotations

import os
from types import SimpleNamespace

import numpy as np
from numpy.polynomial.legendre import leggauss
from scipy.sparse import coo_matrix

_HERE = os.path.dirname(os.path.abspath(__file__))
_QTAB = None


# --------------------------------------------------------------------------
# reference-domain tables (refdom.py:55-209)
# --------------------------------------------------------------------------
REFDOM = {
    "tri": dict(dim=2, nnodes=3, nedges=0,
                facets=[[0, 1], [1, 2], [0, 2]], edges=None),
    "tet": dict(dim=3, nnodes=4, nedges=6,
                facets=[[0, 1, 2], [0, 1, 3], [0, 2, 3], [1, 2, 3]],
                edges=[[0, 1], [1, 2], [0, 2], [0, 3], [1, 3], [2, 3]]),
    "hex": dict(dim=3, nnodes=8, nedges=12,
                facets=[[0, 1, 4, 2], [0, 2, 6, 3], [0, 3, 5, 1],
                        [2, 4, 7, 6], [1, 5, 7, 4], [3, 6, 7, 5]],
                edges=[[0, 1], [0, 2], [0, 3], [1, 4], [1, 5], [2, 4],
                       [2, 6], [3, 5], [3, 6], [4, 7], [5, 7], [6, 7]]),
}


# --------------------------------------------------------------------------
# meshes (mesh/mesh_tet_1.py:326-393, mesh_hex_1.py:97-155,
#         mesh_tri_1.py:14-28,209-253, mesh.py:544-607,1065-1082)
# --------------------------------------------------------------------------
def _finish(p, t, refdom, sort_t=False):
    if sort_t:  # mesh.py:555-556 (MeshTri1.sort_t, mesh_tri_1.py:28)
        t = np.sort(t, axis=0)
    return SimpleNamespace(
        p=np.ascontiguousarray(p, dtype=np.float64),
        t=np.ascontiguousarray(t, dtype=np.int32),
        refdom=refdom,
    )


def _tensor_cells(x, y, z):
    """Vertex grid + the 8 corner index rows of every cell, in the
    reference's Fortran-order numbering (mesh_tet_1.py:343-384)."""
    npx, npy, npz = len(x), len(y), len(z)
    X, Y, Z = np.meshgrid(np.sort(x), np.sort(y), np.sort(z))
    p = np.vstack((X.flatten('F'), Y.flatten('F'), Z.flatten('F')))
    ix = np.arange(npx * npy * npz).reshape(npy, npx, npz, order='F')
    ne = (npx - 1) * (npy - 1) * (npz - 1)
    lo = (slice(0, npy - 1), slice(0, npx - 1), slice(0, npz - 1))
    hi = (slice(1, npy), slice(1, npx), slice(1, npz))
    corners = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1),
               (1, 1, 0), (1, 0, 1), (0, 1, 1), (1, 1, 1)]
    t = np.zeros((8, ne), dtype=np.int64)
    for k, (a, b, c) in enumerate(corners):
        sl = (hi[0] if a else lo[0], hi[1] if b else lo[1],
              hi[2] if c else lo[2])
        t[k] = ix[sl].reshape(ne, order='F')
    return p, t, ne


def mesh_tet_tensor(x, y, z):
    """MeshTet1.init_tensor (mesh_tet_1.py:326-393): 6 Kuhn tets per cell."""
    p, t, ne = _tensor_cells(x, y, z)
    T = np.hstack([t[rows] for rows in ([0, 1, 5, 7], [0, 1, 4, 7],
                                        [0, 2, 4, 7], [0, 3, 5, 7],
                                        [0, 2, 6, 7], [0, 3, 6, 7])])
    return _finish(p, T, "tet")


def mesh_hex_tensor(x, y, z):
    """MeshHex1.init_tensor (mesh_hex_1.py:97-155)."""
    p, t, _ = _tensor_cells(x, y, z)
    return _finish(p, t, "hex")


def mesh_tri_default():
    """MeshTri1 defaults (mesh_tri_1.py:16-28)."""
    p = np.array([[0., 0.], [1., 0.], [0., 1.], [1., 1.]]).T
    t = np.array([[0, 1, 2], [1, 3, 2]]).T
    return _finish(p, t, "tri", sort_t=True)


def mesh_tet_default():
    """MeshTet1 defaults (mesh_tet_1.py:16-42)."""
    p = np.array([[0., 0., 0.], [0., 0., 1.], [0., 1., 0.], [1., 0., 0.],
                  [0., 1., 1.], [1., 0., 1.], [1., 1., 0.], [1., 1., 1.]]).T
    t = np.array([[0, 1, 2, 3], [3, 5, 1, 7], [2, 3, 6, 7], [2, 3, 1, 7],
                  [1, 2, 4, 7]]).T
    return _finish(p, t, "tet")


def build_entities(t, indices, sort=True):
    """Mesh.build_entities (mesh/mesh.py:1065-1082)."""
    indexing = np.hstack(tuple([t[ix] for ix in indices]))
    sorted_indexing = np.sort(indexing, axis=0)
    sorted_indexing, ixa, ixb = np.unique(sorted_indexing, axis=1,
                                          return_index=True,
                                          return_inverse=True)
    mapping = ixb.reshape((len(indices), t.shape[1]))
    if sort:
        return np.ascontiguousarray(sorted_indexing), mapping
    return np.ascontiguousarray(indexing[:, ixa]), mapping


def facets_of(m):
    """Mesh._init_facets (mesh.py:530-535); hexes unsorted
    (mesh_hex_1.py:49-55)."""
    if not hasattr(m, "_facets"):
        m._facets, m._t2f = build_entities(
            m.t, REFDOM[m.refdom]["facets"], sort=(m.refdom != "hex"))
    return m._facets, m._t2f


def edges_of(m):
    """Mesh._init_edges (mesh.py:537-542)."""
    if not hasattr(m, "_edges"):
        m._edges, m._t2e = build_entities(m.t, REFDOM[m.refdom]["edges"])
    return m._edges, m._t2e


def refine_tri(m, n=1):
    """MeshTri1._uniform (mesh_tri_1.py:209-227), n times."""
    for _ in range(n):
        p, t = m.p, m.t
        sz = p.shape[1]
        facets, t2f = facets_of(m)
        newp = np.hstack((p, p[:, facets].mean(axis=1)))
        newt = np.hstack((
            np.vstack((t[0], t2f[0] + sz, t2f[2] + sz)),
            np.vstack((t[1], t2f[0] + sz, t2f[1] + sz)),
            np.vstack((t[2], t2f[2] + sz, t2f[1] + sz)),
            np.vstack((t2f[0] + sz, t2f[1] + sz, t2f[2] + sz)),
        ))
        m = _finish(newp, newt, "tri", sort_t=True)
    return m


# --------------------------------------------------------------------------
# quadrature (quadrature.py:12-77, 2839-2844)
# --------------------------------------------------------------------------
def quadrature(refdom, order):
    global _QTAB
    if refdom in ("tri", "tet"):
        if _QTAB is None:
            _QTAB = np.load(os.path.join(_HERE, "quadrature_tables.npz"))
        order = max(order, 1)
        key = f"{refdom}_{order}_X"
        if key not in _QTAB:
            raise NotImplementedError("quadrature order not tabulated")
        return _QTAB[key], _QTAB[f"{refdom}_{order}_W"]
    if order <= 1:
        order = 2
    X, W = leggauss(int(np.ceil((order + 1.0) / 2.0)))
    X, W = np.array([0.5 * X + 0.5]), W / 2.0
    if refdom == "line":
        return X, W
    if refdom == "hex":
        A, B, C = np.meshgrid(X, X, X)
        Y = np.vstack((A.flatten(order="F"), B.flatten(order="F"),
                       C.flatten(order="F")))
        A, B, C = np.meshgrid(W, W, W)
        return Y, (A * B * C).flatten(order="F")
    raise NotImplementedError(refdom)


# --------------------------------------------------------------------------
# elements: lbasis on the reference domain
# --------------------------------------------------------------------------
def _lb_tri_p1(X, i):  # element_tri/element_tri_p1.py:18-33
    x, y = X
    if i == 0:
        return 1. - x - y, np.array([-1. + 0. * x, -1. + 0. * x])
    if i == 1:
        return x, np.array([1. + 0. * x, 0. * x])
    return y, np.array([0. * x, 1. + 0. * x])


def _lb_tri_p2(X, i):  # element_tri/element_tri_p2.py:22-50
    x, y = X
    if i == 0:
        return (1. - 3. * x - 3. * y + 2. * x ** 2 + 4. * x * y + 2. * y ** 2,
                np.array([-3 + 4. * x + 4. * y, -3 + 4. * x + 4. * y]))
    if i == 1:
        return 2. * x ** 2 - x, np.array([4. * x - 1, 0. * x])
    if i == 2:
        return 2. * y ** 2 - y, np.array([0. * x, 4. * y - 1])
    if i == 3:
        return 4. * x - 4. * x ** 2 - 4. * x * y, np.array([4 - 8. * x - 4. * y, -4. * x])
    if i == 4:
        return 4. * x * y, np.array([4. * y, 4. * x])
    return 4. * y - 4. * x * y - 4. * y ** 2, np.array([-4. * y, 4 - 4. * x - 8. * y])


def _lb_tet_p1(X, i):  # element_tet/element_tet_p1.py:19-45
    x, y, z = X
    one, zero = 1 + 0 * x, 0 * x
    if i == 0:
        return 1 - x - y - z, np.array([-1 + 0 * x, -1 + 0 * x, -1 + 0 * x])
    d = [zero, zero, zero]
    d[i - 1] = one
    return X[i - 1], np.array(d)


def _lb_tet_p2(X, i):  # element_tet/element_tet_p2.py:26-103
    x, y, z = X
    o = 0 * x
    if i == 0:
        phi = (1. - 3. * x + 2. * x ** 2 - 3. * y + 4. * x * y + 2. * y ** 2
               - 3. * z + 4. * x * z + 4. * y * z + 2. * z ** 2)
        g = -3. + 4. * x + 4. * y + 4. * z
        return phi, np.array([g, g, g])
    if i == 1:
        return -1. * x + 2. * x ** 2, np.array([-1 + 4 * x, o, o])
    if i == 2:
        return -1. * y + 2. * y ** 2, np.array([o, -1. + 4. * y, o])
    if i == 3:
        return -1. * z + 2. * z ** 2, np.array([o, o, -1. + 4. * z])
    if i == 4:
        return (4. * x - 4. * x ** 2 - 4. * x * y - 4 * x * z,
                np.array([4. - 8. * x - 4. * y - 4. * z, -4. * x, -4. * x]))
    if i == 5:
        return 4. * x * y, np.array([4. * y, 4. * x, o])
    if i == 6:
        return (0. + 4. * y - 4. * x * y - 4. * y ** 2 - 4. * y * z,
                np.array([-4. * y, 4. - 4. * x - 8. * y - 4. * z, -4. * y]))
    if i == 7:
        return (0. + 4. * z - 4. * x * z - 4. * y * z - 4. * z ** 2,
                np.array([-4. * z, -4. * z, 4. - 4. * x - 4. * y - 8. * z]))
    if i == 8:
        return 0. + 4. * x * z, np.array([4. * z, o, 4 * x])
    return 0. + 4. * y * z, np.array([o, 4 * z, 4 * y])


def _lb_hex1(X, i):  # element_hex/element_hex1.py:23-69
    x, y, z = X
    # vertex i sits at RefHex.p[:, i] (refdom.py:173-180); factor is the
    # coordinate itself (corner value 1) or its complement (corner value 0)
    corner = [(1, 1, 1), (1, 1, 0), (1, 0, 1), (0, 1, 1),
              (1, 0, 0), (0, 1, 0), (0, 0, 1), (0, 0, 0)][i]
    f = [c if k else (1 - c) for c, k in zip((x, y, z), corner)]
    s = [1 if k else -1 for k in corner]
    phi = f[0] * f[1] * f[2]

    def signed(sign, a, b):
        return a * b if sign > 0 else -a * b
    return phi, np.array([signed(s[0], f[1], f[2]),
                          signed(s[1], f[0], f[2]),
                          signed(s[2], f[0], f[1])])


ELEMENTS = {
    # name: refdom, nodal, edge, facet, interior dofs, maxdeg, nbf, lbasis
    "tri_p1": dict(refdom="tri", nodal=1, edge=0, facet=0, interior=0,
                   maxdeg=1, nbf=3, lbasis=_lb_tri_p1),
    "tri_p2": dict(refdom="tri", nodal=1, edge=0, facet=1, interior=0,
                   maxdeg=2, nbf=6, lbasis=_lb_tri_p2),
    "tet_p1": dict(refdom="tet", nodal=1, edge=0, facet=0, interior=0,
                   maxdeg=1, nbf=4, lbasis=_lb_tet_p1),
    "tet_p2": dict(refdom="tet", nodal=1, edge=1, facet=0, interior=0,
                   maxdeg=2, nbf=10, lbasis=_lb_tet_p2),
    "hex1": dict(refdom="hex", nodal=1, edge=0, facet=0, interior=0,
                 maxdeg=3, nbf=8, lbasis=_lb_hex1),
}


def element(name, vector=False):
    e = dict(ELEMENTS[name])
    e["name"] = name
    e["vector"] = bool(vector)
    e["dim"] = REFDOM[e["refdom"]]["dim"]
    return SimpleNamespace(**e)


# --------------------------------------------------------------------------
# global DOF numbering (assembly/dofs.py:264-334)
# --------------------------------------------------------------------------
def dofs(m, elem):
    rd = REFDOM[m.refdom]
    mult = elem.dim if elem.vector else 1  # element_vector.py:13-17
    nverts = int(np.max(m.t)) + 1          # mesh.py:71-73
    nel = m.t.shape[1]
    offset = 0
    nd = elem.nodal * mult
    nodal = np.reshape(np.arange(nd * nverts, dtype=np.int32),
                       (nd, nverts), order='F') + offset
    offset += nd * nverts
    ed = elem.edge * mult
    if rd["dim"] == 3 and ed > 0:
        edges, t2e = edges_of(m)
        edge = np.reshape(np.arange(ed * edges.shape[1], dtype=np.int32),
                          (ed, edges.shape[1]), order='F') + offset
        offset += ed * edges.shape[1]
    fd = elem.facet * mult
    if fd > 0:
        facets, t2f = facets_of(m)
        facet = np.reshape(np.arange(fd * facets.shape[1], dtype=np.int32),
                           (fd, facets.shape[1]), order='F') + offset
        offset += fd * facets.shape[1]
    idf = elem.interior * mult
    interior = np.reshape(np.arange(idf * nel, dtype=np.int32),
                          (idf, nel), order='F') + offset
    rows = [np.zeros((0, nel), dtype=np.int32)]
    for itr in range(m.t.shape[0]):
        rows.append(nodal[:, m.t[itr]])
    if rd["dim"] == 3 and ed > 0:
        for itr in range(t2e.shape[0]):
            rows.append(edge[:, t2e[itr]])
    if rd["dim"] >= 2 and fd > 0:
        for itr in range(t2f.shape[0]):
            rows.append(facet[:, t2f[itr]])
    rows.append(interior)
    element_dofs = np.vstack(rows)
    return element_dofs, int(np.max(element_dofs)) + 1


# --------------------------------------------------------------------------
# mappings
# --------------------------------------------------------------------------
def affine_geometry(m, tind=None):
    """MappingAffine._init_Ab/_init_invA (mapping/mapping_affine.py:55-131)."""
    p, t = m.p, (m.t if tind is None else m.t[:, tind])
    dim = p.shape[0]
    nt = t.shape[1]
    A = np.empty((dim, dim, nt))
    b = np.empty((dim, nt))
    for i in range(dim):
        b[i] = p[i, t[0]]
        for j in range(dim):
            A[i, j] = p[i, t[j + 1]] - p[i, t[0]]
    invA = np.empty((dim, dim, nt))
    if dim == 2:
        detA = A[0, 0] * A[1, 1] - A[0, 1] * A[1, 0]
        invA[0, 0] = A[1, 1] / detA
        invA[0, 1] = -A[0, 1] / detA
        invA[1, 0] = -A[1, 0] / detA
        invA[1, 1] = A[0, 0] / detA
    else:
        detA = (A[0, 0] * (A[1, 1] * A[2, 2] - A[1, 2] * A[2, 1]) -
                A[0, 1] * (A[1, 0] * A[2, 2] - A[1, 2] * A[2, 0]) +
                A[0, 2] * (A[1, 0] * A[2, 1] - A[1, 1] * A[2, 0]))
        invA[0, 0] = (-A[1, 2] * A[2, 1] + A[1, 1] * A[2, 2]) / detA
        invA[1, 0] = (A[1, 2] * A[2, 0] - A[1, 0] * A[2, 2]) / detA
        invA[2, 0] = (-A[1, 1] * A[2, 0] + A[1, 0] * A[2, 1]) / detA
        invA[0, 1] = (A[0, 2] * A[2, 1] - A[0, 1] * A[2, 2]) / detA
        invA[1, 1] = (-A[0, 2] * A[2, 0] + A[0, 0] * A[2, 2]) / detA
        invA[2, 1] = (A[0, 1] * A[2, 0] - A[0, 0] * A[2, 1]) / detA
        invA[0, 2] = (-A[0, 2] * A[1, 1] + A[0, 1] * A[1, 2]) / detA
        invA[1, 2] = (A[0, 2] * A[1, 0] - A[0, 0] * A[1, 2]) / detA
        invA[2, 2] = (-A[0, 1] * A[1, 0] + A[0, 0] * A[1, 1]) / detA
    return SimpleNamespace(kind="affine", A=A, b=b, invA=invA, detA=detA,
                           dim=dim, nel=nt)


def iso_geometry(m, X, tind=None):
    """MappingIsoparametric J/detDF/invDF/F for a Hex1 mesh
    (mapping/mapping_isoparametric.py:112-126,170-226)."""
    t = m.t if tind is None else m.t[:, tind]
    nel, nqp = t.shape[1], X.shape[1]
    dim = m.p.shape[0]
    tab = [_lb_hex1(X, n) for n in range(t.shape[0])]
    J = [[None] * dim for _ in range(dim)]
    F = []
    for i in range(dim):
        out = np.zeros((nel, nqp))
        for n in range(t.shape[0]):           # :106-110 (F)
            out += np.outer(m.p[i, t[n]], tab[n][0])
        F.append(out)
        for j in range(dim):
            acc = np.zeros((nel, nqp))
            for n in range(t.shape[0]):       # :121-125 (J)
                acc += np.outer(m.p[i, t[n]], tab[n][1][j])
            J[i][j] = acc
    det = (J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) -
           J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
           J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]))
    if np.sum(det == 0) > 0:
        raise Exception("Zero Jacobian determinant")
    inv = np.empty((3, 3, nel, nqp))
    inv[0, 0] = -J[1][2] * J[2][1] + J[1][1] * J[2][2]
    inv[1, 0] = J[1][2] * J[2][0] - J[1][0] * J[2][2]
    inv[2, 0] = -J[1][1] * J[2][0] + J[1][0] * J[2][1]
    inv[0, 1] = J[0][2] * J[2][1] - J[0][1] * J[2][2]
    inv[1, 1] = -J[0][2] * J[2][0] + J[0][0] * J[2][2]
    inv[2, 1] = J[0][1] * J[2][0] - J[0][0] * J[2][1]
    inv[0, 2] = -J[0][2] * J[1][1] + J[0][1] * J[1][2]
    inv[1, 2] = J[0][2] * J[1][0] - J[0][0] * J[1][2]
    inv[2, 2] = -J[0][1] * J[1][0] + J[0][0] * J[1][1]
    inv = inv / det
    return SimpleNamespace(kind="iso", invDF=inv, detDF=det, F=np.array(F),
                           dim=dim, nel=nel)


# --------------------------------------------------------------------------
# DiscreteField stand-in (element/discrete_field.py:7-63): an ndarray that
# carries .grad and degrades to a plain array under ufuncs / indexing.
# --------------------------------------------------------------------------
class Field(np.ndarray):
    def __new__(cls, value, grad=None):
        obj = np.asarray(value).view(cls)
        obj.grad = grad
        return obj

    def __array_finalize__(self, obj):
        if obj is not None:
            self.grad = getattr(obj, 'grad', None)

    def __getitem__(self, key):
        return np.array(self)[key]

    def __array_wrap__(self, out_arr, context=None, return_scalar=False):
        return np.array(out_arr)


# --------------------------------------------------------------------------
# CellBasis (assembly/basis/cell_basis.py:94-141, abstract_basis.py:45-88)
# --------------------------------------------------------------------------
def cell_basis(m, elem, intorder=None, elements=None, quadrature_rule=None):
    if REFDOM[m.refdom]["dim"] != elem.dim or m.refdom != elem.refdom:
        raise ValueError("Incompatible Mesh and Element.")
    edofs, N = dofs(m, elem)
    if quadrature_rule is not None:
        X, W = quadrature_rule
    else:
        X, W = quadrature(m.refdom,
                          intorder if intorder is not None
                          else 2 * elem.maxdeg)
    tind = None if elements is None else np.asarray(elements)
    nqp = W.shape[-1]
    if m.refdom == "hex":
        geo = iso_geometry(m, X, tind)
        invDF, detDF = geo.invDF, geo.detDF
        x = geo.F
    else:
        geo = affine_geometry(m, tind)
        detDF = np.tile(geo.detA, (nqp, 1)).T        # mapping_affine.py:205-211
        x = (np.einsum('ijk,jl', geo.A, X).T + geo.b.T).T   # :183-193
        invDF = None
    nel = geo.nel
    nbf_scalar = elem.nbf
    scalar = []
    for j in range(nbf_scalar):
        phi, dphi = elem.lbasis(X, j)
        if invDF is None:
            # ElementH1.gbasis re-materialises invDF per call
            # (element_h1.py:10-18, mapping_affine.py:225-232)
            inv = np.einsum('ijk,l->ijkl', geo.invA, np.ones(nqp))
        else:
            inv = invDF
        scalar.append(Field(np.broadcast_to(phi, (nel, nqp)),
                            np.einsum('ijkl,il->jkl', inv, dphi)))
    if elem.vector:                                   # element_vector.py:36-48
        basis = []
        for i in range(nbf_scalar * elem.dim):
            f = scalar[i // elem.dim]
            n = i % elem.dim
            val = np.zeros((elem.dim,) + f.shape)
            val[n] = np.array(f)
            grd = np.zeros((elem.dim,) + f.grad.shape)
            grd[n] = f.grad
            basis.append(Field(val, grd))
    else:
        basis = scalar
    dx = np.abs(detDF) * np.broadcast_to(W, (nel, nqp))   # cell_basis.py:104
    edofs_l = edofs if tind is None else edofs[:, tind]
    h = np.abs(detDF) ** (1. / elem.dim)               # cell_basis.py:136-141
    return SimpleNamespace(mesh=m, elem=elem, X=X, W=W, basis=basis, dx=dx,
                           element_dofs=edofs_l, N=N, Nbfun=len(basis),
                           nelems=nel, x=Field(x), h=Field(h), geo=geo,
                           dofs_full=edofs)


# --------------------------------------------------------------------------
# FacetBasis (assembly/basis/facet_basis.py:76-140) on affine meshes
# --------------------------------------------------------------------------
def f2t_of(m):
    """Mesh.f2t = build_inverse(t, t2f) (mesh/mesh.py:1085-1100): first and
    last element touching each facet, -1 when they coincide."""
    if not hasattr(m, "_f2t"):
        _, t2f = facets_of(m)
        flat = t2f.flatten(order='C')
        owner = np.tile(np.arange(m.t.shape[1]), t2f.shape[0])
        f_first, ix_first = np.unique(flat, return_index=True)
        f_last, ix_rev = np.unique(flat[::-1], return_index=True)
        out = np.zeros((2, flat.max() + 1), dtype=np.int32)
        out[0, f_first] = owner[ix_first]
        out[1, f_last] = owner[flat.shape[0] - ix_rev - 1]
        out[1, out[0] == out[1]] = -1
        m._f2t = out
    return m._f2t


_NREF = {2: np.array([[0., -1.], [1., 1.], [-1., 0.]]),
         3: np.array([[0., 0., -1.], [0., -1., 0.], [-1., 0., 0.], [1., 1., 1.]])}
_BREFDOM = {"tri": "line", "tet": "tri"}


def facet_basis(m, elem, intorder=None, facets=None, side=0, quadrature_rule=None):
    if m.refdom not in _BREFDOM:
        raise NotImplementedError("affine meshes only")
    if m.refdom != elem.refdom:
        raise ValueError("Incompatible Mesh and Element.")
    edofs, N = dofs(m, elem)
    fac, t2f = facets_of(m)
    f2t = f2t_of(m)
    if quadrature_rule is not None:
        X, W = quadrature_rule
    else:
        X, W = quadrature(_BREFDOM[m.refdom],
                          intorder if intorder is not None else 2 * elem.maxdeg)
    find = (np.nonzero(f2t[1] == -1)[0].astype(np.int32) if facets is None
            else np.asarray(facets))
    tind, tind_n = f2t[side, find], f2t[0, find]
    p, dim, nqp, nf = m.p, m.p.shape[0], W.shape[-1], len(find)
    # boundary mapping (mapping_affine.py:154-181)
    B = np.empty((dim, dim - 1, fac.shape[1]))
    c = np.empty((dim, fac.shape[1]))
    for i in range(dim):
        c[i] = p[i, fac[0]]
        for j in range(dim - 1):
            B[i, j] = p[i, fac[j + 1]] - p[i, fac[0]]
    if dim == 2:
        detB = np.sqrt(B[0, 0] ** 2 + B[1, 0] ** 2)
    else:
        detB = np.sqrt((B[1, 0] * B[2, 1] - B[2, 0] * B[1, 1]) ** 2 +
                       (-B[0, 0] * B[2, 1] + B[2, 0] * B[0, 1]) ** 2 +
                       (B[0, 0] * B[1, 1] - B[1, 0] * B[0, 1]) ** 2)
    geo = affine_geometry(m)
    x = (np.einsum('ijk,jl', B[:, :, find], X).T + c[:, find].T).T      # G, :234-246
    Y = np.einsum('ijk,jkl->ikl', geo.invA[:, :, tind],
                  (x.T - geo.b[:, tind].T).T)                           # invF, :195-203
    # normals (:248-281)
    inv_n = np.einsum('ijk,l->ijkl', geo.invA[:, :, tind_n], np.ones(nqp))
    Nf = np.empty((dim, nf))
    for itr in range(_NREF[dim].shape[0]):
        ix = np.nonzero(t2f[itr, tind_n] == find)[0]
        for jtr in range(dim):
            Nf[jtr, ix] = _NREF[dim][itr, jtr]
    n = np.einsum('ijkl,ik->jkl', inv_n, Nf)
    n = np.einsum('ijk,jk->ijk', n, 1. / np.sqrt(np.sum(n ** 2, axis=0)))
    inv = np.einsum('ijk,l->ijkl', geo.invA[:, :, tind], np.ones(nqp))
    scalar = []
    for j in range(elem.nbf):                        # element_h1.py:10-18, 3-D X
        phi, dphi = elem.lbasis(Y, j)
        scalar.append(Field(np.broadcast_to(phi, (nf, nqp)),
                            np.einsum('ijkl,ikl->jkl', inv, dphi)))
    if elem.vector:
        basis = []
        for i in range(elem.nbf * elem.dim):
            f, k = scalar[i // elem.dim], i % elem.dim
            val = np.zeros((elem.dim,) + f.shape)
            val[k] = np.array(f)
            grd = np.zeros((elem.dim,) + f.grad.shape)
            grd[k] = f.grad
            basis.append(Field(val, grd))
    else:
        basis = scalar
    detDG = np.tile(detB[find], (nqp, 1)).T
    dx = np.abs(detDG) * np.broadcast_to(W, (nf, nqp))                  # facet_basis.py:114
    h = np.abs(detDG) ** (1. / (dim - 1.))
    return SimpleNamespace(mesh=m, elem=elem, X=X, W=W, basis=basis, dx=dx,
                           element_dofs=edofs[:, tind], N=N, Nbfun=len(basis),
                           nelems=nf, x=Field(x), h=Field(h), extra=dict(n=Field(n)),
                           find=find, tind=tind, Y=Y, dofs_full=edofs)


def _params(basis, kw):
    return _W(x=basis.x, h=basis.h, **getattr(basis, "extra", {}),
              **normalize_kwargs(kw, basis))


def interpolate(basis, w):
    """AbstractBasis.interpolate (abstract_basis.py:271-322), scalar H1."""
    if w.shape[0] != basis.N:
        raise ValueError("Input array has wrong size.")

    def linear_combination(get):
        out = 0. * np.einsum('...,...j->...j', w[basis.element_dofs[0]], get(basis.basis[0]))
        for i in range(basis.Nbfun):
            out += np.einsum('...,...j->...j', w[basis.element_dofs[i]], get(basis.basis[i]))
        return out
    return Field(linear_combination(lambda f: np.array(f)),
                 linear_combination(lambda f: f.grad))


def normalize_kwargs(kw, basis):
    """Form._normalize_asm_kwargs (assembly/form/form.py:91-121)."""
    out = {}
    for k, v in kw.items():
        if isinstance(v, np.ndarray) and v.ndim == 1:
            out[k] = interpolate(basis, v)
        elif isinstance(v, np.ndarray) and not isinstance(v, Field):
            out[k] = Field(v)
        else:
            out[k] = v
    return out


def functional_elemental(form, basis, **kw):
    """Functional.elemental / _kernel (assembly/form/functional.py:19-36)."""
    w = _params(basis, kw)
    return (form(w) * basis.dx).sum(-1)


# --------------------------------------------------------------------------
# helpers (helpers.py:22-150) and the models library
# (models/poisson.py:7-24, models/elasticity.py:7-53)
# --------------------------------------------------------------------------
def grad(u):
    return u.grad


def dot(u, v):
    return np.einsum('i...,i...', u, v)


def ddot(u, v):
    return np.einsum('ij...,ij...', u, v)


def transpose(T):
    return np.einsum('ij...->ji...', T)


def trace(T):
    return np.einsum('ii...', T)


def sym_grad(u):
    return .5 * (u.grad + transpose(u.grad))


def eye(w, n):
    return np.array([[w if i == j else 0. * w for i in range(n)]
                     for j in range(n)])


def identity(w, N=None):                        # helpers.py:153-161
    if N is None:
        N = w.shape[-3]
    return eye(np.ones(w.shape[-2:]), N)


def div(u):                                     # helpers.py:27-36
    if len(u.grad.shape) == 4:
        return np.einsum('ii...', u.grad)
    return u.grad[0]


def mul(A, x):                                  # helpers.py:132-134
    return np.einsum('ij...,j...->i...', A, x)


def det(A):                                     # helpers.py:164-176
    if A.shape[0] == 3:
        return A[0, 0] * (A[1, 1] * A[2, 2] - A[1, 2] * A[2, 1]) - \
            A[0, 1] * (A[1, 0] * A[2, 2] - A[1, 2] * A[2, 0]) + \
            A[0, 2] * (A[1, 0] * A[2, 1] - A[1, 1] * A[2, 0])
    return A[0, 0] * A[1, 1] - A[1, 0] * A[0, 1]


def inv(A):                                     # helpers.py:179-207
    out = np.zeros_like(A)
    d = det(A)
    if A.shape[0] == 3:
        out[0, 0] = (-A[1, 2] * A[2, 1] + A[1, 1] * A[2, 2]) / d
        out[1, 0] = (A[1, 2] * A[2, 0] - A[1, 0] * A[2, 2]) / d
        out[2, 0] = (-A[1, 1] * A[2, 0] + A[1, 0] * A[2, 1]) / d
        out[0, 1] = (A[0, 2] * A[2, 1] - A[0, 1] * A[2, 2]) / d
        out[1, 1] = (-A[0, 2] * A[2, 0] + A[0, 0] * A[2, 2]) / d
        out[2, 1] = (A[0, 1] * A[2, 0] - A[0, 0] * A[2, 1]) / d
        out[0, 2] = (-A[0, 2] * A[1, 1] + A[0, 1] * A[1, 2]) / d
        out[1, 2] = (A[0, 2] * A[1, 0] - A[0, 0] * A[1, 2]) / d
        out[2, 2] = (-A[0, 1] * A[1, 0] + A[0, 0] * A[1, 1]) / d
    else:
        out[0, 0], out[0, 1] = A[1, 1] / d, -A[0, 1] / d
        out[1, 0], out[1, 1] = -A[1, 0] / d, A[0, 0] / d
    return out


def curl(u):                                    # helpers.py:39-57
    g = u.grad
    if g.ndim == 3 and g.shape[0] == 2:
        return np.array([g[1], -g[0]])
    if g.ndim == 4 and g.shape[0] == 2:
        return g[1, 0] - g[0, 1]
    return np.array([g[2, 1] - g[1, 2], g[0, 2] - g[2, 0], g[1, 0] - g[0, 1]])


def cross(A, B):                                # helpers.py:210-218
    if A.shape[0] == 2:
        return A[0] * B[1] - A[1] * B[0]
    return np.array([A[1] * B[2] - A[2] * B[1], A[2] * B[0] - A[0] * B[2],
                     A[0] * B[1] - A[1] * B[0]])


def curluv(u, v, _):                            # models/general.py:15-17
    return dot(curl(u), v)


def rot(v, w):                                  # models/general.py:20-22
    return dot(curl(v), w['w'])


def vrot(v, w):                                 # models/general.py:25-27
    return dot(v, curl(w['w']))


def divu(u, v, _):                              # models/general.py:7-9
    return div(u) * v


def laplace(u, v, _):
    return dot(grad(u), grad(v))


def vector_laplace(u, v, _):
    return ddot(grad(u), grad(v))


def mass(u, v, _):
    return u * v


def unit_load(v, _):
    return v


def lame_parameters(E, nu):
    return (E * nu / ((1. + nu) * (1. - 2. * nu)), E / (2. * (1. + nu)))


def linear_elasticity(Lambda=1., Mu=1.):
    def C(T):
        return 2. * Mu * T + Lambda * eye(trace(T), T.shape[0])

    def weakform(u, v, w):
        return ddot(C(sym_grad(u)), sym_grad(v))
    return weakform


# --------------------------------------------------------------------------
# forms -> COO -> CSR / vector
# --------------------------------------------------------------------------
class _W(dict):
    def __getattr__(self, k):
        return self[k]


def bilinear_coo(form, basis, nthreads=0, vbasis=None, **kw):
    """BilinearForm._assemble (assembly/form/bilinear_form.py:58-128);
    ``nthreads > 0`` restates the reference's Python-thread split of the
    Nbfun^2 loop (:100-119,153-161).  ``basis`` is the trial (u) basis,
    ``vbasis`` the optional test basis."""
    ub = basis
    vb = basis if vbasis is None else vbasis
    if ub.X.shape[-1] != vb.X.shape[-1]:
        raise ValueError("Quadrature mismatch: trial and test functions "
                         "should have same number of integration points.")
    nt, nb, nbv = ub.nelems, ub.Nbfun, vb.Nbfun
    w = _params(ub, kw)
    data = np.zeros((nb, nbv, nt))
    rows = np.zeros(nb * nbv * nt, dtype=np.int32)
    cols = np.zeros(nb * nbv * nt, dtype=np.int32)

    def kernel(j, i):
        data[j, i, :] = np.sum(form(ub.basis[j], vb.basis[i], w)
                               * ub.dx, axis=1)            # :150-151

    for j in range(nb):
        for i in range(nbv):
            ixs = slice(nt * (nbv * j + i), nt * (nbv * j + i + 1))
            rows[ixs] = vb.element_dofs[i]
            cols[ixs] = ub.element_dofs[j]
            if nthreads <= 0:
                kernel(j, i)
    if nthreads > 0:
        from threading import Thread
        pairs = np.array([[i, j] for j in range(nb) for i in range(nbv)])
        threads = [Thread(target=lambda ix: [kernel(j, i) for i, j in ix], args=(ix,))
                   for ix in np.array_split(pairs, nthreads, axis=0)]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
    return np.array([rows, cols]), data.flatten('C'), (vb.N, ub.N)


def coo_to_csr(indices, data, shape):
    """COOData._assemble_scipy_csr (assembly/form/coo_data.py:27-36)."""
    K = coo_matrix((data, (indices[0], indices[1])), shape=shape)
    K.eliminate_zeros()
    return K.tocsr()


def assemble_bilinear(form, basis, **kw):
    return coo_to_csr(*bilinear_coo(form, basis, **kw))


def linear_coo(form, basis, **kw):
    """LinearForm._assemble (assembly/form/linear_form.py:18-49)."""
    nt, nb = basis.nelems, basis.Nbfun
    w = _params(basis, kw)
    data = np.zeros(nb * nt)
    rows = np.zeros(nb * nt, dtype=np.int32)
    for i in range(nb):
        ixs = slice(nt * i, nt * (i + 1))
        rows[ixs] = basis.element_dofs[i]
        data[ixs] = np.sum(form(basis.basis[i], w) * basis.dx, axis=1)
    return np.array([rows]), data, (basis.N,)


def assemble_linear(form, basis, **kw):
    """COOData.toarray 1-tensor branch (coo_data.py:102-108)."""
    idx, data, shape = linear_coo(form, basis, **kw)
    return coo_matrix((data, (idx[0], np.zeros_like(idx[0]))),
                      shape=shape + (1,)).toarray().T[0]


def assemble_bilinear_chunked(form, m, elem, chunk=100000, **kw):
    """concatenate chunk COO blocks, one tocsr()
    (assembly/__init__.py:92-95, coo_data.py:79-91)."""
    nel = m.t.shape[1]
    idx, dat, shape = [], [], None
    for s in range(0, nel, chunk):
        b = cell_basis(m, elem, elements=np.arange(s, min(nel, s + chunk)))
        i, d, shape = bilinear_coo(form, b, **kw)
        idx.append(i)
        dat.append(d)
    return coo_to_csr(np.hstack(idx), np.hstack(dat), shape)


def pairwise_sum(a):
    """numpy's pairwise summation of a contiguous 1-D double array
    (numpy/_core/src/umath/loops_utils.h.src DOUBLE_pairwise_sum), restated
    scalar-wise so CUDA's in-kernel reduction order can be checked against
    ``np.sum``."""
    n = len(a)
    if n < 8:
        r = 0.
        for v in a:
            r = r + v
        return r
    if n <= 128:
        r = [a[k] for k in range(8)]
        i = 8
        while i < n - (n % 8):
            for k in range(8):
                r[k] = r[k] + a[i + k]
            i += 8
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while i < n:
            res = res + a[i]
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return pairwise_sum(a[:n2]) + pairwise_sum(a[n2:])


# --------------------------------------------------------------------------
# boundary conditions (utils.py:282-400, 462-603), restated entry by entry
# --------------------------------------------------------------------------
def enforce(A, b=None, x=None, D=None, diag=1.):
    """skfem.utils.enforce: rows D zeroed in place in the pattern, diagonal set."""
    A = A.tocsr().copy()
    if x is None:
        x = np.zeros(A.shape[0])
    for r in np.asarray(D):
        found = False
        for s in range(A.indptr[r], A.indptr[r + 1]):
            on_diag = A.indices[s] == r
            A.data[s] = diag if on_diag else 0.
            found |= bool(on_diag)
        assert found, "missing diagonal"
    if b is None:
        return A
    if hasattr(b, "tocsr"):
        return A, enforce(b, D=D, diag=0.)
    bout = b.copy()
    bout[D] = x[D]
    return A, bout


def condense(A, b=None, x=None, D=None):
    """skfem.utils.condense for D given: (A[I][:, I], b[I] - A[I][:, D] @ x[D], I),
    the matrix-vector product accumulated like scipy's csr_matvec."""
    from scipy.sparse import csr_matrix
    A = A.tocsr()
    n = A.shape[0]
    I = np.setdiff1d(np.arange(n, dtype=np.int32), D)
    if x is None:
        x = np.zeros(n)
    elif b is None:
        b = np.zeros_like(x)
    colmap = np.full(n, -1, dtype=np.int64)
    colmap[I] = np.arange(len(I))
    indptr, indices, data, bout = [0], [], [], []
    for r in I:
        y = 0.
        for s in range(A.indptr[r], A.indptr[r + 1]):
            c = A.indices[s]
            if colmap[c] >= 0:
                indices.append(colmap[c])
                data.append(A.data[s])
            else:
                y = y + A.data[s] * x[c]
        indptr.append(len(indices))
        if b is not None and not hasattr(b, "tocsr"):
            bout.append(b[r] - y)
    AII = csr_matrix((np.array(data), np.array(indices, dtype=np.int32),
                      np.array(indptr, dtype=np.int32)), shape=(len(I), len(I)))
    if b is None:
        return AII, None, I
    if hasattr(b, "tocsr"):
        return AII, condense(b, D=D)[0], I
    return AII, np.array(bout), I

"""Reference domains: local vertex / edge / facet tables.

These tables define the local DOF ordering of higher-order elements (edge k of
a tetrahedron joins local vertices ``RefTet.edges[k]``), so they must agree
with the reference (skfem/refdom.py:55-209) for DOF numbering to be bit-exact.
"""
import numpy as np


class Refdom:
    name = "abstract"
    nnodes = 0
    nfacets = 0
    nedges = 0
    facets = None
    edges = None
    brefdom = None

    @classmethod
    def dim(cls):
        return cls.p.shape[0]


class RefPoint(Refdom):
    name = "point"
    p = np.zeros((0, 1))
    nnodes = 1


class RefLine(Refdom):
    name = "line"
    p = np.array([[0., 1.]])
    nnodes, nfacets = 2, 2
    facets = [[0], [1]]
    brefdom = RefPoint


class RefTri(Refdom):
    name = "tri"
    p = np.array([[0., 1., 0.],
                  [0., 0., 1.]])
    nnodes, nfacets = 3, 3
    facets = [[0, 1], [1, 2], [0, 2]]
    brefdom = RefLine


class RefTet(Refdom):
    name = "tet"
    p = np.array([[0., 1., 0., 0.],
                  [0., 0., 1., 0.],
                  [0., 0., 0., 1.]])
    nnodes, nfacets, nedges = 4, 4, 6
    facets = [[0, 1, 2], [0, 1, 3], [0, 2, 3], [1, 2, 3]]
    edges = [[0, 1], [1, 2], [0, 2], [0, 3], [1, 3], [2, 3]]
    brefdom = RefTri


class RefQuad(Refdom):
    name = "quad"
    p = np.array([[0., 1., 1., 0.],
                  [0., 0., 1., 1.]])
    nnodes, nfacets = 4, 4
    facets = [[0, 1], [1, 2], [2, 3], [0, 3]]
    brefdom = RefLine


class RefHex(Refdom):
    name = "hex"
    # vertex k of the reference cube (note the reversed, "ones first" order)
    p = np.array([[1., 1., 1.], [1., 1., 0.], [1., 0., 1.], [0., 1., 1.],
                  [1., 0., 0.], [0., 1., 0.], [0., 0., 1.], [0., 0., 0.]]).T
    nnodes, nfacets, nedges = 8, 6, 12
    facets = [[0, 1, 4, 2], [0, 2, 6, 3], [0, 3, 5, 1],
              [2, 4, 7, 6], [1, 5, 7, 4], [3, 6, 7, 5]]
    edges = [[0, 1], [0, 2], [0, 3], [1, 4], [1, 5], [2, 4],
             [2, 6], [3, 5], [3, 6], [4, 7], [5, 7], [6, 7]]
    brefdom = RefQuad

"""ctypes binding of the C-ABI library (include/skfem_b200.h).

The shared object is built in-tree by ``__graft_entry__.build()`` (nvcc,
sm_100a).  There is NO fallback: if it is missing, importing the compute path
fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libskfem_b200.so")

SKB_MAP_AFFINE, SKB_MAP_ISO_HEX1 = 0, 1
FORM_LAPLACE, FORM_MASS, FORM_VECTOR_LAPLACE, FORM_ELASTICITY = 0, 1, 2, 3
LFORM_UNIT_LOAD = 0

SKB_EINVAL = -1
SKB_ETOOBIG = -2
ERRORS = {-1: "SKB_EINVAL: bad argument / unsupported combination",
          -2: "SKB_ETOOBIG: tables do not fit on-chip or index overflow",
          -3: "Zero Jacobian determinant"}


class SkbSpace(C.Structure):
    _fields_ = [("dim", C.c_int32), ("nnodes", C.c_int32), ("mapping", C.c_int32),
                ("nbs", C.c_int32), ("ncomp", C.c_int32), ("nqp", C.c_int32),
                ("npts", C.c_int64), ("nel_total", C.c_int64),
                ("p", C.c_void_p), ("t", C.c_void_p), ("tind", C.c_void_p),
                ("nel", C.c_int64),
                ("phi", C.c_void_p), ("dphi", C.c_void_p), ("W", C.c_void_p),
                ("mdphi", C.c_void_p), ("mphi", C.c_void_p), ("X", C.c_void_p)]


_P, _I64, _I32, _INT = C.c_void_p, C.c_int64, C.c_int32, C.c_int
_SP = C.POINTER(SkbSpace)
_PD = C.POINTER(C.c_double)

SIGNATURES = {
    "skb_local_bilinear": (_INT, [_SP, _INT, _PD, _P, _P]),
    "skb_local_bilinear_em": (_INT, [_SP, _INT, _PD, _P, _P]),
    "skb_local_linear": (_INT, [_SP, _INT, _PD, _P, _P]),
    "skb_local_hex_sumfact": (_INT, [_SP, _INT, _I32, C.POINTER(C.c_int32), _PD, _PD,
                                     C.POINTER(C.c_uint8), C.POINTER(C.c_uint8), _I32, _P, _P]),
    "skb_plan_scratch_bytes": (_I64, [_I64]),
    "skb_plan_symbolic": (_INT, [_P, _P, _I32, _I32, _I64, _I64, _I64, _P, _INT,
                                 _P, _P, _P, _P, _P, _P, _I64, C.POINTER(_I64), _P]),
    "skb_plan_finalize": (_INT, [_I64, _I64, _I64, _I64, _I64, _P, _P, _P,
                                 _P, _P, _P, _P, _P]),
    "skb_plan_rows_count": (_INT, [_P, _I32, _I32, _I64, _I64, _P, _INT, _P, _P, _P, _P, _P, _P,
                                   _P]),
    "skb_plan_rows_sort": (_INT, [_P, _P, _I32, _I32, _I64, _I64, _P, _P, _P, _P, _P, _P, _P, _P,
                                  _P, _P, _P, _P, _P, _P, _P]),
    "skb_plan_rows_emit": (_INT, [_I64, _I64, _I64, _P, _P, _P, _P, _P, _P, _P]),
    "skb_mesh_tensor": (_INT, [_P, _P, _P, _I32, _I32, _I32, _I32, _I32, _P, _P, _P, _P]),
    "skb_element_dofs": (_INT, [_P, _I32, _I64, _P, _P]),
    "skb_entity_masks": (_INT, [_P, _I32, _I32, _I64, _P, _P, _P, _P]),
    "skb_plan_slot_of_entry": (_INT, [_P, _P, _I64, _P, _P]),
    "skb_csr_reduce": (_INT, [_P, _P, _P, _I64, _P, _P]),
    "skb_csr_reduce_em": (_INT, [_P, _I64, _I32, _I32, _P, _P, _I64, _P, _P]),
    "skb_vec_reduce": (_INT, [_P, _P, _P, _P, _I64, _P, _P]),
    "skb_tabulate": (_INT, [_SP, _INT, _P, _P, _P, _P, _P]),
    "skb_mapping": (_INT, [_SP, _P, _P, _P, _P]),
    "skb_qp_reduce": (_INT, [_P, _P, _I64, _I32, _INT, _P, _P]),
    "skb_p1tet_laplace_fused": (_INT, [_P, _I64, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32,
                                       _I32, C.c_double, _I32, _P, _P, _P]),
    "skb_p1_fused_smem_bytes": (_I64, [_I32, _I32, _I32, _I32]),
    "skb_p1tet_laplace_fused2": (_INT, [_P, _I64, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32,
                                        _I32, _I32, _I32, _I32, _I32, C.c_double, _I32, _P, _P,
                                        _P, _P, _P]),
    "skb_p1tet_mass_fused2": (_INT, [_P, _P, _I64, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32,
                                     _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P]),
    "skb_p1_fused2_smem_bytes": (_I64, [_I32, _I32, _I32, _I32, _I32]),
    "skb_p1_combine2": (_INT, [_P, _P, _P, _P, _I64, _P, _P]),
    "skb_pack_interface": (_INT, [_P, _P, _I64, _P, _P]),
    "skb_unpack_add_interface": (_INT, [_P, _P, _P, _I64, _P]),
    "skb_debug_flags": (None, [_INT]),
    "skb_sm_reserve": (None, [_INT]),
    "skb_l2_window": (_INT, [_P, _I64, _P]),
    "skb_p1_combine": (_INT, [_P, _P, _P, _P, _I64, _P, _P]),
    "skb_p1_plan_spread": (_INT, [_P, _P, _P, _I64, _I32, _P]),
    "skb_p1_plan_renumber": (_INT, [_P, _P, _I32, _I32, _P]),
    "skb_facet_geometry": (_INT, [_SP, _P, _I64, _P, _P, _P, _P, _I64, _P, _P, _I32,
                                  _P, _P, _P, _P, _P, _P]),
    "skb_facet_basis": (_INT, [_SP, _P, _I64, _I32, _P, _P, _P, _P, _I32, _P, _P, _P]),
    "skb_csr_enforce": (_INT, [_P, _P, _P, _P, _I64, C.c_double, _P, _P]),
    "skb_csr_condense_count": (_INT, [_P, _P, _P, _I64, _P, _P, _P]),
    "skb_csr_condense_fill": (_INT, [_P, _P, _P, _P, _I64, _P, _P, _P, _P, _P, _P, _P, _P]),
    "skb_csr_spmv": (_INT, [_P, _P, _P, _P, _P, _I64, _P]),
    "skb_launch_count": (_I64, [_INT]),
    "skb_version": (C.c_char_p, []),
}

_lib = None


class NativeLibraryMissing(ImportError):
    pass


def lib():
    """Load (once) and return the native library; raise if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryMissing(
                "skfem_b200: native library {} not found. Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` - there is "
                "no CPU fallback.".format(LIB_PATH))
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


class nvtx:
    """NVTX range around a seam of the pipeline (visible in nsys / ncu timelines); the
    reference only logs at these seams (form.py:76,79; bilinear_form.py:145-147;
    cell_basis.py:80,106).  A no-op where NVTX is unavailable."""

    def __init__(self, name):
        self.name = name
        self.on = False

    def __enter__(self):
        try:
            import torch
            if torch.cuda.is_available():
                torch.cuda.nvtx.range_push(self.name)
                self.on = True
        except Exception:
            self.on = False
        return self

    def __exit__(self, *exc):
        if self.on:
            import torch
            torch.cuda.nvtx.range_pop()
        return False


def check(code, what=""):
    if code == 0:
        return
    if code == -3:
        raise Exception("Zero Jacobian determinant")
    if code < 0:
        raise ValueError("{}: {}".format(what, ERRORS.get(code, code)))
    raise RuntimeError("{}: CUDA error {}".format(what, code))

"""Forms and their assembly: ``BilinearForm``, ``LinearForm``, ``Functional``,
``COOData``, ``asm`` - the reference's ``skfem.assembly.form`` surface
(skfem/assembly/form/form.py:28-121, bilinear_form.py:17-161,
linear_form.py:10-49, functional.py:11-61, coo_data.py:19-242,
skfem/assembly/__init__.py:63-97) on top of the CUDA engine.

Pipeline of ``form.assemble(basis)``:

1. element-local data ``(Nbu, Nbv, nel)`` on the device
   - library forms (``skfem_b200.models``): one fused kernel
     (``skb_local_bilinear`` / ``skb_local_linear``),
   - user forms: *traced* - the Python callable runs once per local entry on
     device fields, the quadrature reduction is ``skb_qp_reduce``;
2. sparsity plan (cached per basis/form): ``skb_plan_symbolic`` +
   ``skb_plan_finalize`` - value-dependent pattern like scipy's
   ``eliminate_zeros().tocsr()``;
3. numeric phase ``skb_csr_reduce`` / ``skb_vec_reduce`` - deterministic
   segmented sums;
4. the result is handed back as ``scipy.sparse.csr_matrix`` / ``ndarray`` or,
   with ``assemble_device``, kept on the GPU as :class:`DeviceCSR`.
"""
from __future__ import annotations

import ctypes as C
import logging
import numbers
import warnings
from copy import deepcopy
from dataclasses import dataclass, replace
from functools import partial
from inspect import signature
from typing import Any, Optional, Tuple

import numpy as np

from . import _lib
from . import hex_sumfact as _hex_sumfact
from .field import DeviceArray, DiscreteField, combine_layout, layout_of

logger = logging.getLogger(__name__)


def _torch():
    import torch
    return torch


# Engine options (set_options):
#   fused             use the fused P1-tet Laplace path for warm re-assembly
#   fused_tile        elements per tile (128 | 256 | 384 | 512 | 768, csrc/skb_p1_fused.cu)
#   fused_threads     reduce threads per CTA (with fused_tile selects the kernel variant)
#   fused_ring        record buffers in shared memory (4 | 5)
#   fused_arith       "exact": reference operation order, bit-identical local data;
#                     "fast": FMA + one reciprocal per element in the warm fused kernel
#                     (values within a few ulp per term, inside the rtol 1e-12 bar; the
#                     pattern is unaffected, it always comes from the exact cold pass)
#   fused_spread      plan-time bank spreading of the P2 lists       (csrc/skb_p1_plan.cu)
#   fused_renumber    plan-time conflict-free tile-local vertex ids  (csrc/skb_p1_plan.cu)
#   fused_l2_persist  L2 persisting window on the per-tile partial sums between the fused
#                     kernel and the combine kernel
#   fused_tiling      "morton": tiles = a Morton curve over the element centroids cut every
#                     fused_tile elements; "kd": compact boxes from a balanced k-d tree
#                     (fused._kd_order; fewer CSR slots shared between tiles - plan-only
#                     change, not yet timed on a B200, hence opt-in)
#   hex_sumfact      ElementHex2 laplace / mass at the default rule by sum factorisation
#                     (csrc/skb_hex_sf.cu); False: the Gram-matrix tensor-core kernel
#   element_major    warm generic path: kernels that can (Hex2 sum factorisation) write the
#                     local data element-major and skb_csr_reduce_em gathers whole sectors
#   fused_version    2: super-tile / pool kernel (csrc/skb_p1_fused2.cu, fused2.py; options
#                     fused2_tile, fused2_ring, fused2_pool, fused2_ctas, fused2_S);
#                     1: the first-generation warp-specialised kernel (options above)
_CONFIG = {"fused": True, "fused_tile": 512, "fused_threads": 480, "fused_ring": 4,
           "fused_arith": "exact", "fused_spread": True, "fused_renumber": True,
           "fused_l2_persist": True, "fused_tiling": "morton",
           "fused_version": 2, "fused2_tile": 256, "fused2_ring": 3, "fused2_pool": 2048,
           "fused2_ctas": 0, "fused2_S": None, "fused2_ept": 1, "plan_method": "rows",
           "hex_sumfact": True, "element_major": True}


def set_options(**kw):
    """Set engine options (see the table above ``_CONFIG``); unknown names raise KeyError.
    Options that shape the fused plan (tile, threads, ring, spread, renumber) apply to
    plans built afterwards."""
    for k, v in kw.items():
        if k not in _CONFIG:
            raise KeyError(k)
        _CONFIG[k] = v


def use_fused():
    return bool(_CONFIG["fused"])


def fused_tile():
    return int(_CONFIG["fused_tile"])


class FormExtraParams(dict):
    """Passed to forms as 'w'."""

    def __getattr__(self, attr):
        if attr in self:
            return self[attr]
        raise AttributeError("Attribute '{}' not found in 'w'.".format(attr))


# ---------------------------------------------------------------------------
# device CSR + sparsity plan
# ---------------------------------------------------------------------------
class DeviceCSR:
    """CSR matrix resident on the GPU: ``indptr``/``indices`` int32, ``data``
    float64 torch tensors (canonical format: sorted, no duplicates)."""

    def __init__(self, indptr, indices, data, shape):
        self.indptr, self.indices, self.data, self.shape = indptr, indices, data, tuple(shape)

    @property
    def nnz(self):
        return int(self.data.shape[0])

    def to_scipy(self):
        from scipy.sparse import csr_matrix
        torch = _torch()

        def d2h(t):  # pinned staging + async copy: the three arrays overlap
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t, non_blocking=True)
            return h
        hd, hi, hp = d2h(self.data), d2h(self.indices), d2h(self.indptr)
        torch.cuda.current_stream().synchronize()
        A = csr_matrix((hd.numpy(), hi.numpy(), hp.numpy()), shape=self.shape, copy=False)
        A.has_sorted_indices = True
        A.has_canonical_format = True
        return A

    tocsr = to_scipy

    def copy(self):
        """New values, shared (immutable) pattern."""
        return DeviceCSR(self.indptr, self.indices, self.data.clone(), self.shape)

    @classmethod
    def from_scipy(cls, A, device=None):
        torch = _torch()
        from .basis import default_device
        device = default_device() if device is None else device
        A = A.tocsr()
        A.sort_indices()
        return cls(torch.as_tensor(A.indptr.astype(np.int32), device=device),
                   torch.as_tensor(A.indices.astype(np.int32), device=device),
                   torch.as_tensor(A.data.astype(np.float64), device=device), A.shape)

    def to_torch(self):
        torch = _torch()
        return torch.sparse_csr_tensor(self.indptr.long(), self.indices.long(), self.data,
                                       size=self.shape)

    def matvec(self, x):
        return self.to_torch() @ x


class PatternChanged(RuntimeError):
    """A warm re-assembly found local entries whose zero / nonzero status differs from the
    cached plan's: the reference would produce another CSR pattern (coo_data.py:35)."""


class Plan:
    """indptr/indices + the entry->slot permutation of one (basis, form)."""
    __slots__ = ("indptr", "indices", "segptr", "perm", "nnz", "nkeep", "shape", "ncoo")


def _stream():
    return C.c_void_p(_torch().cuda.current_stream().cuda_stream)


_TOKENS = iter(range(1, 1 << 62))


def _basis_token(basis):
    """Process-unique integer of a basis object (never reused, unlike ``id``)."""
    tok = getattr(basis, "_token", None)
    if tok is None:
        tok = basis._token = next(_TOKENS)
    return tok


MASK_PLANS_KEPT = 4


def _mask_plan_get(ubasis, vbasis, nz):
    """Plan of a traced form whose local data has exactly the zero mask ``nz`` (bool device
    tensor (Nbu, Nbv, nel)) on this pair of bases, or None.  The plan (indptr / indices /
    entry -> slot permutation, coo_data.py:34-36) depends on the form only through that mask,
    so forms are interchangeable here and a form whose coefficients changed is re-planned
    exactly when its pattern did."""
    torch = _torch()
    tok = _basis_token(vbasis)
    for ent in ubasis._plans.get("by-mask", ()):
        if ent[0] == tok and ent[1].shape == nz.shape and bool(torch.equal(ent[1], nz)):
            return ent[2]
    return None


def _mask_plan_put(ubasis, vbasis, nz, plan):
    lst = ubasis._plans.setdefault("by-mask", [])
    lst.append((_basis_token(vbasis), nz, plan))
    del lst[:-MASK_PLANS_KEPT]


def _build_plan_rows(plan, dofs_v, dofs_u, nel, nrows, local, drop_zeros, mask=None):
    """Row buckets + per-row sorts (csrc/skb_plan_rows.cu).  Returns False when a limit of
    that path is hit (the caller then takes the radix-sort path)."""
    torch = _torch()
    lib = _lib.lib()
    dev = dofs_v.device
    nbv = dofs_v.shape[0]
    nbu = 1 if dofs_u is None else dofs_u.shape[0]
    i32, i64 = torch.int32, torch.int64

    def u32(n):
        return torch.empty(max(int(n), 1), dtype=i32, device=dev)
    # mask (optional, int32 tensor of nbv * nel words): the surviving columns of every
    # incidence, supplied by the caller instead of being derived from `local`
    drop = 2 if mask is not None else (1 if drop_zeros else 0)
    if mask is None:
        mask = u32(nbv * nel)
    rc = torch.empty(nrows, dtype=i64, device=dev)
    sums = u32(4096)
    incstart, candstart = u32(nrows + 1), u32(nrows + 1)
    counts = (C.c_int64 * 2)()
    code = lib.skb_plan_rows_count(
        dofs_v.data_ptr(), nbv, nbu, nel, nrows, None if local is None else local.data_ptr(),
        drop, mask.data_ptr(), rc.data_ptr(), sums.data_ptr(),
        incstart.data_ptr(), candstart.data_ptr(), counts, _stream())
    if code == _lib.SKB_ETOOBIG:
        return False
    _lib.check(code, "skb_plan_rows_count")
    nkeep, ninc = int(counts[0]), int(counts[1])
    cursor, nuniq, inc = u32(2 * nrows), u32(nrows), u32(2 * ninc)
    dofs_ut = u32(0 if dofs_u is None else nbu * nel)
    perm, ucol, uoff = u32(nkeep), u32(nkeep), u32(nkeep)
    indptr = torch.empty(nrows + 1, dtype=i32, device=dev)
    flag = torch.empty(3, dtype=i32, device=dev)
    nnz_c = C.c_int64(0)
    code = lib.skb_plan_rows_sort(
        dofs_v.data_ptr(), None if dofs_u is None else dofs_u.data_ptr(), nbv, nbu, nel, nrows,
        mask.data_ptr(), incstart.data_ptr(), candstart.data_ptr(), cursor.data_ptr(),
        inc.data_ptr(), dofs_ut.data_ptr(), sums.data_ptr(), perm.data_ptr(), ucol.data_ptr(),
        uoff.data_ptr(),
        nuniq.data_ptr(), indptr.data_ptr(), flag.data_ptr(), C.byref(nnz_c), _stream())
    if code == _lib.SKB_ETOOBIG:
        return False
    _lib.check(code, "skb_plan_rows_sort")
    nnz = int(nnz_c.value)
    plan.nnz, plan.nkeep = nnz, nkeep
    plan.indptr, plan.perm = indptr, perm
    plan.indices = torch.empty(nnz, dtype=i32, device=dev)
    plan.segptr = torch.empty(nnz + 1, dtype=i32, device=dev)
    code = lib.skb_plan_rows_emit(nrows, nnz, nkeep, candstart.data_ptr(), indptr.data_ptr(),
                                  ucol.data_ptr(), uoff.data_ptr(), plan.indices.data_ptr(),
                                  plan.segptr.data_ptr(), _stream())
    _lib.check(code, "skb_plan_rows_emit")
    return True


def build_plan(dofs_v, dofs_u, nel, shape, local, drop_zeros=True, method=None, mask=None):
    """The CSR structure + the per-slot lists of COO entries, built on the device:
    ``method`` "rows" (default, ``set_options(plan_method=...)``): row buckets and per-row
    sorts (skb_plan_rows_*); "sort": one global radix sort (skb_plan_symbolic +
    skb_plan_finalize), also the fallback when a row is too long for the first.  Both give
    the same arrays bit for bit."""
    torch = _torch()
    lib = _lib.lib()
    dev = dofs_v.device
    nbv = dofs_v.shape[0]
    nbu = 1 if dofs_u is None else dofs_u.shape[0]
    ncoo = nbv * nbu * nel
    nrows = int(shape[0])
    ncols = int(shape[1]) if len(shape) > 1 else 1
    plan = Plan()
    plan.shape, plan.ncoo = tuple(shape), ncoo
    i32, i64 = torch.int32, torch.int64
    if ncoo == 0:
        plan.nnz = plan.nkeep = 0
        plan.indptr = torch.zeros(nrows + 1, dtype=i32, device=dev)
        plan.indices = torch.zeros(0, dtype=i32, device=dev)
        plan.segptr = torch.zeros(1, dtype=i32, device=dev)
        plan.perm = torch.zeros(0, dtype=i32, device=dev)
        return plan
    if mask is not None or (method or _CONFIG["plan_method"]) == "rows":
        if _build_plan_rows(plan, dofs_v, dofs_u, nel, nrows, local, drop_zeros, mask):
            return plan
        if mask is not None:
            return None                                  # (mesh entities: the caller falls back)
    keys_a = torch.empty(ncoo, dtype=i64, device=dev)
    keys_b = torch.empty(ncoo, dtype=i64, device=dev)
    vals_a = torch.empty(ncoo, dtype=i32, device=dev)
    vals_b = torch.empty(ncoo, dtype=i32, device=dev)
    slot = torch.empty(ncoo, dtype=i32, device=dev)
    tmp_bytes = int(lib.skb_plan_scratch_bytes(ncoo))
    tmp = torch.empty(tmp_bytes, dtype=torch.uint8, device=dev)
    counts = (C.c_int64 * 3)()
    code = lib.skb_plan_symbolic(
        dofs_v.data_ptr(), None if dofs_u is None else dofs_u.data_ptr(), nbv, nbu, nel,
        nrows, ncols, None if local is None else local.data_ptr(), 1 if drop_zeros else 0,
        keys_a.data_ptr(), keys_b.data_ptr(), vals_a.data_ptr(), vals_b.data_ptr(),
        slot.data_ptr(), tmp.data_ptr(), tmp_bytes, counts, _stream())
    _lib.check(code, "skb_plan_symbolic")
    nnz, nkeep, which = int(counts[0]), int(counts[1]), int(counts[2])
    ks, vs = (keys_a, vals_a) if which == 0 else (keys_b, vals_b)
    plan.nnz, plan.nkeep = nnz, nkeep
    plan.indptr = torch.empty(nrows + 1, dtype=i32, device=dev)
    plan.indices = torch.empty(nnz, dtype=i32, device=dev)
    plan.segptr = torch.empty(nnz + 1, dtype=i32, device=dev)
    plan.perm = torch.empty(max(nkeep, 1), dtype=i32, device=dev)
    code = lib.skb_plan_finalize(ncoo, nrows, ncols, nnz, nkeep, ks.data_ptr(), vs.data_ptr(),
                                 slot.data_ptr(), plan.indptr.data_ptr(),
                                 plan.indices.data_ptr(), plan.segptr.data_ptr(),
                                 plan.perm.data_ptr(), _stream())
    _lib.check(code, "skb_plan_finalize")
    return plan


# ---------------------------------------------------------------------------
# COOData
# ---------------------------------------------------------------------------
@dataclass
class COOData:
    """Element-local (COO) form of an assembled tensor, host numpy arrays with
    the reference's layout (coo_data.py:19-25): ``data`` flattened
    entry-major / element-minor, ``indices`` int32 ``[rows, cols]``."""
    indices: np.ndarray
    data: np.ndarray
    shape: Tuple[int, ...]
    local_shape: Optional[Tuple[int, ...]]

    def tolocal(self, basis=None):
        if self.local_shape is None:
            raise NotImplementedError("Cannot build local matrices if "
                                      "local_shape is not specified.")
        return np.moveaxis(self.data.reshape(self.local_shape + (-1,), order='C'), -1, 0)

    def fromlocal(self, local):
        return replace(self, data=np.moveaxis(local, 0, -1).flatten('C'))

    def __add__(self, other):
        if isinstance(other, int):
            return self
        return replace(self, indices=np.hstack((self.indices, other.indices)),
                       data=np.hstack((self.data, other.data)),
                       shape=tuple(max(a, b) for a, b in zip(self.shape, other.shape)),
                       local_shape=None)

    __radd__ = __add__

    def astuple(self):
        return self.indices, self.data, self.shape

    def _device_reduce(self):
        """Assemble arbitrary (possibly concatenated) COO triplets on the GPU."""
        torch = _torch()
        from .basis import default_device
        dev = default_device()
        data = torch.as_tensor(self.data, dtype=torch.float64, device=dev)
        rows = torch.as_tensor(self.indices[0].astype(np.int32), device=dev).reshape(1, -1)
        n = rows.shape[1]
        if len(self.shape) == 2:
            cols = torch.as_tensor(self.indices[1].astype(np.int32), device=dev).reshape(1, -1)
            plan = build_plan(rows, cols, n, self.shape, data, drop_zeros=True)
            out = torch.empty(plan.nnz, dtype=torch.float64, device=dev)
            _lib.check(_lib.lib().skb_csr_reduce(data.data_ptr(), plan.perm.data_ptr(),
                                                 plan.segptr.data_ptr(), plan.nnz,
                                                 out.data_ptr(), _stream()), "skb_csr_reduce")
            return DeviceCSR(plan.indptr, plan.indices, out, self.shape)
        plan = build_plan(rows, None, n, self.shape, None, drop_zeros=False)
        out = torch.empty(self.shape[0], dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().skb_vec_reduce(data.data_ptr(), plan.perm.data_ptr(),
                                             plan.segptr.data_ptr(), plan.indptr.data_ptr(),
                                             self.shape[0], out.data_ptr(), _stream()),
                   "skb_vec_reduce")
        return out

    def tocsr(self):
        return self._device_reduce().to_scipy()

    def toarray(self):
        if len(self.shape) == 1:
            return self._device_reduce().cpu().numpy()
        if len(self.shape) == 2:
            return self.tocsr().toarray()
        raise NotImplementedError

    def todefault(self):
        if len(self.shape) == 0:
            return np.sum(self.data, axis=0)
        if len(self.shape) == 1:
            return self.toarray()
        if len(self.shape) == 2:
            return self.tocsr()
        return self


# ---------------------------------------------------------------------------
# forms
# ---------------------------------------------------------------------------
class Form:
    form = None
    native = None  # ("bilinear"|"linear", kernel id, params, "scalar"|"vector")

    def __init__(self, form=None, dtype=np.float64, nthreads=0, **params):
        self.form = form.form if isinstance(form, Form) else form
        if isinstance(form, Form):
            self.native = form.native
        self.nargs = len(signature(self.form).parameters) if self.form is not None else None
        if dtype not in (np.float64, float):
            raise NotImplementedError("skfem_b200 assembles in float64 only")
        self.dtype = dtype
        self.nthreads = nthreads  # accepted for API compatibility; the GPU path ignores it
        self.params = params

    def partial(self, *args, **kwargs):
        form = deepcopy(self)
        name = form.form.__name__
        form.form = partial(form.form, *args, **kwargs)
        form.form.__name__ = name
        form.native = None
        return form

    def __call__(self, *args):
        if self.form is None:  # used as a decorator factory
            return type(self)(form=args[0], dtype=self.dtype, nthreads=self.nthreads,
                              **self.params)
        return self.form(*args)

    # -- shared machinery -------------------------------------------------------------
    @staticmethod
    def _normalize_asm_kwargs(w, basis):
        """Accepted formats of the extra parameters (form.py:91-121)."""
        torch = _torch()
        out = {}
        for k, v in w.items():
            if isinstance(v, DiscreteField):
                if v.shape[-1] != basis.X.shape[-1]:
                    raise ValueError("Quadrature mismatch: '{}' should have same number of "
                                     "integration points as the basis object.".format(k))
                out[k] = v
            elif isinstance(v, DeviceArray):
                out[k] = DiscreteField(v.t)
            elif isinstance(v, numbers.Number) or isinstance(v, tuple):
                out[k] = v
            elif torch.is_tensor(v):
                out[k] = basis.interpolate(v) if v.dim() == 1 else DiscreteField(v)
            elif isinstance(v, np.ndarray) and v.ndim == 1:
                out[k] = basis.interpolate(v)
            elif isinstance(v, np.ndarray) and v.ndim > 1:
                dev = basis._dev()["device"]
                out[k] = DiscreteField(torch.as_tensor(np.asarray(v, dtype=np.float64), device=dev))
            elif isinstance(v, list):
                warnings.warn("Use Basis.interpolate instead of passing lists to assemble",
                              DeprecationWarning)
                out[k] = v
            else:
                raise ValueError("The given type '{}' for the list of extra form parameters w "
                                 "cannot be converted to DiscreteField.".format(type(v)))
        return out

    def _wdict(self, basis, kwargs):
        return FormExtraParams({**basis.default_parameters(),
                                **self._normalize_asm_kwargs(kwargs, basis)})

    def _reduce_qp(self, basis, integrand, out_row):
        """out_row[e] = numpy-pairwise sum_q integrand[e, q] * dx[e, q]."""
        torch = _torch()
        nel, nqp = basis.nelems, basis.nqp
        t = integrand.t if isinstance(integrand, DeviceArray) else integrand
        if not torch.is_tensor(t):
            t = torch.as_tensor(t, dtype=torch.float64, device=out_row.device)
        if t.dtype != torch.float64:
            t = t.to(torch.float64)
        t = t.expand(nel, nqp).contiguous()
        # np.sum(form * dx, axis=1): pairwise if numpy's product is C-ordered,
        # left-to-right if it is Fortran-ordered (field.combine_layout)
        seq = combine_layout([layout_of(integrand), basis._dx_layout]) == "F"
        code = _lib.lib().skb_qp_reduce(t.data_ptr(), basis._dx_dev().data_ptr(), nel, nqp,
                                        1 if seq else 0, out_row.data_ptr(), _stream())
        _lib.check(code, "skb_qp_reduce")

    def assemble(self, *args, **kwargs) -> Any:
        raise NotImplementedError

    def elemental(self, *args, **kwargs):
        return self.coo_data(*args, **kwargs)


class BilinearForm(Form):
    """``a(u, v)``: ``form(u, v, w)`` integrated over the basis.

    >>> form = BilinearForm(lambda u, v, _: u * v)
    >>> form.assemble(Basis(MeshTri(), ElementTriP1())).toarray()  # 4x4 mass matrix
    """

    def _native_applicable(self, ubasis, vbasis, kwargs):
        if self.native is None or kwargs or (vbasis is not None and vbasis is not ubasis):
            return False
        if not getattr(ubasis, "_native_ok", True):   # FacetBasis: traced path
            return False
        kind, _, _, field = self.native
        if kind != "bilinear":
            return False
        if field == "scalar":
            return ubasis.ncomp == 1
        return ubasis.ncomp > 1

    def _local_element_major(self, ubasis):
        """Element-local data as (nel, Nbv, Nbu) for the warm generic path, or None when no
        kernel of this form writes that layout (ElementHex2 by sum factorisation; P1 / P2
        library forms on affine meshes at their default rules).  ``skb_csr_reduce_em`` then
        gathers whole sectors instead of one value per sector."""
        if not (_CONFIG["element_major"] and self._native_applicable(ubasis, None, {})):
            return None
        nb = ubasis.Nbfun
        if nb * nb * ubasis.nelems >= 2 ** 32 or ubasis._plans.get(("no-em", self.native[1])):
            return None
        torch = _torch()
        d = ubasis._dev()
        _, kid, params, _ = self.native
        out = torch.empty((ubasis.nelems, nb, nb), dtype=torch.float64, device=d["device"])
        tab = _hex_sumfact.tables(ubasis) if _CONFIG["hex_sumfact"] and kid in (
            _lib.FORM_LAPLACE, _lib.FORM_MASS) else None
        if tab is not None:
            code = _hex_sumfact.launch(_lib.lib(), d["space"], kid, tab, out.data_ptr(),
                                       _stream(), element_major=True)
            _lib.check(code, "skb_local_hex_sumfact")
            return out
        cparams = None if params is None else (C.c_double * len(params))(*params)
        code = _lib.lib().skb_local_bilinear_em(C.byref(d["space"]), kid, cparams,
                                                out.data_ptr(), _stream())
        if code == _lib.SKB_EINVAL:               # no element-major kernel for this space
            ubasis._plans[("no-em", kid)] = True
            return None
        _lib.check(code, "skb_local_bilinear_em")
        return out

    def _local(self, ubasis, vbasis=None, **kwargs):
        """Element-local data (Nbu, Nbv, nel) as a device tensor."""
        torch = _torch()
        if vbasis is None:
            vbasis = ubasis
        elif ubasis.X.shape[-1] != vbasis.X.shape[-1]:
            raise ValueError("Quadrature mismatch: trial and test functions "
                             "should have same number of integration points.")
        d = ubasis._dev()
        nel = ubasis.nelems
        out = torch.empty((ubasis.Nbfun, vbasis.Nbfun, nel), dtype=torch.float64,
                          device=d["device"])
        if self._native_applicable(ubasis, vbasis if vbasis is not ubasis else None, kwargs):
            _, kid, params, _ = self.native
            cparams = None if params is None else (C.c_double * len(params))(*params)
            if kid in (_lib.FORM_LAPLACE, _lib.FORM_MASS) and _CONFIG["hex_sumfact"]:
                # ElementHex2 at its default rule: sum factorisation (csrc/skb_hex_sf.cu)
                tab = _hex_sumfact.tables(ubasis)
                if tab is not None:
                    code = _hex_sumfact.launch(_lib.lib(), d["space"], kid, tab,
                                               out.data_ptr(), _stream())
                    _lib.check(code, "skb_local_hex_sumfact")
                    return out
            code = _lib.lib().skb_local_bilinear(C.byref(d["space"]), kid, cparams,
                                                 out.data_ptr(), _stream())
            _lib.check(code, "skb_local_bilinear")
            return out
        # traced path (bilinear_form.py:86-98 with device fields)
        w = self._wdict(ubasis, kwargs)
        ub = [ubasis._basis_field_dev(j) for j in range(ubasis.Nbfun)]
        vb = ub if vbasis is ubasis else [vbasis._basis_field_dev(i) for i in range(vbasis.Nbfun)]
        for j in range(ubasis.Nbfun):
            for i in range(vbasis.Nbfun):
                self._reduce_qp(ubasis, self.form(ub[j], vb[i], w), out[j, i])
        return out

    def _plan_key(self, ubasis, vbasis, kwargs):
        """Cache key of the sparsity plan - only for library forms without parameters: their
        local data, hence the value-dependent pattern, is a function of the basis alone.
        Traced forms (closures, ``w`` fields, ``partial`` copies) get no key: their plans
        are cached by *zero mask* (``_mask_plan``)."""
        if kwargs or self.native is None:
            return None
        return (self.native[:3], _basis_token(vbasis) if vbasis is not None else None)

    def assemble_device(self, ubasis, vbasis=None, out=None, slot_map=None,
                        **kwargs) -> DeviceCSR:
        """Assemble into a device-resident CSR (no host transfer).

        ``out``: optional float64 device tensor of length nnz that receives
        the values (warm calls only - the plan must exist).  With ``out`` the
        call allocates nothing and can be captured in a CUDA graph, which is
        how re-assembly loops (time stepping, Newton) should drive it.
        ``slot_map`` (with ``out``): int64 tensor, CSR slot -> index in ``out``;
        the kernels then scatter the values straight to those positions (used by
        the multi-GPU path to write [row block | send buffer] in one pass)."""
        assert self.form is not None
        torch = _torch()
        vb = ubasis if vbasis is None else vbasis
        key = self._plan_key(ubasis, vbasis, kwargs)
        plan = ubasis._plans.get(key) if key is not None else None
        if slot_map is not None and (out is None or plan is None):
            raise ValueError("slot_map= needs out= and an existing plan")
        # warm re-assembly of a fusable form: geometry -> CSR values in one
        # pass (csrc/skb_p1_fused.cu); the tile plan is built on the first
        # warm call from the pattern the cold call established
        if (plan is not None and vbasis is None and not kwargs and use_fused()
                and ubasis.nelems > 0 and plan.nnz > 0):
            from . import fused
            if fused.applicable(ubasis, self, version=int(_CONFIG["fused_version"])):
                fkey = ("fused", key) if slot_map is None else ("fused-mapped", key,
                                                                id(slot_map))
                fp = ubasis._plans.get(fkey, False)
                if fp is False and int(_CONFIG["fused_version"]) == 2:
                    from . import fused2
                    fp = fused2.build_auto(ubasis, plan, T=int(_CONFIG["fused2_tile"]),
                                           ring=int(_CONFIG["fused2_ring"]),
                                           pool_cap=int(_CONFIG["fused2_pool"]),
                                           S=_CONFIG["fused2_S"], slot_map=slot_map,
                                           spread=bool(_CONFIG["fused_spread"]),
                                           renumber=bool(_CONFIG["fused_renumber"]),
                                           ctas_per_sm=int(_CONFIG["fused2_ctas"]) |
                                           (int(_CONFIG["fused2_ept"]) << 8),
                                           form_id=self.native[1])
                    ubasis._plans[fkey] = fp
                if fp is False:
                    fp = fused.build_auto(ubasis, plan, T=fused_tile(),
                                          threads=int(_CONFIG["fused_threads"]),
                                          ring=int(_CONFIG["fused_ring"]), slot_map=slot_map,
                                          spread=bool(_CONFIG["fused_spread"]),
                                          renumber=bool(_CONFIG["fused_renumber"]),
                                          tiling=str(_CONFIG["fused_tiling"]))
                    ubasis._plans[fkey] = fp      # None: tiles too big, stay generic
                if fp is not None:
                    data = out if out is not None else torch.empty(
                        plan.nnz, dtype=torch.float64, device=fp.p.device)
                    if getattr(fp, "version", 1) == 2:
                        from . import fused2
                        with _lib.nvtx("skfem_b200:fused"):
                            fused2.run(fp, data, _stream(),
                                       fast=_CONFIG["fused_arith"] == "fast",
                                       l2_persist=bool(_CONFIG["fused_l2_persist"]))
                        # first run after basis.update_points: did the zero mask of any local
                        # matrix change?  Then this pattern is no longer the reference's.
                        if getattr(fp, "unchecked", False) and \
                                not torch.cuda.is_current_stream_capturing():
                            fp.unchecked = False
                            if fused2.pattern_changed(fp):
                                if out is not None:
                                    raise PatternChanged(
                                        "the sparsity pattern changed with the new vertex "
                                        "coordinates: assemble once without out=")
                                del ubasis._plans[fkey], ubasis._plans[key]
                                return self.assemble_device(ubasis, vbasis)
                    else:
                        fused.run(fp, data, _stream(), fast=_CONFIG["fused_arith"] == "fast",
                                  l2_persist=bool(_CONFIG["fused_l2_persist"]))
                    return DeviceCSR(plan.indptr, plan.indices, data, plan.shape)
        if (plan is not None and vbasis is None and not kwargs and slot_map is None
                and ubasis.nelems > 0 and plan.nnz > 0):
            # warm call of a form whose kernel can write element-major local data
            with _lib.nvtx("skfem_b200:local"):
                local = self._local_element_major(ubasis)
            if local is not None:
                data = out if out is not None else torch.empty(
                    plan.nnz, dtype=torch.float64, device=local.device)
                code = _lib.lib().skb_csr_reduce_em(
                    local.data_ptr(), ubasis.nelems, ubasis.Nbfun, ubasis.Nbfun,
                    plan.perm.data_ptr(), plan.segptr.data_ptr(), plan.nnz, data.data_ptr(),
                    _stream())
                _lib.check(code, "skb_csr_reduce_em")
                return DeviceCSR(plan.indptr, plan.indices, data, plan.shape)
        with _lib.nvtx("skfem_b200:local"):
            local = self._local(ubasis, vbasis, **kwargs)
        nz = None
        if plan is None and key is None:
            # traced form: the pattern is a function of (element_dofs, zero mask of the
            # local data) only - reuse a cached plan iff the mask is identical
            nz = local != 0
            plan = _mask_plan_get(ubasis, vb, nz)
        if plan is None:
            if out is not None:
                raise ValueError("out= needs an existing plan: assemble once without it")
            with _lib.nvtx("skfem_b200:plan"):
                plan = build_plan(vb._dev()["edofs"], ubasis._dev()["edofs"], ubasis.nelems,
                                  (vb.N, ubasis.N), local, drop_zeros=True)
            if key is not None:
                ubasis._plans[key] = plan
            else:
                _mask_plan_put(ubasis, vb, nz, plan)
        data = out if (out is not None and slot_map is None) else torch.empty(
            plan.nnz, dtype=torch.float64, device=local.device)
        code = _lib.lib().skb_csr_reduce(local.data_ptr(), plan.perm.data_ptr(),
                                         plan.segptr.data_ptr(), plan.nnz, data.data_ptr(),
                                         _stream())
        _lib.check(code, "skb_csr_reduce")
        if slot_map is not None:        # generic path: scatter through the map afterwards
            out[slot_map] = data
            data = out
        return DeviceCSR(plan.indptr, plan.indices, data, plan.shape)

    def assemble(self, ubasis, vbasis=None, **kwargs):
        """Assemble into ``scipy.sparse.csr_matrix`` (bilinear_form.py:130-148)."""
        logger.info("Assembling '{}'.".format(getattr(self.form, "__name__", "form")))
        with _lib.nvtx("skfem_b200:assemble"):
            A = self.assemble_device(ubasis, vbasis, **kwargs).to_scipy()
        logger.info("Assembling finished.")
        return A

    def coo_data(self, ubasis, vbasis=None, **kwargs) -> COOData:
        """Element-local matrices as host COO data (form.py:82-89)."""
        vb = ubasis if vbasis is None else vbasis
        local = self._local(ubasis, vbasis, **kwargs)
        nel = ubasis.nelems
        rows = np.tile(vb.element_dofs, (ubasis.Nbfun, 1)).reshape(-1)
        cols = np.repeat(ubasis.element_dofs, vb.Nbfun, axis=0).reshape(-1)
        assert rows.shape[0] == ubasis.Nbfun * vb.Nbfun * nel
        return COOData(np.array([rows, cols]), local.reshape(-1).cpu().numpy(),
                       (vb.N, ubasis.N), (vb.Nbfun, ubasis.Nbfun))

    def _assemble(self, ubasis, vbasis=None, **kwargs):
        c = self.coo_data(ubasis, vbasis, **kwargs)
        return c.indices, c.data, c.shape, c.local_shape


class LinearForm(Form):
    """``l(v)``: ``form(v, w)`` integrated over the basis."""

    def _local(self, basis, **kwargs):
        torch = _torch()
        d = basis._dev()
        out = torch.empty((basis.Nbfun, basis.nelems), dtype=torch.float64, device=d["device"])
        if (self.native is not None and not kwargs and self.native[0] == "linear"
                and basis.ncomp == 1 and getattr(basis, "_native_ok", True)):
            code = _lib.lib().skb_local_linear(C.byref(d["space"]), self.native[1], None,
                                               out.data_ptr(), _stream())
            _lib.check(code, "skb_local_linear")
            return out
        w = self._wdict(basis, kwargs)
        for i in range(basis.Nbfun):
            self._reduce_qp(basis, self.form(basis._basis_field_dev(i), w), out[i])
        return out

    def assemble_device(self, basis, vbasis=None, **kwargs):
        assert vbasis is None
        assert self.form is not None
        torch = _torch()
        local = self._local(basis, **kwargs)
        plan = basis._plans.get("linear")
        if plan is None:
            plan = build_plan(basis._dev()["edofs"], None, basis.nelems, (basis.N,), None,
                              drop_zeros=False)
            basis._plans["linear"] = plan
        vec = torch.empty(basis.N, dtype=torch.float64, device=local.device)
        code = _lib.lib().skb_vec_reduce(local.data_ptr(), plan.perm.data_ptr(),
                                         plan.segptr.data_ptr(), plan.indptr.data_ptr(),
                                         basis.N, vec.data_ptr(), _stream())
        _lib.check(code, "skb_vec_reduce")
        return vec

    def assemble(self, basis, vbasis=None, **kwargs):
        return self.assemble_device(basis, vbasis, **kwargs).cpu().numpy()

    def coo_data(self, basis, vbasis=None, **kwargs) -> COOData:
        assert vbasis is None
        local = self._local(basis, **kwargs)
        rows = basis.element_dofs.reshape(-1)
        return COOData(np.array([rows]), local.reshape(-1).cpu().numpy(), (basis.N,),
                       (basis.Nbfun,))

    def _assemble(self, basis, vbasis=None, **kwargs):
        c = self.coo_data(basis, vbasis, **kwargs)
        return c.indices, c.data, c.shape, c.local_shape


class Functional(Form):
    """Scalar functional ``form(w)`` integrated over the basis
    (functional.py:11-61)."""

    def elemental(self, basis, **kwargs):
        torch = _torch()
        if self.form is None:
            raise Exception("Form function handle not defined.")
        w = self._wdict(basis, kwargs)
        out = torch.empty(basis.nelems, dtype=torch.float64, device=basis._dev()["device"])
        self._reduce_qp(basis, self.form(w), out)
        return out.cpu().numpy()

    def assemble(self, basis, vbasis=None, **kwargs):
        assert vbasis is None
        return np.sum(self.elemental(basis, **kwargs), axis=0)

    def coo_data(self, basis, vbasis=None, **kwargs):
        return COOData(np.array([]), np.array([self.assemble(basis, **kwargs)]), (), ())


def asm(form, *args, to=None, **kwargs):
    """Shorthand for ``form.assemble`` (skfem/assembly/__init__.py:69-97).
    Bare callables are wrapped by argument count.  Lists of bases are summed
    through concatenated COO data, assembled once on the device."""
    if not isinstance(form, Form) and callable(form):
        nargs = form.__code__.co_argcount
        form = [Functional, LinearForm, BilinearForm][nargs - 1](form)
    assert form.form is not None
    if any(isinstance(a, list) for a in args):
        from itertools import product
        lists = [a if isinstance(a, list) else [a] for a in args]
        blocks = [form.coo_data(*combo, idx=ix, **kwargs)
                  for ix, combo in zip(product(*(range(len(x)) for x in lists)), product(*lists))]
        out = sum(blocks)
        return out.todefault() if to is None else to(blocks)
    # the reference hands the position of the basis in its (one-element) lists to the form
    # as w.idx (assembly/__init__.py:91-93); library forms with a dedicated kernel never
    # read w, for them the kwarg would only disable the kernel
    if getattr(form, "native", None) is None and "idx" not in kwargs:
        kwargs["idx"] = (0,) * len(args)
    if to is not None:
        return to([form.coo_data(*args, **kwargs)])
    return form.assemble(*args, **kwargs)

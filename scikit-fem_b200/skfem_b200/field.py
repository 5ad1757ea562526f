"""Device-resident arrays for the *traced* path of user-defined forms.

The reference hands a form callable numpy ``DiscreteField`` objects
(skfem/element/discrete_field.py:7-110): an ndarray of shape
``(..., nel, nqp)`` carrying ``.grad/.div/.curl/.hess`` that degrades to a
plain array under ufuncs and indexing.  Here the same callable is executed once
per local entry with ``DeviceArray``/``DiscreteField`` objects wrapping float64
CUDA tensors: every numpy ufunc / operator the form applies is executed as one
elementwise device op (torch is only the elementwise executor; reductions and
scatter are this package's own kernels), in the order the user wrote them, so
+,-,*,/ and sqrt round exactly like numpy's.

Supported numpy surface: arithmetic operators, all common unary/binary ufuncs,
``np.einsum`` (sequential row-major accumulation like numpy's), ``np.array`` /
``np.stack`` of device arrays, ``zeros_like/ones_like``, ``np.sum``,
``np.where``, indexing, ``.T``-free reshapes.  Transcendentals (sin, exp, pow
with non-trivial exponents) use CUDA's libdevice and may differ from the host
libm in the last ulp - documented in DESIGN.md.
"""
from __future__ import annotations

import numbers

import numpy as np
import torch

_UNARY = {
    "negative": torch.neg, "positive": lambda a: a, "absolute": torch.abs, "fabs": torch.abs,
    "sqrt": torch.sqrt, "square": lambda a: a * a, "exp": torch.exp, "expm1": torch.expm1,
    "log": torch.log, "log2": torch.log2, "log10": torch.log10, "log1p": torch.log1p,
    "sin": torch.sin, "cos": torch.cos, "tan": torch.tan, "arcsin": torch.asin,
    "arccos": torch.acos, "arctan": torch.atan, "sinh": torch.sinh, "cosh": torch.cosh,
    "tanh": torch.tanh, "sign": torch.sign, "floor": torch.floor, "ceil": torch.ceil,
    "reciprocal": torch.reciprocal, "isnan": torch.isnan, "isfinite": torch.isfinite,
    "logical_not": torch.logical_not, "cbrt": lambda a: torch.sign(a) * torch.abs(a) ** (1. / 3.),
}
_BINARY = {
    "add": torch.add, "subtract": torch.sub, "multiply": torch.mul, "divide": torch.div,
    "true_divide": torch.div, "maximum": torch.maximum, "minimum": torch.minimum,
    "arctan2": torch.atan2, "hypot": torch.hypot, "greater": torch.gt, "less": torch.lt,
    "greater_equal": torch.ge, "less_equal": torch.le, "equal": torch.eq,
    "not_equal": torch.ne, "logical_and": torch.logical_and, "logical_or": torch.logical_or,
    "fmax": torch.fmax, "fmin": torch.fmin,
}


def _power(a, b):
    # numpy evaluates x**2 as x*x and x**0.5 as sqrt(x); keep those exact
    if isinstance(b, numbers.Number):
        if b == 2:
            return a * a
        if b == 1:
            return a
        if b == 0.5:
            return torch.sqrt(a)
        if b == -1:
            return torch.reciprocal(a)
    return torch.pow(a, b)


_BINARY["power"] = _power
_BINARY["float_power"] = _power


def raw(x):
    """torch tensor / python scalar behind any operand."""
    if isinstance(x, DeviceArray):
        return x.t
    if isinstance(x, np.ndarray):
        if x.ndim == 0:
            return x.item()
        raise TypeError("mixing host numpy arrays with device fields inside a form; "
                        "wrap them with skfem_b200.asdevice(...) first")
    if isinstance(x, (list, tuple)):
        return _stack(x).t
    return x


def _like(x):
    for a in x:
        if isinstance(a, DeviceArray):
            return a.t
        if isinstance(a, (list, tuple)):
            r = _like(a)
            if r is not None:
                return r
    return None


def _stack(seq):
    """np.array([[a, b], [c, d]]) for nested lists of device arrays/scalars."""
    ref = _like(seq)
    items = []
    for a in seq:
        if isinstance(a, (list, tuple)):
            items.append(_stack(a).t)
        elif isinstance(a, DeviceArray):
            items.append(a.t)
        else:
            items.append(torch.as_tensor(a, dtype=ref.dtype, device=ref.device))
    shape = torch.broadcast_shapes(*[i.shape for i in items])
    return DeviceArray(torch.stack([i.expand(shape) for i in items]))


def combine_layout(lays):
    """Memory order numpy would give the result of an elementwise op.

    numpy allocates ufunc outputs in 'K' order: C order wins any conflict, an
    operand whose element axis has stride 0 (a broadcast basis value, 'B') has
    no say, Fortran order ('F', e.g. the affine ``dx``) survives only if no
    C-ordered operand takes part (nditer's axis-ordering rule).  The order
    matters because ``np.sum(axis=1)`` is pairwise on C-ordered data and plain
    left-to-right on F-ordered data ("layout rule")."""
    lays = [x for x in lays if x is not None]
    if "C" in lays:
        return "C"
    if "F" in lays:
        return "F"
    return "C"


def layout_of(x):
    return x.lay if isinstance(x, DeviceArray) else None


class DeviceArray:
    """A float64 CUDA tensor speaking enough of the ndarray protocol for form
    definitions.  ``lay`` tracks the memory order ('C', 'F' or 'B'roadcast) the
    equivalent numpy array would have; the device data itself is always
    C-ordered."""
    __array_priority__ = 1000.0
    __slots__ = ("t", "lay")

    def __init__(self, t, lay="C"):
        if isinstance(t, DeviceArray):
            t, lay = t.t, t.lay
        self.t = t
        self.lay = lay

    # -- ndarray-like attributes -------------------------------------------------
    @property
    def shape(self):
        return tuple(self.t.shape)

    @property
    def ndim(self):
        return self.t.dim()

    @property
    def dtype(self):
        return np.dtype(str(self.t.dtype).replace("torch.", ""))

    @property
    def size(self):
        return self.t.numel()

    def __len__(self):
        return self.t.shape[0]

    def __iter__(self):
        for k in range(self.t.shape[0]):
            yield DeviceArray(self.t[k])

    def __getitem__(self, key):
        if isinstance(key, DeviceArray):
            key = key.t
        elif isinstance(key, tuple):
            key = tuple(k.t if isinstance(k, DeviceArray) else k for k in key)
        return DeviceArray(self.t[key])

    def numpy(self):
        return self.t.detach().cpu().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype)

    def __repr__(self):
        return "DeviceArray(shape={}, device={})".format(self.shape, self.t.device)

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return DeviceArray(self.t.reshape(shape))

    def sum(self, axis=None, **kw):
        return DeviceArray(self.t.sum() if axis is None else self.t.sum(dim=axis))

    def copy(self):
        return DeviceArray(self.t.clone())

    def astype(self, dtype):
        return DeviceArray(self.t.to(getattr(torch, np.dtype(dtype).name)))

    # -- operators -----------------------------------------------------------------
    def _bin(self, other, fn, swap=False):
        o = raw(other)
        if o is NotImplemented:
            return NotImplemented
        return DeviceArray(fn(o, self.t) if swap else fn(self.t, o),
                           combine_layout([self.lay, layout_of(other)]))

    def __add__(self, o): return self._bin(o, torch.add)
    def __radd__(self, o): return self._bin(o, torch.add, True)
    def __sub__(self, o): return self._bin(o, torch.sub)
    def __rsub__(self, o): return self._bin(o, lambda a, b: torch.sub(torch.as_tensor(a, dtype=b.dtype, device=b.device), b), True)
    def __mul__(self, o): return self._bin(o, torch.mul)
    def __rmul__(self, o): return self._bin(o, torch.mul, True)
    def __truediv__(self, o): return self._bin(o, torch.div)
    def __rtruediv__(self, o): return self._bin(o, lambda a, b: torch.div(torch.as_tensor(a, dtype=b.dtype, device=b.device), b), True)
    def __pow__(self, o): return self._bin(o, _power)
    def __rpow__(self, o): return self._bin(o, lambda a, b: torch.pow(torch.as_tensor(a, dtype=b.dtype, device=b.device), b), True)
    def __neg__(self): return DeviceArray(torch.neg(self.t), combine_layout([self.lay]))
    def __pos__(self): return DeviceArray(self.t, self.lay)
    def __abs__(self): return DeviceArray(torch.abs(self.t), combine_layout([self.lay]))
    def __lt__(self, o): return self._bin(o, torch.lt)
    def __le__(self, o): return self._bin(o, torch.le)
    def __gt__(self, o): return self._bin(o, torch.gt)
    def __ge__(self, o): return self._bin(o, torch.ge)

    # -- numpy protocols --------------------------------------------------------------
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__" or kwargs.get("out") is not None:
            return NotImplemented
        name = ufunc.__name__
        args = [raw(a) for a in inputs]
        ref = _like(inputs)
        lay = combine_layout([layout_of(a) for a in inputs])
        if name in _UNARY and len(args) == 1:
            return DeviceArray(_UNARY[name](args[0]), lay)
        if name in _BINARY and len(args) == 2:
            fn = _BINARY[name]
            if name in ("power", "float_power"):
                if not torch.is_tensor(args[0]):
                    args[0] = torch.as_tensor(args[0], dtype=ref.dtype, device=ref.device)
                return DeviceArray(fn(args[0], args[1]), lay)
            args = [a if torch.is_tensor(a) else torch.as_tensor(a, dtype=ref.dtype, device=ref.device)
                    for a in args]
            return DeviceArray(fn(*args), lay)
        raise NotImplementedError("numpy ufunc '{}' is not available on device fields".format(name))

    def __array_function__(self, func, types, args, kwargs):
        impl = _FUNCTIONS.get(func)
        if impl is None:
            raise NotImplementedError("numpy function '{}' is not available on device "
                                      "fields".format(func.__name__))
        return impl(*args, **kwargs)


class DiscreteField(DeviceArray):
    """Value plus derivative attributes, all device arrays
    (skfem/element/discrete_field.py:7-63).  Arithmetic and indexing return
    plain ``DeviceArray`` objects, like the reference invalidates the
    attributes after ufuncs."""
    __slots__ = ("grad", "div", "curl", "hess", "grad3", "grad4", "grad5", "grad6")
    _extra_attrs = ("grad", "div", "curl", "hess", "grad3", "grad4", "grad5", "grad6")

    def __init__(self, value=None, grad=None, div=None, curl=None, hess=None,
                 grad3=None, grad4=None, grad5=None, grad6=None, lay="C"):
        super().__init__(value, lay)

        def wrap(a):
            return None if a is None else (a if isinstance(a, DeviceArray) else DeviceArray(a))
        self.grad, self.div, self.curl, self.hess = wrap(grad), wrap(div), wrap(curl), wrap(hess)
        self.grad3, self.grad4, self.grad5, self.grad6 = (wrap(grad3), wrap(grad4),
                                                          wrap(grad5), wrap(grad6))

    def get(self, n):
        if n == 0:
            return DeviceArray(self.t, self.lay)
        return getattr(self, self._extra_attrs[n - 1])

    @property
    def astuple(self):
        return tuple(self.get(i) for i in range(len(self._extra_attrs) + 1))

    @property
    def value(self):
        return DeviceArray(self.t, self.lay)

    def zeros(self):
        return DiscreteField(*tuple(None if c is None else torch.zeros_like(c.t)
                                    for c in self.astuple))

    def __repr__(self):
        return "<skfem_b200 DiscreteField (device) shape={}>".format(self.shape)


# ---------------------------------------------------------------------------
# einsum with numpy's accumulation order: output index tuple fixed, contracted
# indices visited in row-major order of their first appearance, acc += product
# ---------------------------------------------------------------------------
def _parse(spec, operands):
    spec = spec.replace(" ", "")
    if "->" in spec:
        lhs, out = spec.split("->")
    else:
        lhs, out = spec, None
    ins = lhs.split(",")
    if len(ins) != len(operands):
        raise ValueError("einsum: operand count mismatch")
    return ins, out


def einsum(spec, *operands):
    ins, out = _parse(spec, operands)
    ts = [raw(o) for o in operands]
    named, ell_nd = [], 0
    for sub, t in zip(ins, ts):
        base = sub.replace("...", "")
        if "..." in sub:
            ell_nd = max(ell_nd, t.dim() - len(base))
        named.append(base)
    # label sizes
    size = {}
    for sub, t, base in zip(ins, ts, named):
        lead = sub.index("...") if "..." in sub else len(base)
        for pos, lab in enumerate(base):
            ax = pos if pos < lead else t.dim() - (len(base) - pos)
            size[lab] = t.shape[ax]
    all_labels = []
    for base in named:
        for lab in base:
            if lab not in all_labels:
                all_labels.append(lab)
    if out is None:
        counts = {lab: sum(b.count(lab) for b in named) for lab in all_labels}
        out_named = "".join(sorted(lab for lab in all_labels if counts[lab] == 1))
        out_has_ell = ell_nd > 0
        out_lead = True  # implicit mode: ellipsis dims first
    else:
        out_named = out.replace("...", "")
        out_has_ell = "..." in out
        out_lead = out.startswith("...") if out_has_ell else True
    contracted = [lab for lab in all_labels if lab not in out_named]

    def pick(t, sub, base, assign):
        """Slice operand t at the given label assignment -> ellipsis-shaped."""
        lead = sub.index("...") if "..." in sub else len(base)
        idx = []
        for pos in range(lead):
            idx.append(assign[base[pos]])
        nell = t.dim() - len(base)
        idx += [slice(None)] * nell
        for pos in range(lead, len(base)):
            idx.append(assign[base[pos]])
        return t[tuple(idx)]

    def ranges(labels):
        if not labels:
            yield {}
            return
        first, rest = labels[0], labels[1:]
        for v in range(size[first]):
            for r in ranges(rest):
                d = {first: v}
                d.update(r)
                yield d

    pieces = []
    for oa in ranges(list(out_named)):
        acc = None
        for ca in ranges(contracted):
            assign = dict(oa)
            assign.update(ca)
            term = None
            for t, sub, base in zip(ts, ins, named):
                s = pick(t, sub, base, assign)
                term = s if term is None else term * s
            acc = term if acc is None else acc + term
        pieces.append(acc)
    if not out_named:
        return DeviceArray(pieces[0])
    shape = torch.broadcast_shapes(*[p.shape for p in pieces])
    res = torch.stack([p.expand(shape) for p in pieces])
    # named output labels first, broadcast ('...') dimensions last: the layout of
    # every helper ('ij...->ji...', 'i...,j...->ij...', 'ij...,j...->i...')
    if out is not None and out_has_ell and out_lead and out_named:
        raise NotImplementedError("einsum outputs of the form '...ij' are not supported "
                                  "on device fields")
    res = res.reshape(tuple(size[lab] for lab in out_named) + tuple(shape))
    return DeviceArray(res)


def _np_array(obj, *a, **k):
    if isinstance(obj, DeviceArray):
        return DeviceArray(obj.t)
    return _stack(obj)


def _np_sum(a, axis=None, **k):
    return a.sum(axis=axis)


def _np_where(c, a, b):
    ref = _like([c, a, b])
    a, b = (x if torch.is_tensor(x) else torch.as_tensor(x, dtype=torch.float64, device=ref.device)
            for x in (raw(a), raw(b)))
    return DeviceArray(torch.where(raw(c), a, b))


_FUNCTIONS = {
    np.einsum: einsum,
    np.array: _np_array,
    np.asarray: _np_array,
    np.stack: lambda seq, axis=0: DeviceArray(torch.stack([raw(s) for s in seq], dim=axis)),
    np.zeros_like: lambda a, **k: DeviceArray(torch.zeros_like(raw(a))),
    np.ones_like: lambda a, **k: DeviceArray(torch.ones_like(raw(a))),
    np.sum: _np_sum,
    np.where: _np_where,
    np.shape: lambda a: a.shape,
    np.ndim: lambda a: a.ndim,
    np.broadcast_to: lambda a, shape: DeviceArray(raw(a).expand(tuple(shape))),
    np.moveaxis: lambda a, s, d: DeviceArray(torch.movedim(raw(a), s, d)),
    np.transpose: lambda a, axes=None: DeviceArray(raw(a).permute(*axes) if axes is not None
                                                   else raw(a).permute(*reversed(range(a.ndim)))),
    np.clip: lambda a, lo, hi: DeviceArray(torch.clamp(raw(a), lo, hi)),
}


def asdevice(a, device="cuda"):
    """Upload a host array for use inside forms."""
    if isinstance(a, DeviceArray):
        return a
    return DeviceArray(torch.as_tensor(np.asarray(a, dtype=np.float64), device=device))

"""Helper functions for writing forms (the reference's ``skfem.helpers``,
skfem/helpers.py:13-218).

Every contraction goes through ``np.einsum``; for device fields numpy
dispatches to :func:`skfem_b200.field.einsum`, which accumulates sequentially
in row-major order like numpy does, so traced user forms round identically.
Host ``ndarray`` inputs are handled by numpy itself, i.e. the helpers also work
on data a user pulled back to the host.
"""
import numpy as np

from .field import DeviceArray, _stack


def _array(items):
    """``np.array([...])`` of the reference's helpers.  numpy does not dispatch ``np.array``
    through ``__array_function__``, so a list of device fields has to be stacked explicitly
    (user forms: use ``np.stack`` or ``skfem_b200.helpers.array``); host arrays go to numpy."""
    if any(isinstance(a, DeviceArray) for a in items):
        return _stack(items)
    return np.array(items)


array = _array


def _contract(spec, *ops):
    return np.einsum(spec, *ops)


def grad(u):
    return u.grad


def div(u):
    if getattr(u, "div", None) is not None:
        return u.div
    if u.grad is not None:
        if u.grad.ndim >= 4:
            return _contract('ii...', u.grad)
        return u.grad[0]
    raise NotImplementedError


def curl(u):
    if getattr(u, "curl", None) is not None:
        return u.curl
    g = u.grad
    if g is None:
        raise NotImplementedError
    if g.ndim == 3 and g.shape[0] == 2:
        return _array([g[1], -g[0]])
    if g.ndim == 4 and g.shape[0] == 2:
        return g[1, 0] - g[0, 1]
    if g.ndim == 4 and g.shape[0] == 3:
        return _array([g[2, 1] - g[1, 2], g[0, 2] - g[2, 0], g[1, 0] - g[0, 1]])
    raise NotImplementedError


def d(u):
    for name in ("grad", "div", "curl"):
        if getattr(u, name, None) is not None:
            return getattr(u, name)
    raise NotImplementedError


def sym_grad(u):
    return .5 * (u.grad + transpose(u.grad))


def dd(u):
    return u.hess


def ddd(u):
    return u.grad3


def dddd(u):
    return u.grad4


def dot(u, v):
    return _contract('i...,i...', u, v)


def ddot(u, v):
    return _contract('ij...,ij...', u, v)


def dddot(u, v):
    return _contract('ijk...,ijk...', u, v)


def inner(u, v):
    if isinstance(u, tuple) and isinstance(v, tuple):
        return sum(inner(a, b) for a, b in zip(u, v))
    if u.ndim == 2:
        return u * v
    if u.ndim == 3:
        return dot(u, v)
    if u.ndim == 4:
        return ddot(u, v)
    raise NotImplementedError


def prod(u, v, w=None):
    if w is None:
        return _contract('i...,j...->ij...', u, v)
    return _contract('i...,j...,k...->ijk...', u, v, w)


def mul(A, x):
    return _contract('ij...,j...->i...', A, x)


def trace(T):
    return _contract('ii...', T)


def transpose(T):
    return _contract('ij...->ji...', T)


def eye(w, n):
    rows = [[w if i == j else 0. * w for i in range(n)] for j in range(n)]
    if isinstance(w, DeviceArray):
        return _stack(rows)
    return np.array(rows)


def identity(w, N=None):
    if N is None:
        if w.ndim > 2:
            N = w.shape[-3]
        else:
            raise ValueError("Cannot deduce the size of the identity matrix. "
                             "Give an explicit keyword argument N.")
    return eye(np.ones_like(w[(0,) * (w.ndim - 2)]) if w.ndim > 2 else np.ones_like(w), N)


def det(A):
    if A.shape[0] == 3:
        return (A[0, 0] * (A[1, 1] * A[2, 2] - A[1, 2] * A[2, 1])
                - A[0, 1] * (A[1, 0] * A[2, 2] - A[1, 2] * A[2, 0])
                + A[0, 2] * (A[1, 0] * A[2, 1] - A[1, 1] * A[2, 0]))
    if A.shape[0] == 2:
        return A[0, 0] * A[1, 1] - A[1, 0] * A[0, 1]
    raise NotImplementedError


def inv(A):
    """Inverse over the two leading axes, cofactor / determinant entry by entry
    in the reference's operation order (helpers.py:179-207)."""
    detA = det(A)
    if A.shape[0] == 3:
        rows = [[(-A[1, 2] * A[2, 1] + A[1, 1] * A[2, 2]) / detA,
                 (A[0, 2] * A[2, 1] - A[0, 1] * A[2, 2]) / detA,
                 (-A[0, 2] * A[1, 1] + A[0, 1] * A[1, 2]) / detA],
                [(A[1, 2] * A[2, 0] - A[1, 0] * A[2, 2]) / detA,
                 (-A[0, 2] * A[2, 0] + A[0, 0] * A[2, 2]) / detA,
                 (A[0, 2] * A[1, 0] - A[0, 0] * A[1, 2]) / detA],
                [(-A[1, 1] * A[2, 0] + A[1, 0] * A[2, 1]) / detA,
                 (A[0, 1] * A[2, 0] - A[0, 0] * A[2, 1]) / detA,
                 (-A[0, 1] * A[1, 0] + A[0, 0] * A[1, 1]) / detA]]
    elif A.shape[0] == 2:
        rows = [[A[1, 1] / detA, -A[0, 1] / detA], [-A[1, 0] / detA, A[0, 0] / detA]]
    else:
        raise NotImplementedError
    return np.stack([np.stack(r) for r in rows])


def cross(A, B):
    if A.shape[0] == 2:
        return A[0] * B[1] - A[1] * B[0]
    if A.shape[0] == 3:
        return _array([A[1] * B[2] - A[2] * B[1],
                       A[2] * B[0] - A[0] * B[2],
                       A[0] * B[1] - A[1] * B[0]])
    raise NotImplementedError


def jump(w, *args):
    if not hasattr(w, 'idx'):
        return args
    out = [(-1.) ** w.idx[i] * a for i, a in enumerate(args)]
    return out[0] if len(out) == 1 else tuple(out)

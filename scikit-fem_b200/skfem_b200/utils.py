"""Boundary conditions and the solver hand-off on the device-resident system
(SURVEY 8f rank 2): ``enforce``, ``condense`` and ``solve`` with the signatures of
``skfem.utils`` (skfem/utils.py:282-400, 462-603, 606-660), operating on
:class:`~skfem_b200.form.DeviceCSR` matrices and torch CUDA vectors so that the
assembled system never visits the host.

    A = laplace.assemble_device(basis)
    b = unit_load.assemble_device(basis)
    x = solve(*condense(A, b, D=basis.get_dofs()))      # torch tensor, on the GPU

``condense`` / ``enforce`` reproduce scipy's results bit for bit (the values are
copied; ``b[I] - A[I][:, D] @ x[D]`` adds in scipy's ``csr_matvec`` order).
``solve`` is a Jacobi-preconditioned conjugate gradient (the reference defaults
to SuperLU, which has no place on the device); it is meant for the symmetric
positive definite systems ``condense`` produces.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .form import DeviceCSR, _stream


def _torch():
    import torch
    return torch


def _flatten_dofs(S, device):
    """None | ndarray | tensor | DofsView-like (``.all()``) | dict of those ->
    int32 device tensor (utils.py:270-279)."""
    torch = _torch()
    if S is None:
        return None
    if isinstance(S, dict):     # np.unique(np.concatenate(...)): named sets may share DOFs
        return torch.unique(torch.cat([_flatten_dofs(v, device) for v in S.values()]))
    if hasattr(S, "all") and not isinstance(S, np.ndarray) and not torch.is_tensor(S):
        S = S.all()
    if torch.is_tensor(S):
        return S.to(device=device, dtype=torch.int32).reshape(-1)
    return torch.as_tensor(np.asarray(S).reshape(-1).astype(np.int32), device=device)


def _vec(v, device):
    torch = _torch()
    if v is None or isinstance(v, DeviceCSR):
        return v
    if torch.is_tensor(v):
        return v.to(device=device, dtype=torch.float64)
    return torch.as_tensor(np.asarray(v, dtype=np.float64), device=device)


def _init_bc(A, b, x, I, D):
    """utils.py:282-323 on the device: complete (I, D), default x / b."""
    torch = _torch()
    if not isinstance(A, DeviceCSR):
        raise TypeError("skfem_b200.utils works on DeviceCSR matrices (form.assemble_device); "
                        "for scipy matrices use skfem.utils or DeviceCSR.from_scipy")
    dev = A.data.device
    n = A.shape[0]
    D, I = _flatten_dofs(D, dev), _flatten_dofs(I, dev)
    if I is None and D is None:
        raise Exception("Either I or D must be given!")
    if I is not None and D is not None:
        raise Exception("Give only I or only D!")
    given = D if I is None else I
    mask = torch.zeros(n, dtype=torch.bool, device=dev)
    mask[given.long()] = True
    other = torch.nonzero(~mask).flatten().to(torch.int32)     # np.setdiff1d: ascending
    if I is None:
        I = other
    else:
        D = other
    b, x = _vec(b, dev), _vec(x, dev)
    if x is None:
        x = torch.zeros(n, dtype=torch.float64, device=dev)
    elif b is None:
        b = torch.zeros_like(x)
    return b, x, I, D


def enforce(A, b=None, x=None, I=None, D=None, diag=1., overwrite=False):
    """Set the rows ``D`` of ``A`` to ``diag`` times identity rows and ``b[D] = x[D]``
    (utils.py:327-400).  Returns ``(A, b)`` (or ``A``), like the reference."""
    torch = _torch()
    b, x, I, D = _init_bc(A, b, x, I, D)
    Aout = A if overwrite else A.copy()
    missing = torch.zeros(1, dtype=torch.int32, device=A.data.device)
    D = D.contiguous()
    code = _lib.lib().skb_csr_enforce(Aout.indptr.data_ptr(), Aout.indices.data_ptr(),
                                      Aout.data.data_ptr(), D.data_ptr(), D.shape[0],
                                      C.c_double(float(diag)), missing.data_ptr(), _stream())
    _lib.check(code, "skb_csr_enforce")
    if int(missing.item()) and float(diag) != 0.:
        # (diag == 0: the row is zero either way; the reference's setdiag would additionally
        # store explicit zeros on the diagonal, utils.py:385-388)
        raise NotImplementedError("enforce: a row of D has no stored diagonal entry")
    if b is None:
        return Aout
    if isinstance(b, DeviceCSR):      # mass matrix of an eigenvalue / initial value problem
        return Aout, enforce(b, D=D, diag=0., overwrite=overwrite)
    bout = b if overwrite else b.clone()
    bout[D.long()] = x[D.long()]
    return Aout, bout


def _submatrix(A, I, colmap, x=None, b=None):
    """A[I][:, colmap >= 0] (+ the condensed right-hand side) with two kernels
    and one scan."""
    torch = _torch()
    lib = _lib.lib()
    dev = A.data.device
    nI = int(I.shape[0])
    counts = torch.empty(nI, dtype=torch.int32, device=dev)
    code = lib.skb_csr_condense_count(A.indptr.data_ptr(), A.indices.data_ptr(), I.data_ptr(), nI,
                                      colmap.data_ptr(), counts.data_ptr(), _stream())
    _lib.check(code, "skb_csr_condense_count")
    indptr = torch.zeros(nI + 1, dtype=torch.int32, device=dev)
    indptr[1:] = torch.cumsum(counts, 0)
    nnz = int(indptr[-1]) if nI else 0
    indices = torch.empty(nnz, dtype=torch.int32, device=dev)
    data = torch.empty(nnz, dtype=torch.float64, device=dev)
    bout = None if b is None else torch.empty(nI, dtype=torch.float64, device=dev)
    code = lib.skb_csr_condense_fill(
        A.indptr.data_ptr(), A.indices.data_ptr(), A.data.data_ptr(), I.data_ptr(), nI,
        colmap.data_ptr(), indptr.data_ptr(), indices.data_ptr(), data.data_ptr(),
        None if b is None else x.data_ptr(), None if b is None else b.data_ptr(),
        None if b is None else bout.data_ptr(), _stream())
    _lib.check(code, "skb_csr_condense_fill")
    ncols = int((colmap >= 0).sum())
    return DeviceCSR(indptr, indices, data, (nI, ncols)), bout


def condense(A, b=None, x=None, I=None, D=None, expand=True):
    """Eliminate the degrees of freedom ``D`` (utils.py:462-603):
    ``A[I][:, I]``, ``b[I] - A[I][:, D] @ x[D]`` and, with ``expand``, ``x`` and
    ``I`` so that :func:`solve` can expand the solution."""
    torch = _torch()
    b, x, I, D = _init_bc(A, b, x, I, D)
    dev = A.data.device
    I = I.contiguous()
    if I.shape[0] > 1 and not bool((I[1:] > I[:-1]).all()):
        raise NotImplementedError("condense: I must be ascending and unique")
    colmap = torch.full((A.shape[1],), -1, dtype=torch.int32, device=dev)
    colmap[I.long()] = torch.arange(I.shape[0], dtype=torch.int32, device=dev)
    if b is None:
        ret = (_submatrix(A, I, colmap)[0],)
    elif isinstance(b, DeviceCSR):    # generalized eigenvalue problem: rhs matrix untouched
        ret = (_submatrix(A, I, colmap)[0], _submatrix(b, I, colmap)[0])
    else:
        ret = _submatrix(A, I, colmap, x=x.contiguous(), b=b.contiguous())
    if expand:
        ret += (x, I)
    return ret if len(ret) > 1 else ret[0]


def matvec(A, x):
    """``A @ x`` with the row sums in scipy's order (``skb_csr_spmv``)."""
    torch = _torch()
    y = torch.empty(A.shape[0], dtype=torch.float64, device=A.data.device)
    code = _lib.lib().skb_csr_spmv(A.indptr.data_ptr(), A.indices.data_ptr(), A.data.data_ptr(),
                                   x.contiguous().data_ptr(), y.data_ptr(), A.shape[0], _stream())
    _lib.check(code, "skb_csr_spmv")
    return y


def solver_iter_pcg(tol=1e-11, maxiter=None, verbose=False):
    """Jacobi-preconditioned CG on the device (the counterpart of
    ``solver_iter_pcg``/``solver_iter_krylov``, utils.py:151-226)."""
    def solver(A, b):
        torch = _torch()
        n = A.shape[0]
        diag = torch.zeros(n, dtype=torch.float64, device=b.device)
        rows = torch.repeat_interleave(torch.arange(n, device=b.device),
                                       (A.indptr[1:] - A.indptr[:-1]).long())
        on_diag = rows == A.indices.long()
        diag[rows[on_diag]] = A.data[on_diag]
        minv = torch.where(diag != 0, 1. / diag, torch.ones_like(diag))
        xk = torch.zeros_like(b)
        r = b.clone()
        z = minv * r
        p = z.clone()
        rz = torch.dot(r, z)
        bnorm = float(torch.linalg.norm(b))
        if bnorm == 0.0:
            return xk
        for it in range(maxiter if maxiter is not None else 10 * n):
            Ap = matvec(A, p)
            alpha = rz / torch.dot(p, Ap)
            xk += alpha * p
            r -= alpha * Ap
            if it % 8 == 7 and float(torch.linalg.norm(r)) <= tol * bnorm:
                break
            z = minv * r
            rz_new = torch.dot(r, z)
            p = z + (rz_new / rz) * p
            rz = rz_new
        if verbose:
            print("pcg: {} iterations, |r|/|b| = {:.2e}".format(
                it + 1, float(torch.linalg.norm(r)) / bnorm))
        return xk
    return solver


def solve(A, b, x=None, I=None, solver=None, **kwargs):
    """Solve ``A y = b`` on the device and, if ``x`` and ``I`` are given (as
    returned by :func:`condense`), expand ``y`` into ``x[I]`` (utils.py:606-660)."""
    if solver is None:
        solver = solver_iter_pcg(**kwargs)
    y = solver(A, _vec(b, A.data.device))
    if x is not None and I is not None:
        out = x.clone()
        out[I.long()] = y
        return out
    return y

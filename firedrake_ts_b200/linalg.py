"""Krylov building blocks for the standalone harness (ts.py): CSR mat-vec and
diagonal through the library's kernels on CUDA tensors; PCG and restarted GMRES
written on torch tensors so that they run on either device.  CPU tensors (only
met in the oracle-driven tests) go through scipy."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def csr_spmv(rowptr, colidx, values, x, kernel="stream"):
    n = rowptr.numel() - 1
    if values.is_cuda:
        from .engine import _check, load_library

        y = torch.empty(n, dtype=torch.float64, device=values.device)
        lib = load_library()
        fn = lib.fdts_csr_spmv if kernel == "stream" else lib.fdts_csr_spmv_warp
        with torch.cuda.device(values.device):
            _check(fn(n, rowptr.data_ptr(), colidx.data_ptr(), values.data_ptr(), x.data_ptr(), y.data_ptr(),
                      _stream()), "fdts_csr_spmv")
        return y
    import scipy.sparse as sp

    A = sp.csr_matrix((values.numpy(), colidx.numpy(), rowptr.numpy()), shape=(n, x.numel()))
    return torch.from_numpy(A @ x.numpy())


def csr_diagonal(rowptr, colidx, values, nown_nodes=None, nnodes=None, ncomp=1, layout=0):
    n = rowptr.numel() - 1
    nown_nodes = n // ncomp if nown_nodes is None else nown_nodes
    nnodes = nown_nodes if nnodes is None else nnodes
    if values.is_cuda:
        from .engine import _check, load_library

        d = torch.empty(n, dtype=torch.float64, device=values.device)
        with torch.cuda.device(values.device):
            _check(load_library().fdts_csr_diagonal(n, rowptr.data_ptr(), colidx.data_ptr(), values.data_ptr(),
                                                    nown_nodes, nnodes, ncomp, layout, d.data_ptr(), _stream()),
                   "fdts_csr_diagonal")
        return d
    rp, ci, v = rowptr.numpy(), colidx.numpy(), values.numpy()
    d = np.zeros(n)
    for r in range(n):
        col = r if layout == 0 else (r // nown_nodes) * nnodes + (r % nown_nodes)
        seg = ci[rp[r]:rp[r + 1]]
        k = np.searchsorted(seg, col)
        if k < len(seg) and seg[k] == col:
            d[r] = v[rp[r] + k]
    return torch.from_numpy(d)


def pcg(matvec, b, prec, dot, rtol, atol, max_it):
    x = torch.zeros_like(b)
    r = b.clone()
    z = prec(r)
    p = z.clone()
    rz = dot(r, z)
    bnorm = dot(b, b) ** 0.5
    tol = max(rtol * bnorm, atol)
    it = 0
    while it < max_it and dot(r, r) ** 0.5 > tol:
        Ap = matvec(p)
        alpha = rz / dot(p, Ap)
        x.add_(p, alpha=alpha)
        r.add_(Ap, alpha=-alpha)
        z = prec(r)
        rz_new = dot(r, z)
        p = z + (rz_new / rz) * p
        rz = rz_new
        it += 1
    return x, it


def pcg_device(matvec, b, prec, rtol, atol, max_it, check_every=4, allreduce=None):
    """PCG with every scalar kept on the device (0-d tensors): no host synchronisation per
    iteration; the norm is read back every ``check_every`` iterations only.  ``allreduce`` sums a
    0-d tensor over the ranks (None on one rank).  Same recurrence as :func:`pcg`."""
    red = allreduce or (lambda t: t)
    dot = lambda a, c: red(torch.dot(a, c))  # noqa: E731
    x = torch.zeros_like(b)
    r = b.clone()
    z = prec(r)
    p = z.clone()
    rz = dot(r, z)
    tol = max(rtol * float(dot(b, b)) ** 0.5, atol)
    it = 0
    while it < max_it:
        if it % check_every == 0 and float(dot(r, r)) ** 0.5 <= tol:
            break
        Ap = matvec(p)
        pAp = dot(p, Ap)
        # an exactly converged iterate between two checks (r = 0, p = 0) must not produce 0/0
        alpha = torch.where(pAp != 0, rz / pAp, torch.zeros_like(rz))
        x.addcmul_(p, alpha)
        r.addcmul_(Ap, -alpha)
        z = prec(r)
        rz_new = dot(r, z)
        p = torch.addcmul(z, p, torch.where(rz != 0, rz_new / rz, torch.zeros_like(rz)))
        rz = rz_new
        it += 1
    return x, it


def gmres(matvec, b, prec, dot, rtol, atol, max_it, restart=30):
    """left-preconditioned restarted GMRES (modified Gram-Schmidt, Givens rotations)."""
    x = torch.zeros_like(b)
    bnorm = dot(prec(b), prec(b)) ** 0.5
    tol = max(rtol * bnorm, atol)
    its = 0
    while its < max_it:
        r = prec(b - matvec(x)) if its else prec(b)
        beta = dot(r, r) ** 0.5
        if beta <= tol:
            break
        m = min(restart, max_it - its)
        V = [r / beta]
        H = np.zeros((m + 1, m))
        cs, sn = np.zeros(m), np.zeros(m)
        g = np.zeros(m + 1)
        g[0] = beta
        k_used = 0
        for k in range(m):
            w = prec(matvec(V[k]))
            for i in range(k + 1):
                H[i, k] = dot(w, V[i])
                w = w - H[i, k] * V[i]
            H[k + 1, k] = dot(w, w) ** 0.5
            for i in range(k):
                t = cs[i] * H[i, k] + sn[i] * H[i + 1, k]
                H[i + 1, k] = -sn[i] * H[i, k] + cs[i] * H[i + 1, k]
                H[i, k] = t
            den = (H[k, k] ** 2 + H[k + 1, k] ** 2) ** 0.5
            cs[k], sn[k] = (H[k, k] / den, H[k + 1, k] / den) if den > 0 else (1.0, 0.0)
            H[k, k] = cs[k] * H[k, k] + sn[k] * H[k + 1, k]
            H[k + 1, k] = 0.0
            g[k + 1] = -sn[k] * g[k]
            g[k] = cs[k] * g[k]
            its += 1
            k_used = k + 1
            if abs(g[k + 1]) <= tol or den == 0:
                break
            V.append(w / den if False else w / max(dot(w, w) ** 0.5, 1e-300))
        y = np.linalg.solve(np.triu(H[:k_used, :k_used]), g[:k_used]) if k_used else np.zeros(0)
        for i in range(k_used):
            x = x + float(y[i]) * V[i]
        if k_used and abs(g[k_used]) <= tol:
            break
    return x, its

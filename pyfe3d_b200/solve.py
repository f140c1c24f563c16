"""Device-side consumers of the assembled CSR matrix (first step): masked SpMV for the
boundary-condition partition and a Jacobi-preconditioned conjugate gradient, so that a static solve never
leaves the GPU.  What it replaces in a reference script (tests/test_quad4_static_point_load.py:84-104)::

    KC0uu = KC0[bu, :][:, bu]                 ->  u = cg_solve(indptr, indices, vals, fext, free=bu)
    uu, info = cg(KC0uu, fext[bu], atol=1e-9)

The SpMV and the diagonal extraction are C-ABI kernels; vector updates and dot products use torch tensors
(device plumbing)."""
import torch

from .batch import _dev, _ptr, context


def masked_spmv(indptr, indices, vals, free, x, out=None):
    """y = P A P x with P = diag(free)."""
    n = indptr.numel() - 1
    if out is None:
        out = torch.empty(n, dtype=torch.float64, device=vals.device)
    context(vals.device).spmv_csr_masked(n, _ptr(indptr), _ptr(indices), _ptr(vals), _ptr(free), _ptr(x), _ptr(out))
    return out


def diagonal(indptr, indices, vals, row0=0):
    n = indptr.numel() - 1
    d = torch.empty(n, dtype=torch.float64, device=vals.device)
    context(vals.device).csr_diagonal(n, _ptr(indptr), _ptr(indices), _ptr(vals), row0, _ptr(d))
    return d


def cg_solve(indptr, indices, vals, b, free=None, rtol=1e-12, maxiter=None, x0=None):
    """Solve (P A P) x = P b for symmetric positive definite A[free, free] with Jacobi-preconditioned CG.
    Returns (x, info): x is zero on constrained DOFs; info = iterations used (negative: not converged)."""
    dev = vals.device
    n = indptr.numel() - 1
    if free is None:
        free = torch.ones(n, dtype=torch.uint8, device=dev)
    free = _dev(free, torch.uint8, dev)
    b = _dev(b, torch.float64, dev) * free
    d = diagonal(indptr, indices, vals)
    minv = torch.where((free > 0) & (d != 0), 1.0 / d, torch.zeros_like(d))
    x = torch.zeros(n, dtype=torch.float64, device=dev) if x0 is None else _dev(x0, torch.float64, dev) * free
    r = b - masked_spmv(indptr, indices, vals, free, x)
    z = minv * r
    p = z.clone()
    rz = torch.dot(r, z)
    bnorm = torch.linalg.vector_norm(b)
    if float(bnorm) == 0.0:
        return x, 0
    maxiter = maxiter or 10 * n
    ap = torch.empty_like(x)
    for it in range(1, maxiter + 1):
        masked_spmv(indptr, indices, vals, free, p, out=ap)
        alpha = rz / torch.dot(p, ap)
        x += alpha * p
        r -= alpha * ap
        if it % 8 == 0 and float(torch.linalg.vector_norm(r)) <= rtol * float(bnorm):
            return x, it
        z = minv * r
        rz_new = torch.dot(r, z)
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, -maxiter

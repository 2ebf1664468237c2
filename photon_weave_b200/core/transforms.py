"""
Transform rules of the seam operations: reverse-mode derivatives and a
batching rule, so that gradient-based use of the reference (``jax.grad`` through
``measure_*_expectation``: tests/state/utils/test_measurements.py:337-371,
tests/state/test_envelope.py:846-870; ``jax.vmap`` over states:
tests/state/test_shape_planning.py:104-121,149-153) runs on the CUDA path.

The reference gets these rules for free from JAX's einsum; a custom call has to state them.
Every rule below is itself a libpwb200 kernel call -- the derivative of a linear apply is
another apply:

    y = (O on T) psi            psi_bar = (O^H on T) y_bar          O_bar[a,b] = sum_rest y_bar[a,.] conj(psi[b,.])
    rho' = sum_k K rho K^H      rho_bar = sum_k K^H rho'_bar K      K_bar = Tr_rest(G K rho^H + G^H K rho)
    p_j = s_j / S, s = |psi|^2  psi_bar = (diag(2 h) on T) psi,  h = (p_bar - <p_bar, p>) / S
    R = Tr_rest psi psi^H       psi_bar = ((G + G^H) on keep) psi
    R = Tr_rest rho             rho_bar = G on keep (x) identity on rest

(cotangents in PyTorch's convention, grad = dL/dRe + i dL/dIm; the JAX glue in
``core/jax_ffi.py`` conjugates where JAX's convention differs).  They are exposed as
``torch.autograd.Function``s -- PyTorch is this package's array plumbing -- and the adapters
route through them whenever an input requires a gradient.

Batching (``vmap``): a leading batch axis of the state is just one more untouched subsystem,
so a batched apply / Kraus is ONE kernel launch over dims (B,) + dims; operations without such
a rule (sampling) fall back to a sequential loop (JAX's ``vmap_method="sequential"``).
"""

from __future__ import annotations

import math
import threading
from typing import Any, Callable, Optional, Sequence

import torch

from . import ffi

_tls = threading.local()


def current_batch() -> Optional[int]:
    """Batch size of the enclosing ``vmap`` (None outside)."""
    return getattr(_tls, "batch", None)


class _Batch:
    def __init__(self, b: Optional[int]):
        self.b, self.prev = b, None

    def __enter__(self):
        self.prev = current_batch()
        _tls.batch = self.b
        return self

    def __exit__(self, *exc):
        _tls.batch = self.prev


def _h(op: torch.Tensor) -> torch.Tensor:
    """Conjugate transpose of a small operator (last two axes), materialised."""
    return op.conj().transpose(-1, -2).resolve_conj().contiguous()


def needs_grad(*xs: Any) -> bool:
    for x in xs:   # plain loop: this sits on the per-call path of the launch-bound configs
        if isinstance(x, torch.Tensor) and x.requires_grad:
            return torch.is_grad_enabled()
    return False


# --------------------------------------------------------------------------------------
# linear maps
# --------------------------------------------------------------------------------------


class ApplyVec(torch.autograd.Function):
    """psi' = (O on targets) psi -- core/kernels.py:62-98."""

    @staticmethod
    def forward(ctx, psi, op, dims, targets):
        ctx.save_for_backward(psi, op)
        ctx.dims, ctx.targets = tuple(dims), tuple(targets)
        return ffi.apply_vec(psi.contiguous().reshape(-1), op, ctx.dims, ctx.targets).reshape(psi.shape)

    @staticmethod
    def backward(ctx, g):
        psi, op = ctx.saved_tensors
        g = g.contiguous().reshape(-1)
        gpsi = gop = None
        if ctx.needs_input_grad[0]:
            gpsi = ffi.apply_vec(g, _h(op), ctx.dims, ctx.targets).reshape(psi.shape)
        if ctx.needs_input_grad[1]:
            gop = ffi.cross_reduced(g, psi.contiguous().reshape(-1), ctx.dims, ctx.targets)
        return gpsi, gop, None, None


class KrausMat(torch.autograd.Function):
    """rho' = sum_k K_k rho K_k^H -- core/kernels.py:101-186 (K = 1: U rho U^H)."""

    @staticmethod
    def forward(ctx, rho, ops, dims, targets):
        ctx.save_for_backward(rho, ops)
        ctx.dims, ctx.targets = tuple(dims), tuple(targets)
        return ffi.kraus_mat(rho.contiguous().reshape(-1), ops, ctx.dims, ctx.targets).reshape(rho.shape)

    @staticmethod
    def backward(ctx, g):
        rho, ops = ctx.saved_tensors
        dims, targets = ctx.dims, ctx.targets
        n = len(dims)
        g = g.contiguous().reshape(-1)
        grho = gops = None
        if ctx.needs_input_grad[0]:
            grho = ffi.kraus_mat(g, _h(ops), dims, targets).reshape(rho.shape)
        if ctx.needs_input_grad[1]:
            rows, cols = list(targets), [n + t for t in targets]
            flat = rho.contiguous().reshape(-1)
            both = dims + dims
            parts = []
            for k in range(ops.shape[0]):
                kop = ops[k].contiguous()
                z = ffi.apply_vec(flat, kop.conj().resolve_conj().contiguous(), both, cols)   # rho K^H
                w = ffi.apply_vec(flat, kop, both, rows)                                      # K rho
                m1 = ffi.cross_reduced(g, z, both, rows)                # Tr_rest(G K rho^H)
                m2 = ffi.cross_reduced(g, w, both, cols).conj().resolve_conj()   # Tr_rest(G^H K rho)
                parts.append(m1 + m2)
            gops = torch.stack(parts).reshape(ops.shape)
        return grho, gops, None, None


# --------------------------------------------------------------------------------------
# reductions
# --------------------------------------------------------------------------------------


def _scale_by_target(x_flat, h_real, dims, targets):
    """x[(j, rest)] * h[j]: a diagonal operator on the targets (the streaming diagonal kernel)."""
    d = torch.diag(h_real.to(x_flat.dtype)).contiguous()
    return ffi.apply_vec(x_flat, d, dims, targets)


class ProbsVec(torch.autograd.Function):
    """p_j = sum_rest |psi_j|^2 / sum |psi|^2 -- core/kernels.py:28-41."""

    @staticmethod
    def forward(ctx, psi, dims, targets):
        ctx.dims, ctx.targets = tuple(dims), tuple(targets)
        flat = psi.contiguous().reshape(-1)
        p = ffi.probs(flat, ctx.dims, ctx.targets, False)
        total = ffi.norm_stats(flat)[:1]
        ctx.save_for_backward(psi, p, total)
        return p

    @staticmethod
    def backward(ctx, g):
        psi, p, total = ctx.saved_tensors
        h = (g - (g * p).sum()) / total.to(g.dtype)
        out = _scale_by_target(psi.contiguous().reshape(-1), 2.0 * h, ctx.dims, ctx.targets)
        return out.reshape(psi.shape), None, None


class ProbsMat(torch.autograd.Function):
    """p_j = Re sum_rest rho[(j,r),(j,r)] / Tr rho -- core/kernels.py:44-59."""

    @staticmethod
    def forward(ctx, rho, dims, targets):
        ctx.dims, ctx.targets = tuple(dims), tuple(targets)
        n_tot = math.prod(dims)
        flat = rho.contiguous().reshape(-1)
        p = ffi.probs(flat, ctx.dims, ctx.targets, True)
        total = torch.diagonal(rho.reshape(n_tot, n_tot)).real.sum().reshape(1)
        ctx.save_for_backward(p, total)
        ctx.shape, ctx.cdtype = rho.shape, rho.dtype
        return p

    @staticmethod
    def backward(ctx, g):
        p, total = ctx.saved_tensors
        n_tot = math.prod(ctx.dims)
        h = (g - (g * p).sum()) / total.to(g.dtype)
        ones = torch.ones(n_tot, dtype=ctx.cdtype, device=g.device)
        diag = _scale_by_target(ones, h, ctx.dims, ctx.targets)       # h[j(index)] along the diagonal
        out = torch.zeros((n_tot, n_tot), dtype=ctx.cdtype, device=g.device)
        torch.diagonal(out).copy_(diag)
        return out.reshape(ctx.shape), None, None


class ReducedFromVec(torch.autograd.Function):
    """R = Tr_rest |psi><psi| -- state/utils/trace_out.py:17-25, core/kernels.py:266-283."""

    @staticmethod
    def forward(ctx, psi, dims, keep):
        ctx.save_for_backward(psi)
        ctx.dims, ctx.keep = tuple(dims), tuple(keep)
        return ffi.reduced_from_vec(psi.contiguous().reshape(-1), ctx.dims, ctx.keep)

    @staticmethod
    def backward(ctx, g):
        (psi,) = ctx.saved_tensors
        sym = (g + _h(g)).contiguous()
        out = ffi.apply_vec(psi.contiguous().reshape(-1), sym, ctx.dims, ctx.keep)
        return out.reshape(psi.shape), None, None


class TraceOutMat(torch.autograd.Function):
    """R = Tr_rest rho -- core/kernels.py:301-319."""

    @staticmethod
    def forward(ctx, rho, dims, keep):
        ctx.dims, ctx.keep = tuple(dims), tuple(keep)
        ctx.shape = rho.shape
        return ffi.trace_out_mat(rho.contiguous().reshape(-1), ctx.dims, ctx.keep)

    @staticmethod
    def backward(ctx, g):
        dims, keep = ctx.dims, list(ctx.keep)
        n = len(dims)
        rest = [i for i in range(n) if i not in keep]
        r = math.prod(dims[i] for i in rest)
        n_tot = math.prod(dims)
        eye = torch.eye(r, dtype=g.dtype, device=g.device)
        m = ffi.kron(g.contiguous(), eye)              # axes (keep..., rest...) on both sides
        cur = keep + rest
        if cur != list(range(n)):
            cur_dims = [dims[i] for i in cur]
            pos = [cur.index(i) for i in range(n)]
            m = ffi.permute(m.reshape(-1), cur_dims + cur_dims, pos + [n + p for p in pos])
        return m.reshape(n_tot, n_tot).reshape(ctx.shape), None, None


# --------------------------------------------------------------------------------------
# batching
# --------------------------------------------------------------------------------------


def vmap(fn: Callable, in_axes: Sequence[Optional[int]] = (0,), method: str = "auto") -> Callable:
    """``jax.vmap`` for functions built from the seam operations.

    ``in_axes`` gives, per positional argument, the batch axis (only 0 is supported) or None.
    method "auto": the function is called ONCE under a batch context; the apply / Kraus adapters
    then treat the leading axis as an untouched subsystem (one launch for the whole batch).  A
    function the batch context cannot serve (it raises ``NotBatchable``) and method
    "sequential" run one call per batch element and stack the results."""

    def wrapped(*args):
        sizes = {a.shape[0] for a, ax in zip(args, in_axes) if ax is not None}
        if any(ax not in (None, 0) for ax in in_axes) or len(sizes) != 1:
            raise ValueError("vmap: batch axes must be axis 0 with one common size")
        b = sizes.pop()
        if method != "sequential":
            try:
                with _Batch(b):
                    return fn(*args)
            except NotBatchable:
                pass
        outs = [fn(*[a[i] if ax is not None else a for a, ax in zip(args, in_axes)]) for i in range(b)]
        def stack(items):
            return torch.stack(list(items)) if all(isinstance(i, torch.Tensor) for i in items) else list(items)

        if isinstance(outs[0], tuple):
            return tuple(stack([o[k] for o in outs]) for k in range(len(outs[0])))
        return stack(outs)

    return wrapped


class NotBatchable(RuntimeError):
    """Raised inside a batch context by an operation that has no batching rule."""

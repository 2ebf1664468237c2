"""
JAX front end of the C ABI: the seam operations of ``photon_weave.core.adapters`` registered as
``jax.ffi`` custom calls (zero-copy device buffers, XLA's stream), with the transform rules the
reference's tests rely on -- ``custom_vjp`` (the derivative of an apply is another apply,
SURVEY.md section 8f.4) and ``vmap_method="sequential"`` / a batching rule for the applies.

GATED: importing this module needs ``jax`` and ``photon_weave_b200/lib/libpwb200_xla.so``
(``make -C photon_weave_b200/csrc xla`` on a machine whose jaxlib ships the XLA FFI headers).
Neither exists in the build image, so nothing in the package imports this module; the same
rules are exercised through PyTorch autograd in ``core/transforms.py`` /
``tests/test_gpu_transforms.py``.  Cotangents follow JAX's convention, which is the complex
conjugate of PyTorch's (``jax.grad(|z|^2) = 2 conj z``): rule_jax(ct) = conj(rule_torch(conj ct)).

A maintainer's binding in the reference is one line per function:
``from photon_weave_b200.core.jax_ffi import apply_operation_vector`` in ``core/adapters.py``.
"""

from __future__ import annotations

import ctypes
import math
import os
from functools import partial

import numpy as np

import jax  # noqa: E402  (ImportError here is the gate)
import jax.numpy as jnp

_HERE = os.path.dirname(os.path.abspath(__file__))
XLA_LIB = os.path.join(_HERE, "..", "lib", "libpwb200_xla.so")

#: ffi target name -> exported handler symbol (csrc/pw_xla_ffi.cc)
TARGETS = {
    "pw_apply_vec": "PwApplyVec",
    "pw_kraus_mat": "PwKrausMat",
    "pw_trace_out_mat": "PwTraceOutMat",
    "pw_reduced_from_vec": "PwReducedFromVec",
    "pw_cross_reduced": "PwCrossReduced",
    "pw_probs": "PwProbs",
    "pw_draw": "PwDraw",
    "pw_collapse": "PwCollapse",
    "pw_povm_probs": "PwPovmProbs",
    "pw_povm_select": "PwPovmSelect",
}

_registered = False


def register() -> None:
    """Load libpwb200_xla.so and register every handler for platform CUDA (idempotent)."""
    global _registered
    if _registered:
        return
    if not os.path.exists(XLA_LIB):
        raise ImportError(f"{XLA_LIB} not built: run `make -C photon_weave_b200/csrc xla` (needs jaxlib headers)")
    lib = ctypes.CDLL(XLA_LIB)
    for name, sym in TARGETS.items():
        jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(lib, sym)), platform="CUDA")
    _registered = True


def _i64(v):
    return np.asarray(tuple(int(x) for x in v), dtype=np.int64)


def _real(dtype):
    return jnp.float32 if dtype == jnp.complex64 else jnp.float64


def _call(name, out, *args, vmap_method="sequential", **attrs):
    register()
    return jax.ffi.ffi_call(name, out, vmap_method=vmap_method)(*args, **attrs)


# ---- primitives with their reverse-mode rules ------------------------------------------------


@partial(jax.custom_vjp, nondiff_argnums=(2, 3))
def apply_vec(psi, op, dims, targets):
    """psi' = (op on targets) psi (core/kernels.py:62-98)."""
    return _call("pw_apply_vec", jax.ShapeDtypeStruct(psi.shape, psi.dtype), psi, op,
                 dims=_i64(dims), targets=_i64(targets))


def _apply_vec_fwd(psi, op, dims, targets):
    return apply_vec(psi, op, dims, targets), (psi, op)


def _apply_vec_bwd(dims, targets, res, ct):
    psi, op = res
    t = op.shape[0]
    psi_bar = apply_vec(ct, op.T, dims, targets)                                   # O^T ct
    op_bar = _call("pw_cross_reduced", jax.ShapeDtypeStruct((t, t), psi.dtype), ct.reshape(-1),
                   jnp.conj(psi).reshape(-1), axes=_i64(dims), keep=_i64(targets))  # sum ct[a] psi[b]
    return psi_bar, op_bar


apply_vec.defvjp(_apply_vec_fwd, _apply_vec_bwd)


@partial(jax.custom_vjp, nondiff_argnums=(2, 3))
def kraus_mat(rho, ops, dims, targets):
    """rho' = sum_k K rho K^dagger, ops (K, t, t) (core/kernels.py:101-186)."""
    return _call("pw_kraus_mat", jax.ShapeDtypeStruct(rho.shape, rho.dtype), rho, ops,
                 dims=_i64(dims), targets=_i64(targets))


def _kraus_fwd(rho, ops, dims, targets):
    return kraus_mat(rho, ops, dims, targets), (rho, ops)


def _kraus_bwd(dims, targets, res, ct):
    rho, ops = res
    n = len(dims)
    both = tuple(dims) + tuple(dims)
    rows, cols = tuple(targets), tuple(n + t for t in targets)
    t = ops.shape[-1]
    # JAX convention = conj of the PyTorch rules of core/transforms.py applied to conj(ct)
    g = jnp.conj(ct)
    rho_bar = jnp.conj(kraus_mat(g, jnp.conj(jnp.swapaxes(ops, -1, -2)), dims, targets))
    flat = rho.reshape(-1)
    parts = []
    for k in range(ops.shape[0]):
        z = apply_vec(flat, jnp.conj(ops[k]), both, cols)
        w = apply_vec(flat, ops[k], both, rows)
        out = jax.ShapeDtypeStruct((t, t), rho.dtype)
        m1 = _call("pw_cross_reduced", out, g.reshape(-1), z, axes=_i64(both), keep=_i64(rows))
        m2 = jnp.conj(_call("pw_cross_reduced", out, g.reshape(-1), w, axes=_i64(both), keep=_i64(cols)))
        parts.append(jnp.conj(m1 + m2))
    return rho_bar, jnp.stack(parts).reshape(ops.shape)


kraus_mat.defvjp(_kraus_fwd, _kraus_bwd)


@partial(jax.custom_vjp, nondiff_argnums=(1, 2, 3))
def probs(state, dims, targets, is_matrix):
    """Outcome probabilities (core/kernels.py:28-59)."""
    t = math.prod(dims[i] for i in targets)
    return _call("pw_probs", jax.ShapeDtypeStruct((t,), _real(state.dtype)), state, dims=_i64(dims),
                 targets=_i64(targets), is_matrix=np.int64(bool(is_matrix)))


def _probs_fwd(state, dims, targets, is_matrix):
    p = probs(state, dims, targets, is_matrix)
    return p, (state, p)


def _probs_bwd(dims, targets, is_matrix, res, ct):
    state, p = res
    n_tot = math.prod(dims)
    if is_matrix:
        total = jnp.real(jnp.trace(state.reshape(n_tot, n_tot)))
        h = (ct - jnp.sum(ct * p)) / total
        ones = jnp.ones((n_tot,), dtype=state.dtype)
        diag = apply_vec(ones, jnp.diag(h.astype(state.dtype)), dims, targets)
        return (jnp.diag(diag).reshape(state.shape),)          # real, so conj is a no-op
    total = jnp.sum(jnp.abs(state) ** 2)
    h = (ct - jnp.sum(ct * p)) / total
    out = apply_vec(state.reshape(-1), jnp.diag((2.0 * h).astype(state.dtype)), dims, targets)
    return (jnp.conj(out).reshape(state.shape),)


probs.defvjp(_probs_fwd, _probs_bwd)


@partial(jax.custom_vjp, nondiff_argnums=(1, 2))
def reduced_from_vec(psi, dims, keep):
    """Tr_rest |psi><psi| (state/utils/trace_out.py:17-25)."""
    t = math.prod(dims[i] for i in keep)
    return _call("pw_reduced_from_vec", jax.ShapeDtypeStruct((t, t), psi.dtype), psi, dims=_i64(dims),
                 keep=_i64(keep))


def _reduced_fwd(psi, dims, keep):
    return reduced_from_vec(psi, dims, keep), psi


def _reduced_bwd(dims, keep, psi, ct):
    g = jnp.conj(ct)
    sym = g + jnp.conj(g.T)
    return (jnp.conj(apply_vec(psi.reshape(-1), sym, dims, keep)).reshape(psi.shape),)


reduced_from_vec.defvjp(_reduced_fwd, _reduced_bwd)


@partial(jax.custom_vjp, nondiff_argnums=(1, 2))
def trace_out_mat(rho, dims, keep):
    """Tr_rest rho (core/kernels.py:301-319)."""
    t = math.prod(dims[i] for i in keep)
    return _call("pw_trace_out_mat", jax.ShapeDtypeStruct((t, t), rho.dtype), rho, dims=_i64(dims),
                 keep=_i64(keep))


def _trace_fwd(rho, dims, keep):
    return trace_out_mat(rho, dims, keep), rho.shape


def _trace_bwd(dims, keep, shape, ct):
    n = len(dims)
    keep = list(keep)
    rest = [i for i in range(n) if i not in keep]
    r = math.prod(dims[i] for i in rest)
    cur = keep + rest
    cur_dims = [dims[i] for i in cur]
    m = jnp.kron(ct, jnp.eye(r, dtype=ct.dtype)).reshape(cur_dims + cur_dims)   # tiny glue: an outer product
    pos = [cur.index(i) for i in range(n)]
    return (jnp.transpose(m, pos + [n + p for p in pos]).reshape(shape),)


trace_out_mat.defvjp(_trace_fwd, _trace_bwd)


# ---- the adapters' surface (photon_weave/core/adapters.py:299-733) -----------------------------


def apply_operation_vector(state_dims, target_indices, product_state, operator, use_contraction=False):
    n = math.prod(state_dims)
    op = operator.astype(product_state.dtype)
    return apply_vec(product_state.reshape(-1), op, tuple(state_dims), tuple(target_indices)).reshape(n, 1)


def apply_operation_matrix(state_dims, target_indices, product_state, operator, use_contraction=False):
    n = math.prod(state_dims)
    ops = operator.astype(product_state.dtype)[None]
    return kraus_mat(product_state.reshape(-1), ops, tuple(state_dims), tuple(target_indices)).reshape(n, n)


def apply_kraus_matrix(state_dims, target_indices, product_state, operators, use_contraction=False):
    n = math.prod(state_dims)
    t = math.prod(state_dims[i] for i in target_indices)
    ops = jnp.asarray(operators).astype(product_state.dtype).reshape(-1, t, t)
    return kraus_mat(product_state.reshape(-1), ops, tuple(state_dims), tuple(target_indices)).reshape(n, n)


def trace_out_matrix(state_dims, target_indices, product_state, use_contraction=False):
    keep = sorted(target_indices) if use_contraction else list(target_indices)
    if len(keep) == len(state_dims):
        return product_state
    return trace_out_mat(product_state.reshape(-1), tuple(state_dims), tuple(keep))


def measure_vector_expectation(state_dims, target_indices, product_state):
    rest = tuple(i for i in range(len(state_dims)) if i not in target_indices)
    flat = product_state.reshape(-1)
    return (probs(flat, tuple(state_dims), tuple(target_indices), False),
            reduced_from_vec(flat, tuple(state_dims), rest))


def measure_matrix_expectation(state_dims, target_indices, product_state):
    rest = tuple(i for i in range(len(state_dims)) if i not in target_indices)
    flat = product_state.reshape(-1)
    return (probs(flat, tuple(state_dims), tuple(target_indices), True),
            trace_out_mat(flat, tuple(state_dims), rest))


def _measure(state_dims, target_indices, product_state, key, is_matrix):
    dims, tg = tuple(state_dims), tuple(target_indices)
    flat = product_state.reshape(-1)
    p = jax.lax.stop_gradient(probs(flat, dims, tg, is_matrix))
    kd = jax.random.key_data(key) if jnp.issubdtype(key.dtype, jax.dtypes.prng_key) else key
    outcome = _call("pw_draw", jax.ShapeDtypeStruct((), jnp.int64), p, kd.astype(jnp.uint32))
    r = math.prod(dims) // math.prod(dims[i] for i in tg)
    shape = (r, r) if is_matrix else (r, 1)
    post = _call("pw_collapse", jax.ShapeDtypeStruct(shape, product_state.dtype), flat, outcome, p,
                 dims=_i64(dims), targets=_i64(tg), is_matrix=np.int64(bool(is_matrix)))
    return outcome, post, p


def measure_vector(state_dims, target_indices, product_state, key):
    o, post, _ = _measure(state_dims, target_indices, product_state, key, False)
    return o, post, key


def measure_matrix(state_dims, target_indices, product_state, key):
    o, post, _ = _measure(state_dims, target_indices, product_state, key, True)
    return o, post, key


def measure_vector_with_probs(state_dims, target_indices, product_state, key):
    o, post, p = _measure(state_dims, target_indices, product_state, key, False)
    return o, post, p, key


def measure_matrix_with_probs(state_dims, target_indices, product_state, key):
    o, post, p = _measure(state_dims, target_indices, product_state, key, True)
    return o, post, p, key

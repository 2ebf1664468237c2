#!/usr/bin/env python
"""
Generates tests/golden/seam_golden.npz by EXECUTING THE REFERENCE'S OWN SOURCE for the hot path
-- /root/reference/photon_weave/core/kernels.py and core/adapters.py (with core/meta.py and
extra/einsum_constructor.py), loaded in place, never copied -- on seeded inputs.

jax / jaxlib / opt_einsum are not installable in the build container, so the two third-party
array libraries those files import are stood in for by NumPy:
    jax.numpy        -> numpy                 jax.jit  -> identity          jax.vmap -> loop + stack
    opt_einsum.contract(expr, *ops) -> numpy.einsum(expr, *ops, optimize=True)
    jax.random.choice -> oracle/rng_threefry.choice (the restated threefry2x32 + cumsum +
                         searchsorted of jax 0.8.1, pinned by its own known-answer tests)
Everything the REFERENCE wrote -- validation, reshapes, transposes, the generated einsum strings,
the +1e-18 regularisation, renormalisation, restoring the subsystem order, the kept-axis order of
the two trace-out routes, Kraus stacking -- runs exactly as written.  The fixture therefore pins
the oracle (tests/test_golden_fixtures.py, no GPU) and the CUDA path (same file, -m gpu) to the
reference's code rather than to a restatement of it.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
"""
import importlib.util
import math
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = "/root/reference/photon_weave"
sys.path.insert(0, ROOT)

from oracle import operators as ops  # noqa: E402  (input generation)
from oracle import rng_threefry as orng  # noqa: E402  (third-party PRNG restated)


def _install_shims():
    jnp = types.ModuleType("jax.numpy")
    jnp.__dict__.update({k: getattr(np, k) for k in dir(np) if not k.startswith("_")})
    jnp.ndarray = np.ndarray
    jax = types.ModuleType("jax")
    jax.numpy = jnp

    def jit(fn=None, **_kw):
        return (lambda f: f) if fn is None else fn

    def vmap(fn, in_axes=0):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else None

        def mapped(*args):
            ax = axes if axes is not None else (in_axes,) * len(args)
            n = next(a.shape[0] for a, x in zip(args, ax) if x is not None)
            return np.stack([fn(*[a[i] if x is not None else a for a, x in zip(args, ax)]) for i in range(n)])

        return mapped

    rnd = types.ModuleType("jax.random")
    rnd.choice = lambda key, a, p=None: np.int64(orng.choice(key, int(a), np.asarray(p)))
    rnd.PRNGKey = orng.PRNGKey
    rnd.split = orng.split
    jax.jit, jax.vmap, jax.random = jit, vmap, rnd
    oe = types.ModuleType("opt_einsum")
    oe.contract = lambda expr, *operands, backend=None, **kw: np.einsum(expr, *operands, optimize=True)
    sys.modules.update({"jax": jax, "jax.numpy": jnp, "jax.random": rnd, "opt_einsum": oe})


def _bare(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    _install_shims()
    for pkg, sub in (("photon_weave", ""), ("photon_weave.core", "core"), ("photon_weave.extra", "extra"),
                     ("photon_weave.state", "state")):
        _bare(pkg, os.path.join(REF, sub))
    _load("photon_weave.state.interfaces", "state/interfaces.py")
    _load("photon_weave.core.meta", "core/meta.py")
    _load("photon_weave.extra.einsum_constructor", "extra/einsum_constructor.py")
    sys.modules["photon_weave.core"].kernels = _load("photon_weave.core.kernels", "core/kernels.py")
    return _load("photon_weave.core.adapters", "core/adapters.py")


CASES = [  # (dims, targets)
    ((2, 3, 2), (1,)), ((3, 2, 4), (2, 0)), ((2, 2, 2, 2), (3, 1)), ((4, 3), (0,)), ((2, 3), (1, 0)),
    ((3, 3, 2), (0, 1, 2)), ((5,), (0,)), ((2, 2, 3, 2), (2,)),
]


def main():
    rad = load_reference()
    out = {}
    for c, (dims, tg) in enumerate(CASES):
        N = math.prod(dims)
        t = math.prod(dims[i] for i in tg)
        psi = ops.random_state(N, 100 + c)
        rho = ops.random_density(N, 200 + c, rank=3)
        u = ops.random_unitary(t, 300 + c)
        ks = np.stack([ops.random_unitary(t, 400 + c + 7 * k) for k in range(3)]) / math.sqrt(3.0)
        pre = f"c{c}_"
        out[pre + "dims"], out[pre + "targets"] = np.array(dims), np.array(tg)
        out[pre + "psi"], out[pre + "rho"], out[pre + "u"], out[pre + "ks"] = psi, rho, u, ks
        for uc in (False, True):
            s = "_uc" if uc else "_ord"
            out[pre + "apply_vec" + s] = rad.apply_operation_vector(dims, tg, psi, u, use_contraction=uc)
            out[pre + "apply_mat" + s] = rad.apply_operation_matrix(dims, tg, rho, u, use_contraction=uc)
            out[pre + "kraus" + s] = rad.apply_kraus_matrix(dims, tg, rho, ks, use_contraction=uc)
            out[pre + "trace" + s] = rad.trace_out_matrix(dims, tg, rho, use_contraction=uc)
        for seed in (0, 1, 2, 3, 7, 11):
            key = orng.PRNGKey(seed)
            o, post, probs, _ = rad.measure_vector_with_probs(dims, tg, psi, key)
            out[pre + f"mv{seed}_o"], out[pre + f"mv{seed}_post"] = np.int64(o), post
            out[pre + "mv_probs"] = probs
            o, post, probs, _ = rad.measure_matrix_with_probs(dims, tg, rho, key)
            out[pre + f"mm{seed}_o"], out[pre + f"mm{seed}_post"] = np.int64(o), post
            out[pre + "mm_probs"] = probs
            o2, post2, _ = rad.measure_vector(dims, tg, psi, key)
            assert int(o2) == int(out[pre + f"mv{seed}_o"]) and np.array_equal(post2, out[pre + f"mv{seed}_post"])
            if len(tg) < len(dims):
                o, post, _ = rad.measure_povm_matrix(dims, tg, ks, rho, key)
                out[pre + f"povm{seed}_o"], out[pre + f"povm{seed}_post"] = np.int64(o), post
        if len(tg) < len(dims):
            p, red = rad.measure_vector_expectation(dims, tg, psi)
            out[pre + "ev_probs"], out[pre + "ev_post"] = p, red
            p, red = rad.measure_matrix_expectation(dims, tg, rho)
            out[pre + "em_probs"], out[pre + "em_post"] = p, red
    out["ncases"] = np.int64(len(CASES))
    path = os.path.join(HERE, "seam_golden.npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in out.items()})
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()

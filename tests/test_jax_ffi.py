"""
The jax.ffi front end (photon_weave_b200/core/jax_ffi.py + csrc/pw_xla_ffi.cc) against the
reference's own transform tests: jit, vmap over states
(tests/state/test_shape_planning.py:104-121) and grad through expectation-mode measurement
(tests/state/utils/test_measurements.py:337-371, tests/state/test_envelope.py:846-870).

GATED: needs jax + a CUDA device + libpwb200_xla.so (`make -C photon_weave_b200/csrc xla`).  The
build image has no jax, so this module skips cleanly there; the same rules run through PyTorch
autograd in tests/test_gpu_transforms.py.
"""

import math

import numpy as np
import pytest

jax = pytest.importorskip("jax")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def jf():
    jax.config.update("jax_enable_x64", True)
    try:
        from photon_weave_b200.core import jax_ffi

        jax_ffi.register()
    except ImportError as exc:
        pytest.skip(str(exc))
    return jax_ffi


def test_apply_under_jit_matches_kron(jf):
    import jax.numpy as jnp

    op = jnp.array([[0, 1], [1, 0]], dtype=jnp.complex128)
    psi = jnp.zeros((4, 1), dtype=jnp.complex128).at[0, 0].set(1.0)
    fn = jax.jit(lambda s: jf.apply_operation_vector((2, 2), (0,), s, op))
    assert jnp.allclose(fn(psi), jnp.kron(op, jnp.eye(2)) @ psi)


def test_vmap_apply_op_vector(jf):
    import jax.numpy as jnp

    op = jnp.array([[0, 1], [1, 0]], dtype=jnp.complex128)
    batch = jnp.zeros((2, 4, 1), dtype=jnp.complex128).at[0, 0, 0].set(1.0).at[1, 1, 0].set(1.0)
    out = jax.vmap(lambda s, o: jf.apply_operation_vector((2, 2), (0,), s, o), in_axes=(0, None))(batch, op)
    full = jnp.kron(op, jnp.eye(2))
    assert jnp.allclose(out, jax.vmap(lambda s: full @ s)(batch))


def test_expectation_gradients(jf):
    import jax.numpy as jnp

    def p0_vec(theta):
        state = jnp.zeros((4, 1), dtype=jnp.complex128).at[0, 0].set(jnp.cos(theta)).at[2, 0].set(jnp.sin(theta))
        probs, _ = jf.measure_vector_expectation((2, 2), (0,), state / jnp.linalg.norm(state))
        return probs[0]

    def p0_mat(theta):
        state = jnp.zeros((4, 1), dtype=jnp.complex128).at[0, 0].set(jnp.cos(theta)).at[2, 0].set(jnp.sin(theta))
        probs, _ = jf.measure_matrix_expectation((2, 2), (0,), state @ jnp.conj(state).T)
        return probs[0]

    for fn in (p0_vec, p0_mat):
        g = jax.grad(fn)(0.3)
        assert jnp.isfinite(g) and abs(float(g) + math.sin(0.6)) < 1e-9


def test_apply_gradient_matches_einsum(jf):
    import jax.numpy as jnp

    rng = np.random.default_rng(0)
    psi = jnp.asarray(rng.normal(size=(24, 1)) + 1j * rng.normal(size=(24, 1)))
    op = jnp.asarray(rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)))

    def loss(f):
        return lambda s, o: jnp.sum(jnp.abs(f(s, o)) ** 2) + jnp.real(jnp.sum(f(s, o)))

    ours = loss(lambda s, o: jf.apply_operation_vector((3, 4, 2), (1,), s, o))
    ref = loss(lambda s, o: jnp.einsum("ab,xbz->xaz", o, s.reshape(3, 4, 2)).reshape(24, 1))
    for a, b in zip(jax.grad(ours, argnums=(0, 1))(psi, op), jax.grad(ref, argnums=(0, 1))(psi, op)):
        assert jnp.allclose(a, b, rtol=1e-10, atol=1e-12)

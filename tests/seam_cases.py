"""
The reference's seam-level tests, restated once and
run against two backends: the NumPy oracle (``-m "not gpu"``, pins the oracle)
and the CUDA product path (``-m gpu``).  Every case cites the reference test it
restates; expected values are the literals / kron constructions those tests use.

A backend is an object with
  ad            module exposing the photon_weave.core.adapters function set
  arr(np_array) host ndarray -> backend array
  np(x)         backend array -> host ndarray
  key(seed)     PRNG key for ``seed`` (jax.random.PRNGKey semantics)
"""

from __future__ import annotations

import numpy as np

from oracle import operators as ops

C128 = np.complex128


def basis(index, dim, dtype=C128):
    v = np.zeros((dim, 1), dtype=dtype)
    v[index, 0] = 1
    return v


def density(v):
    return v @ np.conj(v.T)


X = ops.x_operator()


# --- tests/core/test_core_adapters.py -------------------------------------


def case_apply_vector_matches_full_kron(b, use_contraction):
    """test_core_adapters.py:28-39."""
    out = b.ad.apply_operation_vector(
        (2, 2), (0,), b.arr(basis(0, 4)), b.arr(X), use_contraction=use_contraction
    )
    expected = np.kron(X, np.eye(2)) @ basis(0, 4)
    assert b.np(out).shape == (4, 1)
    assert np.allclose(b.np(out), expected)


def case_apply_matrix_matches_full_kron(b, use_contraction):
    """test_core_adapters.py:42-54."""
    rho = density(basis(0, 4))
    out = b.ad.apply_operation_matrix(
        (2, 2), (0,), b.arr(rho), b.arr(X), use_contraction=use_contraction
    )
    full = np.kron(X, np.eye(2))
    assert np.allclose(b.np(out), full @ rho @ full.conj().T)


def case_kraus_single_op_matches_unitary(b):
    """test_core_adapters.py:57-66."""
    rho = density(basis(0, 4))
    out = b.ad.apply_kraus_matrix((2, 2), (0,), b.arr(rho), b.arr(np.stack([X])))
    full = np.kron(X, np.eye(2))
    assert np.allclose(b.np(out), full @ rho @ full.conj().T)


def case_kraus_doctest_literal(b, use_contraction):
    """
    Inputs of the docstring examples at core/kernels.py:169-178 and
    core/adapters.py:427-436 (rho = I/4, K = |0><1| on qubit 0).  The printed
    output there, diag(0, 0, .5, .5), is NOT what the reference's code computes
    (the docstrings are never executed: no doctest runner in pytest.ini/tox.ini,
    and the value contradicts the executed kron-equivalence tests
    test_core_adapters.py:57-66): (K x I) rho (K x I)^dagger = diag(.25, .25, 0, 0).
    """
    rho = np.eye(4, dtype=C128) / 4
    k = np.array([[0, 1], [0, 0]], dtype=C128)
    out = b.ad.apply_kraus_matrix(
        (2, 2), (0,), b.arr(rho), b.arr(np.stack([k])), use_contraction=use_contraction
    )
    full = np.kron(k, np.eye(2))
    assert np.allclose(b.np(out), full @ rho @ full.conj().T)
    assert np.allclose(b.np(out), np.diag([0.25, 0.25, 0.0, 0.0]))


def case_measure_vector_collapses(b):
    """test_core_adapters.py:69-77."""
    o, post, _ = b.ad.measure_vector((2, 2), (0,), b.arr(basis(0, 4)), b.key(0))
    assert int(o) == 0
    assert b.np(post).shape == (2, 1)
    assert np.allclose(b.np(post), basis(0, 2))


def case_measure_matrix_collapses(b):
    """test_core_adapters.py:80-88."""
    o, post, _ = b.ad.measure_matrix(
        (2, 2), (0,), b.arr(density(basis(0, 4))), b.key(0)
    )
    assert int(o) == 0
    assert b.np(post).shape == (2, 2)
    assert np.allclose(b.np(post), np.array([[1, 0], [0, 0]]))


def case_measure_povm_projective(b):
    """test_core_adapters.py:91-101."""
    rho = density(basis(0, 4))
    operators = np.stack([np.diag([1, 0]), np.diag([0, 1])]).astype(C128)
    o, post, _ = b.ad.measure_povm_matrix(
        (2, 2), (0,), b.arr(operators), b.arr(rho), b.key(0)
    )
    assert int(o) == 0
    assert np.allclose(b.np(post), rho)


def case_trace_out_reduces(b):
    """test_core_adapters.py:104-109."""
    red = b.ad.trace_out_matrix((2, 2), (0,), b.arr(density(basis(0, 4))))
    assert b.np(red).shape == (2, 2)
    assert np.allclose(b.np(red), np.array([[1, 0], [0, 0]]))


def case_measure_vector_with_probs(b):
    """test_core_adapters.py:112-120."""
    o, post, probs, _ = b.ad.measure_vector_with_probs(
        (2, 2), (0,), b.arr(basis(1, 4)), b.key(0)
    )
    assert int(o) == 0
    assert np.allclose(b.np(probs), [1.0, 0.0])
    assert np.allclose(b.np(post), basis(1, 2))


def case_measure_matrix_with_probs(b):
    """test_core_adapters.py:123-131."""
    o, post, probs, _ = b.ad.measure_matrix_with_probs(
        (2, 2), (0,), b.arr(density(basis(1, 4))), b.key(0)
    )
    assert int(o) == 0
    assert np.allclose(b.np(probs), [1.0, 0.0])
    assert np.allclose(b.np(post), np.array([[0, 0], [0, 1]]))


def case_measure_vector_expectation(b):
    """test_core_adapters.py:134-144."""
    psi = np.zeros((4, 1), dtype=C128)
    psi[0, 0] = psi[2, 0] = 1 / np.sqrt(2)
    probs, post = b.ad.measure_vector_expectation((2, 2), (0,), b.arr(psi))
    assert np.allclose(b.np(probs), [0.5, 0.5])
    assert b.np(post).shape == (2, 2)
    assert np.allclose(b.np(post), np.array([[1, 0], [0, 0]]))


def case_measure_matrix_expectation(b):
    """test_core_adapters.py:147-154."""
    probs, post = b.ad.measure_matrix_expectation(
        (2, 2), (0,), b.arr(density(basis(0, 4)))
    )
    assert np.allclose(b.np(probs), [1.0, 0.0])
    assert np.allclose(b.np(post), np.array([[1, 0], [0, 0]]))


# --- tests/core/test_contraction_caching.py ------------------------------


def case_contraction_matches_ordered(b):
    """test_contraction_caching.py:17-61."""
    psi = basis(1, 4)
    rho = density(psi)
    a = b.ad.apply_operation_vector((2, 2), (0,), b.arr(psi), b.arr(X), use_contraction=True)
    c = b.ad.apply_operation_vector((2, 2), (0,), b.arr(psi), b.arr(X), use_contraction=False)
    assert np.allclose(b.np(a), b.np(c))
    a = b.ad.apply_operation_matrix((2, 2), (0,), b.arr(rho), b.arr(X), use_contraction=True)
    c = b.ad.apply_operation_matrix((2, 2), (0,), b.arr(rho), b.arr(X), use_contraction=False)
    assert np.allclose(b.np(a), b.np(c))
    a = b.ad.trace_out_matrix((2, 2), (1,), b.arr(rho), use_contraction=True)
    c = b.ad.trace_out_matrix((2, 2), (1,), b.arr(rho), use_contraction=False)
    assert np.allclose(b.np(a), b.np(c))
    assert np.allclose(b.np(a), np.array([[0, 0], [0, 1]]))


def case_measure_deterministic_with_key(b):
    """test_contraction_caching.py:64-83."""
    o, post, _ = b.ad.measure_vector((2,), (0,), b.arr(basis(1, 2)), b.key(0))
    assert int(o) == 1 and b.np(post).shape == (1, 1)
    o, post, _ = b.ad.measure_matrix((2,), (0,), b.arr(density(basis(0, 2))), b.key(1))
    assert int(o) == 0 and b.np(post).shape == (1, 1)


# --- tests/core/test_jitted_meta.py ----------------------------------------


def case_complex64_operator_is_promoted(b):
    """test_jitted_meta.py:14-22: complex64 operator on a complex128 state."""
    x64 = X.astype(np.complex64)
    out = b.ad.apply_operation_vector((2, 2), (0,), b.arr(basis(0, 4)), b.arr(x64))
    assert b.np(out).dtype == np.complex128
    assert np.allclose(b.np(out), basis(2, 4))


# --- tests/state/utils/test_operations_contraction.py ---------------------


def case_noncontiguous_middle_target(b, use_contraction):
    """test_operations_contraction.py:32-56: dims (2,2,2), middle target."""
    psi = basis(2, 8)
    full = np.kron(np.eye(2), np.kron(X, np.eye(2)))
    out = b.ad.apply_operation_vector(
        (2, 2, 2), (1,), b.arr(psi), b.arr(X), use_contraction=use_contraction
    )
    assert np.allclose(b.np(out), full @ psi)
    rho = density(psi)
    out = b.ad.apply_operation_matrix(
        (2, 2, 2), (1,), b.arr(rho), b.arr(X), use_contraction=use_contraction
    )
    assert np.allclose(b.np(out), full @ rho @ full.conj().T)


def case_dims_2_3_last_target(b, use_contraction):
    """test_operations_contraction.py:59-90: dims (2,3), target 1, 3x3 permutation."""
    op = np.array([[0, 1, 0], [1, 0, 0], [0, 0, 1]], dtype=C128)
    psi = basis(4, 6)
    full = np.kron(np.eye(2), op)
    out = b.ad.apply_operation_vector(
        (2, 3), (1,), b.arr(psi), b.arr(op), use_contraction=use_contraction
    )
    assert np.allclose(b.np(out), full @ psi)
    rho = density(psi)
    out = b.ad.apply_operation_matrix(
        (2, 3), (1,), b.arr(rho), b.arr(op), use_contraction=use_contraction
    )
    assert np.allclose(b.np(out), full @ rho @ full.conj().T)


def case_kraus_last_target(b, use_contraction):
    """test_operations_contraction.py:93-108 and test_shape_planning.py:124-140."""
    rho = density(basis(1, 4))
    full = np.kron(np.eye(2), X)
    out = b.ad.apply_kraus_matrix(
        (2, 2), (1,), b.arr(rho), b.arr(np.stack([X])), use_contraction=use_contraction
    )
    assert np.allclose(b.np(out), full @ rho @ full.conj().T)
    rho = density(basis(0, 12))
    full = np.kron(np.eye(2), np.kron(np.eye(3), X))
    out = b.ad.apply_kraus_matrix(
        (2, 3, 2), (2,), b.arr(rho), b.arr(np.stack([X])), use_contraction=use_contraction
    )
    assert np.allclose(b.np(out), full @ rho @ full.conj().T)


# --- tests/state/utils/test_trace_out_contraction.py ----------------------


def case_trace_out_keep_0_2_of_2_3_2(b, use_contraction):
    """test_trace_out_contraction.py:18-45."""
    rho = density(basis(2, 12))
    out = b.ad.trace_out_matrix(
        (2, 3, 2), (0, 2), b.arr(rho), use_contraction=use_contraction
    )
    t = rho.reshape((2, 3, 2, 2, 3, 2))
    expected = np.trace(np.transpose(t, (0, 2, 1, 3, 5, 4)), axis1=2, axis2=5)
    assert np.allclose(b.np(out), expected.reshape(4, 4))


def case_trace_out_literal_4x4(b, use_contraction):
    """tests/state/test_composite_envelope.py:1331-1375: R (x) V (x) H, keep 0,1."""
    r = np.array([[1], [1j]]) / np.sqrt(2)
    v = np.array([[0], [1]], dtype=C128)
    h = np.array([[1], [0]], dtype=C128)
    rho = density(np.kron(np.kron(r, v), h))
    literal = np.zeros((8, 8), dtype=C128)
    literal[2, 2] = literal[6, 6] = 0.5
    literal[2, 6] = -0.5j
    literal[6, 2] = 0.5j
    assert np.allclose(rho, literal)
    out = b.ad.trace_out_matrix(
        (2, 2, 2), (0, 1), b.arr(rho), use_contraction=use_contraction
    )
    expected = np.array(
        [[0, 0, 0, 0], [0, 0.5, 0, -0.5j], [0, 0, 0, 0], [0, 0.5j, 0, 0.5]]
    )
    assert np.allclose(b.np(out), expected)


def case_trace_out_all_kept_is_identity(b):
    """core/adapters.py:706-709."""
    rho = density(basis(1, 4))
    out = b.ad.trace_out_matrix((2, 2), (0, 1), b.arr(rho))
    assert np.allclose(b.np(out), rho)


# --- container-level goldens restated at the seam --------------------------


def case_beam_splitter_hom(b, use_contraction):
    """tests/operation/test_composite_operation.py:15-71: BS(pi/4) on |1,1>, cutoff 3:
    amplitude 1j/sqrt2 at flat indices 2 and 6 of the 3x3 state."""
    bs = ops.beam_splitter(3, 3, np.pi / 4)
    psi = basis(4, 9)
    out = b.np(
        b.ad.apply_operation_vector(
            (3, 3), (0, 1), b.arr(psi), b.arr(bs), use_contraction=use_contraction
        )
    )
    expected = np.zeros((9, 1), dtype=C128)
    expected[2, 0] = expected[6, 0] = 1j / np.sqrt(2)
    assert np.allclose(out, expected, atol=1e-7)


def case_three_beam_splitters_probs(b):
    """tests/state/utils/test_measurements.py:66-134: |1,0,0,0>, BS(0,1), BS(0,2),
    BS(3,1) -> every mode measures (0.75, 0.25)."""
    bs = ops.beam_splitter(2, 2, np.pi / 4)
    dims = (2, 2, 2, 2)
    psi = b.arr(basis(8, 16))
    for tg in [(0, 1), (0, 2), (3, 1)]:
        psi = b.ad.apply_operation_vector(dims, tg, psi, b.arr(bs), use_contraction=True)
    rho = b.arr(density(b.np(psi)))
    for i in range(4):
        _, _, probs, _ = b.ad.measure_vector_with_probs(dims, (i,), psi, b.key(0))
        assert np.allclose(b.np(probs), [0.75, 0.25], atol=1e-5)
        _, _, probs, _ = b.ad.measure_matrix_with_probs(dims, (i,), rho, b.key(0))
        assert np.allclose(b.np(probs), [0.75, 0.25], atol=1e-5)


def case_expression_phase_flip(b):
    """tests/operation/test_composite_operation.py:74-91: expm(i pi n) on |1> (x) H."""
    from scipy.linalg import expm

    op = expm(1j * np.pi * ops.number_operator(2))
    psi = basis(2, 4)
    out = b.ad.apply_operation_vector((2, 2), (0,), b.arr(psi), b.arr(op))
    assert np.allclose(b.np(out), np.array([[0], [0], [-1], [0]]))


# --- error behaviour ---------------------------------


def case_error_behaviour(b):
    """core/adapters.py:317-327, 365-375, 711-715."""
    import pytest

    with pytest.raises(ValueError):
        b.ad.apply_operation_vector((2, 2), (0,), b.arr(basis(0, 8)), b.arr(X))
    with pytest.raises(AssertionError):
        b.ad.apply_operation_vector((2, 2), (0,), b.arr(basis(0, 4)), b.arr(np.eye(3)))
    with pytest.raises(ValueError):
        b.ad.apply_operation_matrix((2, 2), (0,), b.arr(density(basis(0, 8))), b.arr(X))
    with pytest.raises(AssertionError):
        b.ad.apply_operation_matrix(
            (2, 2), (0,), b.arr(density(basis(0, 4))), b.arr(np.eye(3))
        )
    with pytest.raises(ValueError):
        b.ad.trace_out_matrix((2, 2), (0,), b.arr(density(basis(0, 8))))


CASES_BOTH_ROUTES = [
    case_apply_vector_matches_full_kron,
    case_apply_matrix_matches_full_kron,
    case_kraus_doctest_literal,
    case_noncontiguous_middle_target,
    case_dims_2_3_last_target,
    case_kraus_last_target,
    case_trace_out_keep_0_2_of_2_3_2,
    case_trace_out_literal_4x4,
    case_beam_splitter_hom,
]

CASES_SINGLE = [
    case_kraus_single_op_matches_unitary,
    case_measure_vector_collapses,
    case_measure_matrix_collapses,
    case_measure_povm_projective,
    case_trace_out_reduces,
    case_measure_vector_with_probs,
    case_measure_matrix_with_probs,
    case_measure_vector_expectation,
    case_measure_matrix_expectation,
    case_contraction_matches_ordered,
    case_measure_deterministic_with_key,
    case_complex64_operator_is_promoted,
    case_trace_out_all_kept_is_identity,
    case_three_beam_splitters_probs,
    case_expression_phase_flip,
    case_error_behaviour,
]

"""
Pins the NumPy oracle (oracle/adapters_np.py) against every seam-level golden
the reference's tests hold for the hot path (tests/seam_cases.py cites each),
and cross-checks it on random inputs against an independent dense-kron
construction.  CPU only.
"""

import numpy as np
import pytest

import seam_cases as sc
from oracle import adapters_np as ad
from oracle import einsum_strings as es
from oracle import operators as ops
from oracle import rng_threefry as rng


class OracleBackend:
    ad = ad

    @staticmethod
    def arr(x):
        return np.asarray(x)

    @staticmethod
    def np(x):
        return np.asarray(x)

    @staticmethod
    def key(seed):
        return rng.PRNGKey(seed)


B = OracleBackend()


@pytest.mark.parametrize("use_contraction", [False, True])
@pytest.mark.parametrize("case", sc.CASES_BOTH_ROUTES, ids=lambda c: c.__name__)
def test_seam_case_both_routes(case, use_contraction):
    case(B, use_contraction)


@pytest.mark.parametrize("case", sc.CASES_SINGLE, ids=lambda c: c.__name__)
def test_seam_case(case):
    case(B)


def test_einsum_strings_match_reference_goldens():
    """tests/utils/test_einsum_constructor.py:15-16 ('abcadc->bd'); SURVEY section 8a
    strings 'eb,abcd->aecd' (vector, target 1 of 3) and the 5-mode matrix string."""
    operands, out = es.trace_out_matrix(3, [1])
    assert es.to_string(operands, out) == "abcadc->bd"
    operands, out = es.apply_operator_vector(3, [1])
    assert es.to_string(operands, out) == "eb,abcd->aecd"
    operands, out = es.apply_operator_matrix(5, [1, 3])
    assert es.to_string(operands, out) == "klbd,abcdefghij,mngi->akclefmhnj"
    operands, out = es.kraus_matrix(2, [0])
    assert es.to_string(operands, out) == "kea,abcd,kfc->ebfd"


def test_kraus_string_shares_one_summed_label():
    """core/adapters.py:194-207: the Kraus label leads both operators and is summed."""
    operands, out = es.kraus_matrix(2, [0])
    k = operands[0][0]
    assert operands[2][0] == k and k not in out and k not in operands[1]


def _full_operator(dims, targets, op):
    """Independent construction: embed op (indices in TARGET-GIVEN order) densely."""
    n = len(dims)
    N = int(np.prod(dims))
    td = [dims[t] for t in targets]
    full = np.zeros((N, N), dtype=np.complex128)
    for col in range(N):
        idx = list(np.unravel_index(col, dims))
        b = int(np.ravel_multi_index([idx[t] for t in targets], td))
        for a in range(op.shape[0]):
            if op[a, b] == 0:
                continue
            new = list(idx)
            for t, v in zip(targets, np.unravel_index(a, td)):
                new[t] = int(v)
            full[int(np.ravel_multi_index(new, dims)), col] += op[a, b]
    return full


@pytest.mark.parametrize(
    "dims,targets",
    [
        ((2, 3, 2), (1,)),
        ((3, 2, 4), (2, 0)),
        ((3, 2, 4), (0, 2)),
        ((2, 3, 2, 2), (3, 1)),
        ((4, 3), (1, 0)),
        ((5,), (0,)),
        ((2, 3, 2, 3), (2, 0, 3)),
    ],
)
def test_oracle_matches_dense_embedding_random(dims, targets):
    N = int(np.prod(dims))
    t = int(np.prod([dims[i] for i in targets]))
    op = ops.random_unitary(t, 7) * 1.3
    full = _full_operator(dims, targets, op)
    psi = ops.random_state(N, 1)
    rho = ops.random_density(N, 2)
    kr = np.stack([ops.random_unitary(t, 3) * 0.6, ops.random_unitary(t, 4) * 0.8])
    kfull = [_full_operator(dims, targets, k) for k in kr]
    for uc in (False, True):
        assert np.allclose(
            ad.apply_operation_vector(dims, targets, psi, op, uc), full @ psi, atol=1e-13
        )
        assert np.allclose(
            ad.apply_operation_matrix(dims, targets, rho, op, uc),
            full @ rho @ full.conj().T,
            atol=1e-13,
        )
        assert np.allclose(
            ad.apply_kraus_matrix(dims, targets, rho, kr, uc),
            sum(k @ rho @ k.conj().T for k in kfull),
            atol=1e-13,
        )


def test_trace_out_order_semantics_differ_between_routes():
    """adapters.py:720,732 (given order) vs :80-81 (tensor order)."""
    dims = (2, 3, 2)
    rho = ops.random_density(12, 5)
    given = ad.trace_out_matrix(dims, (2, 0), rho, use_contraction=False)
    tensor = ad.trace_out_matrix(dims, (2, 0), rho, use_contraction=True)
    t = rho.reshape(2, 3, 2, 2, 3, 2)
    keep02 = np.einsum("abcdbf->acdf", t)
    assert np.allclose(tensor, keep02.reshape(4, 4))
    assert np.allclose(given, np.transpose(keep02, (1, 0, 3, 2)).reshape(4, 4))


def test_trace_out_vector_equals_expand_then_trace():
    """state/utils/trace_out.py:17-25."""
    dims = (3, 2, 2)
    psi = ops.random_state(12, 9)
    red = ad.trace_out_vector(dims, (1,), psi)
    t = psi.reshape(3, 2, 2)
    assert np.allclose(red, np.einsum("abc,adc->bd", t, t.conj()))


def test_measure_post_state_and_probs_random():
    dims = (3, 2, 2)
    psi = ops.random_state(12, 11)
    key = rng.PRNGKey(5)
    o, post, probs, _ = ad.measure_vector_with_probs(dims, (1, 2), psi, key)
    t = psi.reshape(3, 4)
    p = np.sum(np.abs(t) ** 2, axis=0)
    assert np.allclose(probs, p / p.sum())
    assert np.allclose(post[:, 0], t[:, o] / np.sqrt(probs[o] + 1e-18))
    rho = psi @ psi.conj().T
    o2, post2, probs2, _ = ad.measure_matrix_with_probs(dims, (1, 2), rho, key)
    assert o2 == o and np.allclose(probs2, probs)
    assert np.allclose(post2, post @ post.conj().T)


def test_loss_channel_is_trace_preserving():
    ks = ops.loss_channel(6, 0.1)
    s = sum(k.conj().T @ k for k in ks)
    assert np.allclose(s, np.eye(6))

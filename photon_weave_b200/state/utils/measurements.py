"""
Projective / POVM measurement bridges (reference: photon_weave/state/utils/measurements.py:20-682).
The PNR detector model (measurements.py:685-1112) is host arithmetic on a handful of
integers: the projective measurement inside it is the CUDA path, the binomial-loss /
Poisson-dark-count / Gaussian-jitter draws are taken from a counter-based stream seeded by
the threefry key (statistically the reference's model; NOT bit-identical to
``jax.random.binomial/poisson/normal`` for efficiency < 1 or dark_rate > 0 -- with the
defaults, efficiency = 1 and dark_rate = 0, no random draw influences the result).

Key threading follows the reference exactly:
* legacy sequential paths: one key per target from ``_key_sequence`` (``split(key, n+1)[1:]``),
* jit paths: ``use_key, next_key = borrow_key(key)`` then ONE joint draw over prod(target dims)
  and a mixed-radix decode of the flat outcome (measurements.py:200-208).
"""

from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

from ...core import adapters
from ...core import rng as _rng
from ...core.meta import DimsMeta
from ...core.rng import borrow_key
from ...photon_weave import Config
from .shape_planning import ShapePlan


def _ensure_key(key):
    """measurements.py:20-24."""
    return key if key is not None else Config().random_key


def _borrow_optional(key):
    """measurements.py:27-47."""
    if key is None:
        return Config().random_key, None
    k = _rng.split(key, 2)
    return k[0], k[1]


def _key_sequence(key, count: int):
    """measurements.py:50-80: drop the first subkey, return (keys, leftover)."""
    if key is None:
        cfg = Config()
        return [cfg.random_key for _ in range(count)], None
    k = _rng.split(key, count + 1)
    return [k[i] for i in range(1, count + 1)], k[0]


def _dims_targets(meta, state_objs, target_states, target_indices):
    if isinstance(meta, ShapePlan):
        dims, tgt = list(meta.dims), list(meta.target_indices)
    elif meta is None:
        objs = list(state_objs)
        dims = [s.dimensions for s in objs]
        tgt = [i for i, s in enumerate(objs) if s in target_states]
    else:
        dims, tgt = list(meta.state_dims), list(meta.target_indices)
    if target_indices is not None:
        tgt = list(target_indices)
    return dims, tgt


def _decode(outcome, dims, tgt, target_states) -> Dict:
    """Mixed-radix decode of the joint outcome (measurements.py:200-208)."""
    tmp = int(outcome)
    decoded = []
    for dim in reversed([dims[i] for i in tgt]):
        decoded.append(tmp % dim)
        tmp //= dim
    decoded.reverse()
    return {ts: int(v) for ts, v in zip(target_states, decoded)}


def _sequential(state_objs, target_states, product_state, prob_callback, key, is_matrix: bool):
    """Legacy path (measurements.py:83-138, 212-281): targets measured one at a time."""
    objs = list(state_objs)
    keys, _ = _key_sequence(key, len(target_states))
    outcomes: Dict = {}
    for state, use_key in zip(target_states, keys):
        dims = [s.dimensions for s in objs]
        idx = objs.index(state)
        # jnp.take / plain slicing in the reference: the collapsed slice is NOT rescaled
        outcome, product_state, probs = adapters._measure(dims, (idx,), product_state, use_key,
                                                           is_matrix, rescale=False)
        if prob_callback is not None:
            prob_callback(state, probs)
        outcomes[state] = int(outcome)
        objs.remove(state)
    return outcomes, product_state


def measure_vector(state_objs, target_states, product_state, prob_callback: Optional[Callable] = None,
                   key=None, target_indices=None):
    """measurements.py:83-138 (the collapsed slice is returned un-normalised, as jnp.take does)."""
    return _sequential(state_objs, target_states, product_state, prob_callback, key, False)


def measure_matrix(state_objs, target_states, product_state, prob_callback: Optional[Callable] = None,
                   key=None, target_indices=None):
    """measurements.py:212-281."""
    return _sequential(state_objs, target_states, product_state, prob_callback, key, True)


def measure_vector_jit(state_objs, target_states, product_state, key, meta=None, target_indices=None):
    """measurements.py:141-209 -> (outcomes, post, next_key)."""
    use_key, next_key = borrow_key(key)
    dims, tgt = _dims_targets(meta, state_objs, target_states, target_indices)
    outcome, post, _ = adapters.measure_vector(dims, tgt, product_state, use_key)
    return _decode(outcome, dims, tgt, target_states), post, next_key


def measure_matrix_jit(state_objs, target_states, product_state, key, meta=None, target_indices=None):
    """measurements.py:284-343."""
    use_key, next_key = borrow_key(key)
    dims, tgt = _dims_targets(meta, state_objs, target_states, target_indices)
    outcome, post, _ = adapters.measure_matrix(dims, tgt, product_state, use_key)
    return _decode(outcome, dims, tgt, target_states), post, next_key


def measure_vector_jit_with_probs(state_objs, target_states, product_state, key, meta=None,
                                  target_indices=None):
    """measurements.py:346-417."""
    use_key, next_key = borrow_key(key)
    dims, tgt = _dims_targets(meta, state_objs, target_states, target_indices)
    outcome, post, probs, _ = adapters.measure_vector_with_probs(dims, tgt, product_state, use_key)
    return _decode(outcome, dims, tgt, target_states), post, probs, next_key


def measure_matrix_jit_with_probs(state_objs, target_states, product_state, key, meta=None,
                                  target_indices=None):
    """measurements.py:420-491."""
    use_key, next_key = borrow_key(key)
    dims, tgt = _dims_targets(meta, state_objs, target_states, target_indices)
    outcome, post, probs, _ = adapters.measure_matrix_with_probs(dims, tgt, product_state, use_key)
    return _decode(outcome, dims, tgt, target_states), post, probs, next_key


def measure_vector_expectation(state_objs, target_states, product_state, meta=None):
    """measurements.py:494-534."""
    dims, tgt = _dims_targets(meta, state_objs, target_states, None)
    return adapters.measure_vector_expectation(dims, tgt, product_state)


def measure_matrix_expectation(state_objs, target_states, product_state, meta=None):
    """measurements.py:537-577."""
    dims, tgt = _dims_targets(meta, state_objs, target_states, None)
    return adapters.measure_matrix_expectation(dims, tgt, product_state)


def measure_POVM_matrix(state_objs, target_states, operators, product_state, key=None, meta=None):
    """measurements.py:580-645 -> (outcome index, post state)."""
    key = _ensure_key(key)
    if meta is None:
        objs = list(state_objs)
        dims = [s.dimensions for s in objs]
        tgt = [objs.index(s) for s in target_states]
    else:
        dims, tgt = list(meta.state_dims), list(meta.target_indices)
    target_dim = int(np.prod([dims[i] for i in tgt]))
    expected = (target_dim, target_dim)
    for op in operators:
        if tuple(op.shape) != expected:
            raise AssertionError(f"POVM operator shape {tuple(op.shape)} does not match expected {expected}")
    use_key, _ = borrow_key(key)
    outcome, post, _ = adapters.measure_povm_matrix(dims, tgt, list(operators), product_state, use_key)
    return outcome, post


def measure_POVM_matrix_jit(state_objs, target_states, operators, product_state, key):
    """measurements.py:648-682 -> (outcome, post, key)."""
    objs = list(state_objs)
    dims = [s.dimensions for s in objs]
    tgt = [objs.index(s) for s in target_states]
    return adapters.measure_povm_matrix(dims, tgt, operators, product_state, key)


# --------------------------------------------------------------------------------------
# photon-number-resolving detection (measurements.py:685-893)
# --------------------------------------------------------------------------------------


def _host_stream(key) -> np.random.Generator:
    """Deterministic NumPy stream keyed by the two threefry words of ``key``."""
    k = np.asarray(key).astype(np.uint32).reshape(-1)
    return np.random.Generator(np.random.Philox(key=int(k[0]) << 32 | int(k[1])))


def _pnr_noise(key, target_count: int, dark_mean: float, jitter_std: float):
    """Dark counts ~ Poisson(dark_mean), timing jitter ~ N(0, jitter_std) (measurements.py:685-703)."""
    key = _ensure_key(key)
    k = _rng.split(key, 2)
    dark = _host_stream(k[0]).poisson(lam=dark_mean, size=(target_count,)) if dark_mean > 0 \
        else np.zeros((target_count,), dtype=np.int64)
    if jitter_std > 0:
        jitter = _host_stream(k[1]).normal(size=(target_count,)) * jitter_std
    else:
        jitter = np.zeros((target_count,))
    return dark, jitter


def _binomial(key, n: int, p: float) -> int:
    p = float(min(max(p, 0.0), 1.0))
    if p >= 1.0 or n <= 0:
        return int(max(n, 0))
    if p <= 0.0:
        return 0
    return int(_host_stream(key).binomial(int(n), p))


def _measure_pnr(is_matrix: bool, state_objs, target_states, product_state, efficiency, dark_rate,
                 detection_window, jitter_std, key, meta, use_jit):
    use_jit_flag = isinstance(meta, (DimsMeta, ShapePlan)) or (Config().use_jit if use_jit is None else use_jit)
    if use_jit_flag:
        if key is None:
            raise ValueError("key is required for jitted PNR measurement (static path)")
        if meta is None:
            # measurements.py:781-783: the PNR path indexes the targets in the GIVEN order (the
            # joint draw and the decode both follow target_states), unlike measure_*_jit above
            objs = list(state_objs)
            dims = [s.dimensions for s in objs]
            tgt = [objs.index(s) for s in target_states]
        else:
            dims, tgt = _dims_targets(meta, state_objs, target_states, None)
        core = adapters.measure_matrix if is_matrix else adapters.measure_vector
        outcome, post, key_after = core(dims, tgt, product_state, key)
        decoded = list(_decode(outcome, dims, tgt, target_states).values())
        k3 = _rng.split(key_after, 3)
        binom_key, noise_seed, next_key = k3[0], k3[1], k3[2]
        draw_keys = _rng.split(binom_key, max(len(decoded), 1))
        detected = [_binomial(draw_keys[i], n, efficiency) for i, n in enumerate(decoded)]
        dark_seed = _rng.split(noise_seed, 2)[0]
        dark, jitter = _pnr_noise(dark_seed, len(target_states), dark_rate * detection_window, jitter_std)
        return ({ts: int(detected[i] + dark[i]) for i, ts in enumerate(target_states)}, post, jitter,
                next_key)
    base_key = _ensure_key(key)
    k2 = _rng.split(base_key, 2)
    meas_key, noise_key = k2[0], k2[1]
    legacy = measure_matrix if is_matrix else measure_vector
    outcomes, post = legacy(state_objs, target_states, product_state, key=meas_key)
    k3 = _rng.split(noise_key, 3)
    binom_key, noise_seed, next_key = k3[0], k3[1], k3[2]
    measured = {}
    subkey = binom_key
    for ts, outcome in outcomes.items():
        k = _rng.split(subkey, 2)
        subkey, draw_key = k[0], k[1]
        measured[ts] = _binomial(draw_key, int(outcome), efficiency)
    dark_seed = _rng.split(noise_seed, 2)[0]
    dark, jitter = _pnr_noise(dark_seed, len(target_states), dark_rate * detection_window, jitter_std)
    return ({ts: int(measured[ts] + dark[i]) for i, ts in enumerate(target_states)}, post, jitter,
            next_key)


def measure_pnr_vector(state_objs, target_states, product_state, efficiency: float = 1.0,
                       dark_rate: float = 0.0, detection_window: float = 1.0, jitter_std: float = 0.0,
                       key=None, meta=None, use_jit=None):
    """measurements.py:718-803 -> (counts per target, post state, jitter, next key)."""
    return _measure_pnr(False, state_objs, target_states, product_state, efficiency, dark_rate,
                        detection_window, jitter_std, key, meta, use_jit)


def measure_pnr_matrix(state_objs, target_states, product_state, efficiency: float = 1.0,
                       dark_rate: float = 0.0, detection_window: float = 1.0, jitter_std: float = 0.0,
                       key=None, meta=None, use_jit=None):
    """measurements.py:806-893."""
    return _measure_pnr(True, state_objs, target_states, product_state, efficiency, dark_rate,
                        detection_window, jitter_std, key, meta, use_jit)


# -- analytic detector model (measurements.py:896-1112): operator-sized host tables ---------


def fwhm_to_sigma(fwhm: float) -> float:
    return fwhm / (2.0 * np.sqrt(2.0 * np.log(2.0)))


def poisson_pmf_direct(k_max: int, mean: float) -> np.ndarray:
    from scipy.special import gammaln

    k = np.arange(k_max + 1, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        pmf = np.exp(-mean) * np.power(mean, k) / np.exp(gammaln(k + 1))
    total = pmf.sum()
    return pmf / total if total > 0 else pmf


def pnr_pmf_single_n(n_photons: int, eta: float, dark_rate: float, window: float,
                     max_counts=None) -> np.ndarray:
    """P(k clicks | n photons): binomial loss convolved with Poisson dark counts."""
    from scipy.special import gammaln

    if not (0.0 <= eta <= 1.0):
        raise ValueError("eta must be in [0, 1].")
    mu_d = dark_rate * window
    n = int(n_photons)
    if max_counts is None:
        spread = np.sqrt(max(n * eta * (1.0 - eta) + mu_d, 1e-12))
        cutoff = int(np.ceil(n * eta + mu_d + 8.0 * spread))
        max_counts = int(max(cutoff, n + int(5 + 5 * mu_d)))
    max_counts = max(0, int(max_counts))
    m = np.arange(n + 1, dtype=np.float64)
    coeff = np.exp(gammaln(n + 1) - gammaln(m + 1) - gammaln(n - m + 1))
    p_sig = coeff * np.power(eta, m) * np.power(1.0 - eta, n - m)
    p_dark = poisson_pmf_direct(max_counts, mu_d)
    k = np.arange(max_counts + 1)[None, :]
    lag = (k - m[:, None]).astype(np.int64)
    table = np.where(lag >= 0, p_dark[np.clip(lag, 0, max_counts)], 0.0)
    total = (p_sig[:, None] * table).sum(axis=0)
    norm = total.sum()
    return total / norm if norm > 0 else total


def pnr_transition_matrix(max_photons: int, eta: float, dark_rate: float, window: float,
                          max_counts=None) -> np.ndarray:
    max_photons = int(max_photons)
    if max_photons < 0:
        raise ValueError("max_photons must be non-negative.")
    if max_counts is None:
        max_counts = pnr_pmf_single_n(max_photons, eta, dark_rate, window, None).shape[0] - 1
    cols = [pnr_pmf_single_n(n, eta, dark_rate, window, int(max_counts)) for n in range(max_photons + 1)]
    return np.stack(cols, axis=1)


def pnr_povm(max_photons: int, eta: float, dark_rate: float, window: float, max_counts=None) -> np.ndarray:
    """Diagonal POVM elements E_k = diag_n P(k | n), shape (K, dim, dim)."""
    P = pnr_transition_matrix(max_photons, eta, dark_rate, window, max_counts)
    povm = np.zeros((P.shape[0], P.shape[1], P.shape[1]), dtype=np.float64)
    idx = np.arange(P.shape[1])
    povm[:, idx, idx] = P
    return povm


def gaussian_jitter_kernel(num_bins: int, bin_width: float, jitter_std: float) -> np.ndarray:
    """Column-stochastic matrix spreading a click time over neighbouring bins."""
    from scipy.special import erf

    if jitter_std <= 0.0:
        return np.eye(num_bins, dtype=np.float64)
    centers = (np.arange(num_bins, dtype=np.float64) + 0.5) * bin_width
    scale = np.sqrt(2.0) * jitter_std
    left = (centers[:, None] - 0.5 * bin_width - centers[None, :]) / scale
    right = (centers[:, None] + 0.5 * bin_width - centers[None, :]) / scale
    cols = 0.5 * (erf(right) - erf(left))
    sums = cols.sum(axis=0, keepdims=True)
    return np.where(sums > 0, cols / np.where(sums > 0, sums, 1.0), cols)

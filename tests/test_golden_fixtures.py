"""
tests/golden/seam_golden.npz holds outputs of the REFERENCE'S OWN core/kernels.py +
core/adapters.py executed on seeded inputs (tests/golden/make_golden.py; NumPy stands in for the
two third-party array libraries).  The oracle is pinned against them without a GPU; the CUDA
path is compared with the same vectors under -m gpu (1e-12 relative in complex128, outcomes
equal).
"""

import os

import numpy as np
import pytest

from oracle import adapters_np as oad
from oracle import rng_threefry as orng

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "seam_golden.npz"))
NC = int(G["ncases"])
SEEDS = (0, 1, 2, 3, 7, 11)


def rel(a, b):
    b = np.asarray(b)
    return float(np.abs(np.asarray(a) - b).max() / max(np.abs(b).max(), 1e-300))


class Oracle:
    tol = 1e-13

    @staticmethod
    def arr(x):
        return x

    @staticmethod
    def host(x):
        return np.asarray(x)

    ad = oad


def _check(be, c):
    pre = f"c{c}_"
    dims, tg = tuple(int(d) for d in G[pre + "dims"]), tuple(int(t) for t in G[pre + "targets"])
    psi, rho, u, ks = (G[pre + k] for k in ("psi", "rho", "u", "ks"))
    ad, A, H, tol = be.ad, be.arr, be.host, be.tol
    for uc, s in ((False, "_ord"), (True, "_uc")):
        assert rel(H(ad.apply_operation_vector(dims, tg, A(psi), A(u), uc)), G[pre + "apply_vec" + s]) < tol
        assert rel(H(ad.apply_operation_matrix(dims, tg, A(rho), A(u), uc)), G[pre + "apply_mat" + s]) < tol
        assert rel(H(ad.apply_kraus_matrix(dims, tg, A(rho), A(ks), uc)), G[pre + "kraus" + s]) < tol
        assert rel(H(ad.trace_out_matrix(dims, tg, A(rho), uc)), G[pre + "trace" + s]) < tol
    for seed in SEEDS:
        key = orng.PRNGKey(seed)
        o, post, probs, _ = ad.measure_vector_with_probs(dims, tg, A(psi), key)
        assert int(o) == int(G[pre + f"mv{seed}_o"])
        assert rel(H(post), G[pre + f"mv{seed}_post"]) < tol * 10 and rel(H(probs), G[pre + "mv_probs"]) < tol
        o, post, probs, _ = ad.measure_matrix_with_probs(dims, tg, A(rho), key)
        assert int(o) == int(G[pre + f"mm{seed}_o"])
        assert rel(H(post), G[pre + f"mm{seed}_post"]) < tol * 10 and rel(H(probs), G[pre + "mm_probs"]) < tol
        if pre + f"povm{seed}_o" in G:
            o, post, _ = ad.measure_povm_matrix(dims, tg, A(ks), A(rho), key)
            assert int(o) == int(G[pre + f"povm{seed}_o"])
            assert rel(H(post), G[pre + f"povm{seed}_post"]) < tol * 10
    if pre + "ev_probs" in G:
        p, red = ad.measure_vector_expectation(dims, tg, A(psi))
        assert rel(H(p), G[pre + "ev_probs"]) < tol and rel(H(red), G[pre + "ev_post"]) < tol
        p, red = ad.measure_matrix_expectation(dims, tg, A(rho))
        assert rel(H(p), G[pre + "em_probs"]) < tol and rel(H(red), G[pre + "em_post"]) < tol


@pytest.mark.parametrize("c", range(NC))
def test_oracle_matches_reference_outputs(c):
    _check(Oracle, c)


@pytest.mark.gpu
@pytest.mark.parametrize("c", range(NC))
def test_cuda_path_matches_reference_outputs(c):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import __graft_entry__ as g

    g.build()
    from photon_weave_b200.core import adapters

    class Cuda:
        tol = 1e-12
        ad = adapters

        @staticmethod
        def arr(x):
            return torch.as_tensor(np.ascontiguousarray(x)).cuda()

        @staticmethod
        def host(x):
            return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)

    _check(Cuda, c)

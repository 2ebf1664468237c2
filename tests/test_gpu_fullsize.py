"""
BASELINE.json full sizes: configs[4] 12-mode cutoff-6 state vector (34.8 GB) and configs[3]
5-mode cutoff-8 density matrix (17.2 GB).  The oracle cannot hold them, so

* after EVERY operation, randomly sampled fibers (vector) / fiber-pair blocks (density matrix)
  of the INPUT are brought to the host, the operator (or Kraus sum) is applied to them with
  NumPy -- the arithmetic of core/kernels.py:62-98, 101-145, 148-186 on that fiber -- and compared
  with the same entries of the CUDA OUTPUT (<= 1e-12 relative): a unitary-but-wrong kernel,
  a wrong permutation or an index overflow beyond 2^32 bytes fails here;
* size-independent properties on top: norm / trace preservation, inverse round trip,
  Hermiticity, agreement of the matrix and vector pictures.
"""

import math

import numpy as np
import pytest

from oracle import operators as ops

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ffi():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    free, total = torch.cuda.mem_get_info()
    if total < 150 * 2**30:
        pytest.skip("needs a 180 GB device")
    import __graft_entry__ as g

    g.build()
    from photon_weave_b200.core import ffi as f

    torch.cuda.empty_cache()
    return f


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x)).cuda()


def max_abs_diff(a, b, chunks=12):
    n = a.numel()
    step = (n + chunks - 1) // chunks
    worst = 0.0
    for i in range(0, n, step):
        worst = max(worst, float((a[i:i + step] - b[i:i + step]).abs().max()))
    return worst


def _strides(dims):
    st, acc = [], 1
    for d in reversed(dims):
        st.append(acc)
        acc *= d
    return list(reversed(st))


def _target_offsets(dims, tg):
    """Element offsets of the operator index j (targets in the GIVEN order, adapters.py:53)."""
    st = _strides(dims)
    off = np.zeros(1, dtype=np.int64)
    for a in tg:
        off = (off[:, None] + np.arange(dims[a], dtype=np.int64)[None, :] * st[a]).reshape(-1)
    return off


def _fiber_bases(dims, tg, rng, count):
    """Random fiber origins (target digits zero), half of them from the last 1 % of the array."""
    st = np.array(_strides(dims), dtype=np.int64)
    idx = rng.integers(0, np.array(dims), size=(count, len(dims)))
    idx[count // 2:, 0] = dims[0] - 1
    idx[count // 2:, 1] = dims[1] - 1
    idx[:, list(tg)] = 0
    return idx @ st


def check_vec_fibers(x_in, x_out, op, dims, tg, rng, count=64, tol=1e-12):
    """out[fiber] == op @ in[fiber] on `count` sampled fibers (NumPy on the sampled INPUT)."""
    pos = _fiber_bases(dims, tg, rng, count)[:, None] + _target_offsets(dims, tg)[None, :]
    pos_d = torch.as_tensor(pos, device=x_in.device)
    xin = x_in.reshape(-1)[pos_d].cpu().numpy()
    got = x_out.reshape(-1)[pos_d].cpu().numpy()
    want = xin @ np.asarray(op).T
    err = np.abs(got - want).max() / max(np.abs(want).max(), 1e-300)
    assert err < tol, (tg, err)


def check_mat_blocks(r_in, r_out, kraus, dims, tg, rng, count=48, tol=1e-12):
    """out[rows(I), cols(J)] == sum_k K in[rows(I), cols(J)] K^dagger on sampled fiber pairs."""
    D = math.prod(dims)
    toff = _target_offsets(dims, tg)
    rows = _fiber_bases(dims, tg, rng, count)[:, None] + toff[None, :]
    cols = _fiber_bases(dims, tg, rng, count)[:, None] + toff[None, :]
    flat = rows[:, :, None] * D + cols[:, None, :]
    flat_d = torch.as_tensor(flat, device=r_in.device)
    bin_ = r_in.reshape(-1)[flat_d].cpu().numpy()
    got = r_out.reshape(-1)[flat_d].cpu().numpy()
    ks = np.asarray(kraus)
    ks = ks[None] if ks.ndim == 2 else ks
    want = sum(np.einsum("ab,nbc,dc->nad", k, bin_, k.conj()) for k in ks)
    err = np.abs(got - want).max() / max(np.abs(want).max(), 1e-300)
    assert err < tol, (tg, err)


def test_config4_vector_full_size_properties(ffi):
    modes, d = 12, 6
    dims = (d,) * modes
    N = d**modes
    gen = torch.Generator(device="cuda").manual_seed(7)
    psi = torch.view_as_complex(torch.randn(N, 2, generator=gen, device="cuda", dtype=torch.float64))
    n0 = float(ffi.norm_stats(psi)[0])
    ffi.scale_(psi, torch.tensor([1.0 / math.sqrt(n0)], device="cuda", dtype=torch.float64))
    a, b = psi.clone(), torch.empty_like(psi)
    us = {ax: ops.random_unitary(d, 50 + ax) for ax in (0, 5, 9, 11)}
    bs = ops.beam_splitter(d, d, math.pi / 5)
    dense2 = ops.random_unitary(d * d, 77)   # dense two-mode operators: FP64 tensor-core kernel
    ph = ops.phase_operator(d, 0.41)           # diagonal operators: streaming kernel
    prog = [((0,), us[0]), ((5,), us[5]), ((9,), us[9]), ((11,), us[11]), ((9,), ph), ((0,), ph), ((11,), ph),
            ((2, 3), bs), ((10, 11), bs), ((11, 0), bs), ((3, 4), dense2), ((11, 10), dense2),
            ((4, 8), bs)]
    prog += [((8,), us[9]), ((10,), us[0]), ((8, 9), bs), ((2, 9), bs), ((6, 1), dense2)]
    rng = np.random.default_rng(0)
    for tg, op in prog:
        ffi.apply_vec(a, dev(op), dims, tg, out=b)
        check_vec_fibers(a, b, op, dims, tg, rng)      # every operation, against NumPy on its input
        a, b = b, a
    assert abs(float(ffi.norm_stats(a)[0]) - 1.0) < 1e-12
    # inverse round trip back to psi
    for tg, op in reversed(prog):
        ffi.apply_vec(a, dev(op.conj().T), dims, tg, out=b)
        a, b = b, a
    assert max_abs_diff(a, psi) < 1e-12 * float(psi.abs().max()) * 100
    del a, b, psi
    torch.cuda.empty_cache()


def test_config3_density_matrix_full_size_properties(ffi):
    modes, d = 5, 8
    dims = (d,) * modes
    D = d**modes
    # rho = sum_r |v_r><v_r| / 3 built blockwise on the device (PSD, trace ~ 1)
    gen = torch.Generator(device="cuda").manual_seed(3)
    vs = [torch.view_as_complex(torch.randn(D, 2, generator=gen, device="cuda", dtype=torch.float64))
          for _ in range(3)]
    vs = [v / torch.linalg.vector_norm(v) for v in vs]
    rho = torch.zeros((D, D), dtype=torch.complex128, device="cuda")
    step = 2048
    for i in range(0, D, step):
        for v in vs:
            rho[i:i + step] += torch.outer(v[i:i + step], v.conj()) / 3.0
    out = torch.empty_like(rho)
    tr0 = float(torch.diagonal(rho).sum().real)
    assert abs(tr0 - 1.0) < 1e-12
    loss_h = ops.loss_channel(d, 0.2)
    loss = dev(loss_h)
    rng_s = np.random.default_rng(11)
    for ax in (0, 2, 4):
        ffi.kraus_mat(rho.reshape(-1), loss, dims, (ax,), out=out.reshape(-1))
        check_mat_blocks(rho, out, loss_h, dims, (ax,), rng_s)
        assert abs(float(torch.diagonal(out).sum().real) - tr0) < 1e-11      # trace preserving
        herm = float((out[:512, :512] - out[:512, :512].conj().T).abs().max())
        assert herm < 1e-15                                                     # stays Hermitian
    # unitary on rho: purity-related invariant Tr(rho) and agreement with the vector picture:
    # (U rho U^dagger) built from the evolved vectors, checked on a 256 x 256 corner
    u = ops.random_unitary(d, 9)
    ffi.apply_mat(rho.reshape(-1), dev(u), dims, (3,), out=out.reshape(-1))
    check_mat_blocks(rho, out, u, dims, (3,), rng_s)
    # dense Kraus channel (K = 3 random contractions: dense superoperator route) and a phase
    kd = np.stack([ops.random_unitary(d, 30 + k) for k in range(3)]) / math.sqrt(3.0)
    scratch = torch.empty_like(rho)
    ffi.kraus_mat(rho.reshape(-1), dev(kd), dims, (1,), out=scratch.reshape(-1))
    check_mat_blocks(rho, scratch, kd, dims, (1,), rng_s)
    ph = ops.phase_operator(d, 0.37)
    ffi.apply_mat(rho.reshape(-1), dev(ph), dims, (4,), out=scratch.reshape(-1))
    check_mat_blocks(rho, scratch, ph, dims, (4,), rng_s)
    del scratch
    ws = [ffi.apply_vec(v, dev(u), dims, (3,)) for v in vs]
    corner = sum(torch.outer(w[:256], w[:256].conj()) for w in ws) / 3.0
    assert float((out[:256, :256] - corner).abs().max()) < 1e-15
    # the tensor-core two-sided kernels (dense on the innermost mode, beam splitter sectors packed
    # into super-blocks, row tiles for the innermost mode pair) against the vector picture on
    # 256 x 256 randomly sampled entries, then the inverse round trip of the whole sequence
    bs = ops.beam_splitter(d, d, math.pi / 5)
    seq = [((4,), u), ((1, 2), bs), ((3, 4), bs), ((4, 0), bs), ((0,), ops.random_unitary(d, 10))]
    a, b = out, torch.empty_like(out)
    rng = np.random.default_rng(5)
    I = torch.as_tensor(rng.choice(D, 256, replace=False), device="cuda")
    J = torch.as_tensor(rng.choice(D, 256, replace=False), device="cuda")
    start = a.clone()
    for tg, op in seq:
        ffi.apply_mat(a.reshape(-1), dev(op), dims, tg, out=b.reshape(-1))
        check_mat_blocks(a, b, op, dims, tg, rng_s, count=12 if len(tg) == 2 else 48)
        ws = [ffi.apply_vec(w, dev(op), dims, tg) for w in ws]
        a, b = b, a
        want = sum(torch.outer(w[I], w[J].conj()) for w in ws) / 3.0
        assert float((a[I][:, J] - want).abs().max()) < 1e-15, tg
        assert abs(float(torch.diagonal(a).sum().real) - tr0) < 1e-11, tg
    for tg, op in reversed(seq):
        ffi.apply_mat(a.reshape(-1), dev(op.conj().T), dims, tg, out=b.reshape(-1))
        a, b = b, a
    assert max_abs_diff(a.reshape(-1), start.reshape(-1)) < 1e-15
    out = a
    del b, start
    ws = [ffi.apply_vec(v, dev(u), dims, (3,)) for v in vs]
    # measurement probabilities of one mode agree between the matrix and vector pictures
    p_mat = ffi.probs(out, dims, (3,), True)
    p_vec = sum(ffi.probs(w, dims, (3,), False) * float(ffi.norm_stats(w)[0]) for w in ws) / 3.0
    assert float((p_mat - p_vec / p_vec.sum()).abs().max()) < 1e-12
    del rho, out, vs, ws
    torch.cuda.empty_cache()

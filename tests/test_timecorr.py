"""Time-correlation functions of the pyPCA chain (K7, cgc_time_corr; reference scripts/depca/depca/sidechain_corr.py:133-242).

CPU: the oracle's restatements against tests/golden/timecorr.npz — outputs of the reference's own function bodies, executed
by tests/golden/make_timecorr_golden.py — and the argument checks of the C-ABI.  GPU: the device path, called through the
shim with the scripts' names and arguments, against the same golden vectors and against the oracle on seeded series.

Tolerance (FP64 sums in a different order): |d| <= 1e-12 * sqrt(sum f1^2 * sum f2^2) / (n - tau) per lag — the Cauchy-Schwarz
bound of the entry, i.e. ~1e4 ulp of the largest value a lag can take."""
import os

import numpy as np
import pytest

from cgromcorrfmo_b200 import capi, pypca

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "timecorr.npz")
RTOL = 1e-12


def _bound(f1, f2, n_lags):
    n = f1.size
    return RTOL * np.sqrt(np.dot(f1, f1) * np.dot(f2, f2)) / (n - np.arange(n_lags)) + 1e-300


def _window(E, a, b, chromo, offset, nframes):
    E = np.asarray(E, dtype=np.float64)
    f1 = E[offset:offset + nframes, chromo - 1, a] - E[:, chromo - 1, a].mean()
    f2 = E[offset:offset + nframes, chromo - 1, b] - E[:, chromo - 1, b].mean()
    return f1, f2


def test_oracle_restatement_matches_reference_outputs(oracle_mod):
    O = oracle_mod
    g = np.load(GOLDEN)
    E = g["E"]
    for k, (ver, chromo, a, b, offset, nframes, corr_len) in enumerate(g["cases"]):
        want = g["c%d" % k]
        if ver == 1:
            got = O.store_ct_v1(E, a, b, chromo=chromo, offset=offset, nframes=nframes)
        elif ver == 2:
            got = O.store_ct_v2(E, a, b, chromo=chromo, offset=offset, nframes=nframes, corr_len=corr_len)
        else:
            got = O.store_ct_v3(E, a, b, chromo=chromo, offset=offset, nframes=nframes, corr_len=corr_len)
        assert got.shape == want.shape
        assert np.array_equal(got, want), (k, np.abs(got - want).max())      # the same numpy calls: bit for bit
    assert np.array_equal(O.time_corr(g["tc_f1"], g["tc_f2"]), g["tc"])
    assert O.time_corr(np.zeros(5), np.zeros(5)) is None
    with pytest.raises(RuntimeError, match="offset > len"):
        O.store_ct_v2(E, 0, 1, offset=E.shape[0], nframes=10, corr_len=5)
    with pytest.raises(RuntimeError, match="offset \\+ nframes > len"):
        O.store_ct_v3(E, 0, 1, offset=10, nframes=E.shape[0], corr_len=5)


def test_circular_is_linear_plus_wrapped_terms(oracle_mod):
    """the identity K7's circular mode is built on: circ[tau] = lin12[tau] + lin21[n - tau] (unnormalised sums)."""
    rng = np.random.default_rng(5)
    n, L = 257, 100
    f1, f2 = rng.normal(size=n), rng.normal(size=n)
    circ = np.real(np.fft.ifft(np.fft.fft(f2) * np.conj(np.fft.fft(f1))))[:L]
    lin12 = np.array([np.dot(f1[:n - t], f2[t:]) for t in range(L)])
    lin21 = np.array([0.0] + [np.dot(f2[:t], f1[n - t:]) for t in range(1, L)])
    assert np.allclose(circ, lin12 + lin21, rtol=0, atol=1e-11)


def test_argument_checks_need_no_device(small_inputs):
    _, top, itp, _ = small_inputs("tiny")
    s = np.zeros((2, 10))
    with capi.Context(top, itp, device=-1) as c:
        for kw, msg in ((dict(corr_len=0), "corr_len"), (dict(corr_len=11), "corr_len")):
            with pytest.raises(capi.CgcError, match=msg) as e:
                c.time_corr(s, [0], [1], **kw)
            assert e.value.code == capi.CGC_ERR_ARG
        with pytest.raises(capi.CgcError, match="pair index") as e:
            c.time_corr(s, [0], [2], 5)
        assert e.value.code == capi.CGC_ERR_ARG
        assert c.time_corr(s, [], [], 5).shape == (0, 5)
        with pytest.raises(capi.CgcError) as e:       # no CPU fallback: a table-only context cannot compute
            c.time_corr(s, [0], [1], 5)
        assert e.value.code == capi.CGC_ERR_CUDA


@pytest.fixture(scope="module")
def gpu_ctx(small_inputs):
    _, top, itp, _ = small_inputs("tiny")
    with capi.Context(top, itp, device=0) as c:
        yield c


@pytest.mark.gpu
def test_gpu_matches_reference_outputs(gpu_ctx):
    g = np.load(GOLDEN)
    E = g["E"]
    for k, (ver, chromo, a, b, offset, nframes, corr_len) in enumerate(g["cases"]):
        want = g["c%d" % k]
        f1, f2 = _window(E, a, b, chromo, offset, nframes)
        if ver == 1:
            got = pypca.HDFStoreCt_v1(gpu_ctx, E, a, b, chromo=chromo, offset=offset, nframes=nframes)
        elif ver == 2:
            got = pypca.HDFStoreCt_v2(gpu_ctx, E, a, b, chromo=chromo, offset=offset, nframes=nframes, corr_len=corr_len)
        else:
            got = pypca.HDFStoreCt_v3(gpu_ctx, E, a, b, chromo=chromo, offset=offset, nframes=nframes, corr_len=corr_len)
        assert got.shape == want.shape
        # the reference's FFT (v3) and np.correlate carry their own rounding: allow both sides the bound
        assert np.all(np.abs(got - want) <= 2 * _bound(f1, f2, want.size)), (k, np.abs(got - want).max())
    got = pypca.TimeCorr_gpu(gpu_ctx, g["tc_f1"], g["tc_f2"])
    assert np.all(np.abs(got - g["tc"]) <= 2 * _bound(g["tc_f1"], g["tc_f2"], got.size))
    assert pypca.TimeCorr_gpu(gpu_ctx, np.zeros(7), np.zeros(7)) is None
    with pytest.raises(RuntimeError, match="offset > len"):
        pypca.HDFStoreCt_v2(gpu_ctx, E, 0, 1, offset=E.shape[0], nframes=10, corr_len=5)
    with pytest.raises(RuntimeError, match="offset \\+ nframes > len"):
        pypca.HDFStoreCt_v3(gpu_ctx, E, 0, 1, offset=10, nframes=E.shape[0], corr_len=5)


@pytest.mark.gpu
@pytest.mark.parametrize("n,corr_len", [(1, 1), (7, 7), (513, 1), (1024, 1024), (1025, 1025), (5000, 3001), (20011, 2049)])
@pytest.mark.parametrize("circular", [False, True])
def test_gpu_matches_oracle_on_seeded_series(gpu_ctx, oracle_mod, n, corr_len, circular):
    """lag-block, chunk and split boundaries (1024 lags per CTA, 512 time steps per chunk), corr_len == n, single samples."""
    O = oracle_mod
    rng = np.random.default_rng(1000 * n + corr_len)
    S = 4
    E = (rng.normal(size=(n + 9, 1, S)).cumsum(axis=0) * 0.1 + rng.normal(size=(n + 9, 1, S)) + [5.0, -3.0, 0.0, 40.0])
    pairs = [(0, 0), (0, 1), (1, 0), (2, 3), (3, 3)]
    got = pypca.StoreCt_pairs_gpu(gpu_ctx, E, pairs, chromo=1, offset=4, nframes=n, corr_len=corr_len, circular=circular)
    assert got.shape == (len(pairs), corr_len)
    ref_fn = O.store_ct_v3 if circular else O.store_ct_v2
    for (a, b), row in zip(pairs, got):
        want = ref_fn(E, a, b, chromo=1, offset=4, nframes=n, corr_len=corr_len)
        f1, f2 = _window(E, a, b, 1, 4, n)
        assert np.all(np.abs(row - want) <= 2 * _bound(f1, f2, corr_len)), ((a, b), np.abs(row - want).max())


@pytest.mark.gpu
def test_gpu_deterministic_and_batched(gpu_ctx, monkeypatch):
    """twice: identical bits; and again with the partial-sum scratch capped so that the pairs take several passes."""
    rng = np.random.default_rng(77)
    s = rng.normal(size=(12, 30000))
    a, b = np.meshgrid(np.arange(12), np.arange(12))
    r1 = gpu_ctx.time_corr(s, a.ravel(), b.ravel(), 2500)
    r2 = gpu_ctx.time_corr(s, a.ravel(), b.ravel(), 2500)
    assert np.array_equal(r1, r2)
    monkeypatch.setenv("CGC_TIMECORR_SCRATCH_DOUBLES", str(40 * 3072))     # a few pairs per pass
    r3 = gpu_ctx.time_corr(s, a.ravel(), b.ravel(), 2500, circular=True)
    monkeypatch.delenv("CGC_TIMECORR_SCRATCH_DOUBLES")
    r4 = gpu_ctx.time_corr(s, a.ravel(), b.ravel(), 2500, circular=True)
    assert np.allclose(r3, r4, rtol=0, atol=1e-15 * 30000)                # (the split count may differ: not bit-identical)
    # C_ab[tau = 0] is symmetric, and a pair's row does not depend on what else is in the batch
    assert np.allclose(r1.reshape(12, 12, -1)[:, :, 0], r1.reshape(12, 12, -1)[:, :, 0].T, rtol=1e-13)
    one = gpu_ctx.time_corr(s, [3], [7], 2500)
    k = int(np.flatnonzero((a.ravel() == 3) & (b.ravel() == 7))[0])
    assert np.allclose(one[0], r1[k], rtol=0, atol=1e-15 * 30000)


@pytest.mark.gpu
def test_gpu_full_size_against_padded_fft(gpu_ctx):
    """the scripts' default size (nframes 200 000, corr_len 20 000): the linear sums against a zero-padded FP64 FFT, the
    circular ones against the unpadded FFT of HDFStoreCt_v3."""
    rng = np.random.default_rng(2014)
    n, L = 200000, 20000
    kern = np.exp(-np.arange(64) / 11.0)
    s = np.stack([np.convolve(rng.normal(size=n + 63), kern, "valid"), np.convolve(rng.normal(size=n + 63), kern, "valid")])
    s -= s.mean(axis=1, keepdims=True)
    lin, ms = gpu_ctx.time_corr(s, [0, 0], [1, 0], L, want_ms=True)
    F = np.fft.rfft(s, 2 * n, axis=1)
    for row, (a, b) in zip(lin, [(0, 1), (0, 0)]):
        want = np.fft.irfft(F[b] * np.conj(F[a]), 2 * n)[:L] / (n - np.arange(L))
        assert np.all(np.abs(row - want) <= 10 * _bound(s[a], s[b], L)), np.abs(row - want).max()
    circ = gpu_ctx.time_corr(s, [0], [1], L, circular=True)[0]
    want = np.real(np.fft.ifft(np.fft.fft(s[1]) * np.conj(np.fft.fft(s[0])))[:L]) / (n - np.arange(L))
    assert np.all(np.abs(circ - want) <= 10 * _bound(s[0], s[1], L)), np.abs(circ - want).max()
    assert ms > 0

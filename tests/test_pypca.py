"""The pyPCA/depca shim (cgromcorrfmo_b200/pypca.py): host-side post-processing against a line-by-line restatement of the
reference scripts, and (gpu) the device-computed reduction against numpy on the rows the same run returns."""
import numpy as np
import pytest

from cgromcorrfmo_b200 import pypca


def _compute_modes_reference(corr):
    """reference scripts/depca/depca/sidechain_corr.py:91-129, restated without the py2 idioms."""
    vi, wi, imp = [], [], []
    for c in corr:
        v, w = np.linalg.eigh(c)
        w = w.transpose().copy()
        n = []
        for k in range(len(v)):
            nf = sum(w[k])
            if (v[k] < 0 and nf > 0) or (v[k] > 0 and nf < 0):
                w[k] *= -1
                nf *= -1
            n.append(nf)
        tup = sorted(zip(v, n, [x.copy() for x in w]), key=lambda x: x[0] * x[1] * x[1])
        vi.append([t[0] for t in tup]); imp.append([t[1] for t in tup]); wi.append([t[2] for t in tup])
    return np.array(vi), np.array(wi), np.array(imp)


def test_host_post_processing_matches_depca():
    rng = np.random.default_rng(90210)                     # the seed of the reference's own script (old/test__pyPCA.py:9)
    E = (rng.normal(size=(1000, 2, 6)) @ np.diag([3, 2, 1, 0.5, 0.2, 0.1]) + 7.0).astype(np.float32)
    corr, avg = pypca.AvgAndCorrelateSidechains(E)
    for i in range(2):
        assert np.allclose(corr[i], np.cov(E[:, i, :].astype(np.float64), rowvar=False))
    vi, wi, imp = pypca.ComputeModes(corr)
    rv, rw, ri = _compute_modes_reference(corr)
    assert np.allclose(vi, rv) and np.allclose(wi, rw) and np.allclose(imp, ri)
    assert np.all(np.diff(vi * imp * imp, axis=1) >= -1e-12)
    proj = pypca.ApplyPCA(E, avg, wi)
    assert proj.shape == (1000, 2, 6)
    for i in range(2):
        assert np.allclose(np.cov(proj[:, i, :], rowvar=False), np.diag(vi[i]), atol=1e-8)   # modes diagonalise the covariance


@pytest.mark.gpu
def test_gpu_reduction_matches_numpy(small_inputs):
    traj, top, itp, frames = small_inputs("tiny")
    E = pypca.load_E_t_ia(traj, top, itp, frames, sites=7)
    assert E.shape[:2] == (frames, 7) and E.dtype == np.float32
    Ek = pypca.load_E_t_ia(traj, top, itp, frames, sites=7, units="kcal")
    assert np.array_equal(E, (Ek.astype(np.float64) * 349.7).astype(np.float32))
    corr, avg = pypca.AvgAndCorrelateSidechains_gpu(traj, top, itp, frames, sites=7, units="kcal")
    rc, ra = pypca.AvgAndCorrelateSidechains(Ek)
    scale = np.maximum(np.abs(rc), np.sqrt(np.einsum("sii,sjj->sij", rc, rc)))
    ok = scale > 0
    assert np.all(np.abs(corr - rc)[ok] <= 1e-9 * scale[ok])
    assert np.allclose(avg, ra, rtol=1e-12, atol=1e-300)


@pytest.mark.gpu
@pytest.mark.parametrize("n_modes", [1, 5, None])
def test_gpu_projection_matches_numpy(n_modes, small_inputs):
    """K4 (cgc_project) against depca's ApplyPCA restated in numpy float64, on the rows and modes of a real run and on
    frame counts that are not a multiple of the 64-frame tile."""
    from cgromcorrfmo_b200 import capi
    traj, top, itp, frames = small_inputs("permol")
    with capi.Context(top, itp, 0) as c:
        c.open(traj)
        E, done, _ = c.run(0, frames)
        S, G = c.n_sites, c.num_tot
        rng = np.random.default_rng(11)
        # a longer series with the statistics of the real rows (the run itself is only a few frames long)
        T = 333
        X = (E.mean(axis=0)[None] + rng.normal(size=(T, S, G)) * (np.abs(E).mean() + 1e-3)).astype(np.float32)
        corr, avg = pypca.AvgAndCorrelateSidechains(X)
        vi, wi, imp = pypca.ComputeModes(corr)
        W = wi if n_modes is None else wi[:, -n_modes:, :]
        ref = pypca.ApplyPCA(X, avg, W)
        got = pypca.ApplyPCA_gpu(c, X, avg, W)
        assert got.shape == ref.shape == (T, S, W.shape[1])
        scale = np.einsum("tsg,smg->tsm", np.abs(X.astype(np.float64) - avg[None]), np.abs(W))
        assert np.all(np.abs(got - ref) <= 1e-13 * scale + 1e-300)
        with pytest.raises(ValueError):
            c.project(X[:, :1], avg, W)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 7, 64, 257])
def test_gpu_eigh_matches_lapack(n, small_inputs):
    """K6 (cgc_eigh, batched one-sided Jacobi) against numpy.linalg.eigh: eigenvalues, orthonormal eigenvectors, and
    A = V^T diag(lambda) V — on covariance-like matrices (positive semi-definite, one exactly zero row/column like the
    own-site column of a dE covariance), an indefinite one and a rank-deficient one (fewer frames than columns)."""
    from cgromcorrfmo_b200 import capi
    traj, top, itp, frames = small_inputs("tiny")
    rng = np.random.default_rng(n)
    mats = []
    X = rng.normal(size=(4 * n + 3, n)) * (1.0 + 10.0 * rng.random(n))
    c = np.cov(X, rowvar=False).reshape(n, n)
    if n > 2:
        c[0, :] = 0.0
        c[:, 0] = 0.0
    mats.append(c)
    b = rng.normal(size=(n, n))
    mats.append((b + b.T) / 2)                                         # indefinite
    Y = rng.normal(size=(max(1, n // 3), n))
    mats.append(Y.T @ Y)                                               # rank-deficient: a large null space
    A = np.stack(mats)
    with capi.Context(top, itp, 0) as ctx:
        ctx.open(traj)
        ev, vt, sweeps = ctx.eigh(A)
    assert 1 <= sweeps <= 40
    for k in range(len(mats)):
        ref = np.linalg.eigvalsh(A[k])
        scale = max(np.abs(ref).max(), 1e-300)
        assert np.all(np.diff(ev[k]) >= 0)
        assert np.all(np.abs(ev[k] - ref) <= 1e-12 * scale)
        assert np.allclose(vt[k] @ vt[k].T, np.eye(n), atol=1e-12)
        assert np.allclose(vt[k].T @ np.diag(ev[k]) @ vt[k], A[k], atol=1e-11 * scale)


@pytest.mark.gpu
def test_compute_modes_gpu_matches_numpy(small_inputs):
    from cgromcorrfmo_b200 import capi
    traj, top, itp, frames = small_inputs("tiny")
    rng = np.random.default_rng(4)
    E = (rng.normal(size=(400, 3, 12)) @ np.diag(np.linspace(3, 0.2, 12)) + 5.0).astype(np.float32)
    corr, avg = pypca.AvgAndCorrelateSidechains(E)
    vi, wi, imp = pypca.ComputeModes(corr)
    with capi.Context(top, itp, 0) as ctx:
        ctx.open(traj)
        gv, gw, gi = pypca.ComputeModes_gpu(ctx, corr)
    assert np.allclose(gv, vi, rtol=1e-10) and np.allclose(gi, imp, atol=1e-8) and np.allclose(gw, wi, atol=1e-8)


@pytest.mark.gpu
def test_whole_pypca_chain_on_the_device(tmp_path):
    """trajectory -> E_t_ia -> per-site covariance (K3) -> modes (K6) -> projections (K4) -> time correlation of mode
    amplitudes (K7), every step against the numpy formulation of the depca scripts on the same rows."""
    from cgromcorrfmo_b200 import capi, synth
    s = synth.make_system(1200, 7, 11, "shared")
    frames = 48
    traj, top, itp = synth.write_inputs(str(tmp_path), s, frames)
    E = pypca.load_E_t_ia(traj, top, itp, frames, sites=7)                     # float32 (T, 7, Ncoarse), cm^-1
    corr, avg = pypca.AvgAndCorrelateSidechains_gpu(traj, top, itp, frames, sites=7)
    rc, ra = pypca.AvgAndCorrelateSidechains(E)
    # the device reduces the kcal/mol rows and scales by 349.7^2; numpy reduces the rounded cm^-1 rows: float32 rounding apart
    assert np.allclose(corr, rc, rtol=2e-5, atol=1e-6 * np.abs(rc).max()) and np.allclose(avg, ra, rtol=1e-6, atol=1e-6)
    with capi.Context(top, itp, 0) as ctx:
        ctx.open(traj)
        ctx.set_option("sites", 7)
        vi, wi, imp = pypca.ComputeModes_gpu(ctx, rc)
        rv, rw, ri = pypca.ComputeModes(rc)
        assert np.allclose(vi, rv, rtol=0, atol=1e-10 * np.abs(rv).max())
        proj = pypca.ApplyPCA_gpu(ctx, E, ra, wi)
        assert np.allclose(proj, pypca.ApplyPCA(E, ra, wi), rtol=0, atol=1e-9 * np.abs(proj).max())
        # modes diagonalise the covariance of the rows they were computed from (rank <= frames - 1: most eigenvalues are 0)
        for i in range(7):
            c = np.cov(proj[:, i, :], rowvar=False)
            assert np.allclose(np.diag(c), vi[i], atol=1e-8 * np.abs(vi[i]).max())
        # time correlation of the two strongest modes of chromophore 1 (sorted last), the scripts' v3 and v2 forms
        M = proj.shape[2]
        for fn, circ in ((pypca.HDFStoreCt_v3, True), (pypca.HDFStoreCt_v2, False)):
            got = fn(ctx, proj, M - 1, M - 2, chromo=1, offset=0, nframes=frames, corr_len=16)
            f1 = proj[:, 0, M - 1] - proj[:, 0, M - 1].mean()
            f2 = proj[:, 0, M - 2] - proj[:, 0, M - 2].mean()
            if circ:
                want = np.real(np.fft.ifft(np.fft.fft(f2) * np.conj(np.fft.fft(f1)))[:16]) / (frames - np.arange(16))
            else:
                want = np.array([np.inner(f1[:frames - t], f2[t:]) / (frames - t) for t in range(16)])
            assert np.allclose(got, want, rtol=0, atol=1e-11 * np.sqrt(np.dot(f1, f1) * np.dot(f2, f2)))

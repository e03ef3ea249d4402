"""Full-size parity on the GPU (-m gpu): BASELINE.json's shapes, checked per entry.

  C2  one frame of the 100 000-atom trimer, all 24 sites x 1099 columns against the float64 restatement
      (oracle_cdc_f64, |d| <= 1e-6 * max(|dE|, sum |term|)) AND against the float32 rows of the UNMODIFIED reference library
      for the same inputs (tests/golden/c2_frame0_ref.npz, made by tests/golden/make_full_size_golden.py)
  C1  the same for the 30 000-atom monomer, 7 sites
  C4  350 000 atoms, every solvent molecule its own group (~1e5 columns: all bins through the atomic path), 3 sites
  C5  K3 at G = 1099 over several launches against np.cov, both ddof, with a look at the edge of the last 64-block

What these replace: ComputeCDC_v1, reference src/grom2atomFMO.cpp:625-693 (rows), the Correlate block
src/FMOparse.cpp:168-206 and depca's np.cov, scripts/depca/depca/sidechain_corr.py:50-72 (covariance).
"""
import os

import numpy as np
import pytest

from conftest import sha

pytestmark = pytest.mark.gpu

DE_TOL = 1e-6
COV_TOL = 1e-9
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def capi_mod():
    from cgromcorrfmo_b200 import capi
    capi.load()
    return capi


def _write_named(tmp, name, frames=1):
    from cgromcorrfmo_b200 import synth
    s = synth.named_system(name)
    top, itp, traj = str(tmp / "topol.top"), str(tmp / "bcx.itp"), str(tmp / "traj.gro")
    synth.write_top(top, s)
    synth.write_itp(itp, s)
    prefix = synth._static_prefix(s)
    with open(traj, "wb") as f:
        for t in range(frames):
            f.write(synth.frame_bytes(s, prefix, t))
    return s, traj, top, itp


def _exact_rows(O, traj, top, itp, sites):
    """float64 restatement of ComputeCDC_v1 fed the float32 inputs the reference reads: (rows, scale) per site."""
    names, masses, charges = O.load_tables(top, itp)
    index = {}
    for j, n in enumerate(names):
        index.setdefault(n, j)
    _, Xi, keys, grp, _ = O.read_frame(traj, None)
    out = {}
    for site in sites:
        q = O.atom_data_lookup(keys, grp, site, names, masses, charges, index)[2]
        out[site] = O.compute_cdc_f64(site, Xi, q, grp)
    return grp, out


@pytest.mark.parametrize("name", ["C1", "C2"])
def test_full_frame_every_entry(name, tmp_path, oracle_mod, capi_mod):
    O, capi = oracle_mod, capi_mod
    s, traj, top, itp = _write_named(tmp_path, name)
    with capi.Context(top, itp, 0) as c:
        c.open(traj)
        n_atoms, n_chromo, n_groups, G = c.dims()
        dE, done, rc = c.run(0, 1)
        groups = c.groups()
    assert rc == capi.CGC_OK and done == 1 and dE.shape == (1, n_chromo, G)
    grp, exact = _exact_rows(O, traj, top, itp, range(1, n_chromo + 1))
    assert np.array_equal(groups, grp)                                   # bit-exact group ids at full size
    worst = 0.0
    for site, (r64, sc) in exact.items():
        ours = dE[0, site - 1].astype(np.float64)
        scale = np.maximum(np.abs(r64), sc)
        m = scale > 0
        worst = max(worst, float((np.abs(ours - r64)[m] / scale[m]).max()))
        assert np.all(dE[0, site - 1][~m] == 0)                          # own-site column and empty columns: exactly 0
        assert abs(ours.sum() - r64.sum()) <= DE_TOL * max(abs(r64.sum()), sc.sum())
    assert worst <= DE_TOL, "max scaled error %g" % worst
    # the unmodified reference's own float32 rows for the same inputs
    g = np.load(os.path.join(GOLD, "%s_frame0_ref.npz" % name.lower()))
    assert [sha(traj), sha(top), sha(itp)] == list(g["sha"]), "the synthetic generator no longer reproduces the golden inputs"
    assert np.array_equal(groups, g["groups"])
    assert list(g["dims"]) == [n_atoms, n_chromo, n_groups, G]
    ref32 = g["dE_f32"].astype(np.float64)
    worst_ref = worst_ours_vs_ref = 0.0
    for site, (r64, sc) in exact.items():
        scale = np.maximum(np.abs(r64), sc)
        m = scale > 0
        worst_ref = max(worst_ref, float((np.abs(ref32[site - 1] - r64)[m] / scale[m]).max()))
        worst_ours_vs_ref = max(worst_ours_vs_ref, float((np.abs(dE[0, site - 1] - ref32[site - 1])[m] / scale[m]).max()))
    assert worst_ours_vs_ref <= DE_TOL                                   # within tolerance of the reference itself
    assert worst <= max(2 * worst_ref, 2e-7)                             # and no noisier than its float32 arithmetic
    print("%s: max scaled error ours %.3g, reference float32 %.3g, ours vs reference %.3g" % (name, worst, worst_ref, worst_ours_vs_ref))


def test_c4_shape_per_molecule_groups(tmp_path, oracle_mod, capi_mod):
    """Membrane-sized system, every solvent molecule its own group: ~1e5 columns, every bin through the in-thread-run +
    atomic path, covariance off (24 x (1e5)^2 sums do not fit).  Three sites per entry, the rest by their totals."""
    O, capi = oracle_mod, capi_mod
    s, traj, top, itp = _write_named(tmp_path, "C4", frames=2)
    with capi.Context(top, itp, 0) as c:
        c.open(traj)
        n_atoms, n_chromo, n_groups, G = c.dims()
        assert n_atoms == 350000 and n_chromo == 24 and G > 50000
        assert c.cov_size() == 0
        dE, done, rc = c.run(0, 2)
        dE2, _, _ = c.run(0, 2)
        groups = c.groups()
    assert done == 2
    assert np.array_equal(dE.view(np.uint32), dE2.view(np.uint32))       # bit-identical from run to run
    grp, exact = _exact_rows(O, traj, top, itp, (1, 12, 24))
    assert np.array_equal(groups, grp)
    for site, (r64, sc) in exact.items():
        ours = dE[0, site - 1].astype(np.float64)
        scale = np.maximum(np.abs(r64), sc)
        m = scale > 0
        e = float((np.abs(ours - r64)[m] / scale[m]).max())
        assert e <= DE_TOL, "site %d: max scaled error %g" % (site, e)
        assert np.all(dE[0, site - 1][~m] == 0)


def test_c5_covariance_at_full_width(tmp_path, capi_mod):
    """K3 on the C2/C5 width (24 sites x 1099 columns: seventeen full 64-blocks and one of 11 columns) over 5 launches of
    uneven length against np.cov in float64 on the identical float32 rows: 1e-9 of max(|C_ij|, sqrt(C_ii C_jj)), ddof 0 and 1."""
    import torch
    capi = capi_mod
    s, traj, top, itp = _write_named(tmp_path, "C2")
    rng = np.random.default_rng(5)
    with capi.Context(top, itp, 0) as c:
        c.open(traj)
        S, G = c.n_sites, c.num_tot
        assert (S, G) == (24, 1099)
        T = 1000
        # correlated columns with means far larger than the fluctuations (what the common shift is for)
        base = rng.normal(size=(T, S, 8)).astype(np.float32)
        mix = rng.normal(size=(S, 8, G)).astype(np.float32)
        X = (np.einsum("tsk,skg->tsg", base, mix) * 0.02 + rng.normal(size=(T, S, G)).astype(np.float32) * 0.01
             + rng.normal(size=(1, S, G)).astype(np.float32) * 30.0).astype(np.float32)
        d = torch.from_numpy(X).cuda()
        c.cov_reset()
        t0 = 0
        for n in (256, 256, 17, 300, 171):
            c.cov_accumulate_device(d[t0:t0 + n].contiguous().data_ptr(), n)
            t0 += n
        assert t0 == T
        torch.cuda.synchronize()
        res = {ddof: c.cov_finalize(ddof) for ddof in (0, 1)}
        assert c.cov_partials()[0] == T
    X64 = X.astype(np.float64)
    for ddof, (mean, cov) in res.items():
        assert np.array_equal(cov, np.swapaxes(cov, 1, 2))
        for k in range(S):
            ref = np.cov(X64[:, k, :], rowvar=False, ddof=ddof)
            sd = np.sqrt(np.diag(ref))
            scale = np.maximum(np.abs(ref), np.outer(sd, sd))
            err = np.abs(cov[k] - ref) / scale
            assert err.max() <= COV_TOL, "site %d ddof %d: %g" % (k, ddof, err.max())
            assert err[1088:, :].max() <= COV_TOL and err[:, 1088:].max() <= COV_TOL   # the 11-column edge block
            assert np.allclose(mean[k], X64[:, k, :].mean(0), rtol=1e-12)

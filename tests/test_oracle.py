"""CPU tests: the oracle (oracle/) against the reference's golden vectors and, where oracle/_ref was built from the
unmodified reference sources, against the reference itself.  Nothing here touches the product path."""
import os

import numpy as np
import pytest

from conftest import scaled_err, sha

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_reference_unit_test_lookup(oracle_mod):
    """The three assertions of the reference's only unit test with content
    (reference src/_test_grom2atomFMO.cpp:24-71, AtomDataLookup_v2)."""
    O = oracle_mod
    names = ["BCL,MG", "BCL,MG,EXC"]
    masses = np.array([33.0, 0.2], np.float32)
    charges = np.array([0.20, 0.70], np.float32)
    keys, groups = ["BCL,MG"], np.array([-1], np.int32)
    m, sz, q, ground = O.atom_data_lookup(keys, groups, 0, names, masses, charges)      # ground-state lookup (:49-54)
    assert m[0] == pytest.approx(33.0) and q[0] == pytest.approx(0.20)
    m, sz, q, ground = O.atom_data_lookup(keys, groups, 1, names, masses, charges)      # excited lookup (:56-62)
    assert q[0] == pytest.approx(0.70) and m[0] == pytest.approx(0.2) and ground[0] == pytest.approx(0.20)
    n_groups = max(int(groups.max()), 0)                                                # group counting (:64-71)
    assert n_groups == 0 and abs(int(groups.min())) == 1
    with pytest.raises(O.InputFileMalformed):
        O.atom_data_lookup(["BCL,XX"], groups, 0, names, masses, charges)


def test_oracle_matches_golden_synthetic(oracle_mod, small_inputs):
    """golden7.npz was produced by the unmodified reference (tests/golden/make_golden.py)."""
    O = oracle_mod
    g = np.load(os.path.join(GOLD, "golden7.npz"))
    traj, top, itp, frames = small_inputs("golden7")
    if [sha(traj), sha(top), sha(itp)] != list(g["sha"]):
        pytest.skip("synthetic generator produced different bytes than when the golden file was made")
    r = O.run(traj, top, itp, frames)
    assert np.array_equal(r["groups"], g["groups"])                       # bit-exact group ids
    assert r["keys"] == list(g["keys"])
    assert np.array_equal(r["q"].view(np.uint32), g["q"].view(np.uint32))
    assert np.array_equal(r["x0"].view(np.uint32), g["x0"].view(np.uint32))
    assert np.array_equal(r["dE_f32"].view(np.uint32), g["dE_f32"].view(np.uint32))   # float32 path: BIT-exact
    e, z = scaled_err(g["dE_f32"], r["dE_f64"], r["scale"])
    assert e < 1e-6 and z == 0.0                                          # f64 restatement agrees with the reference
    assert O.format_time_rows(r["dE_f32"][:, :7].reshape(-1, r["dE_f32"].shape[-1])) == str(g["time_text"])
    assert O.format_mtx(O.mtx_compat_f32(r["dE_f32"][:, :7])) == str(g["mtx_text"])


def test_oracle_matches_golden_reference_fixture(oracle_mod, ref_fixture):
    O = oracle_mod
    g = np.load(os.path.join(GOLD, "fixture.npz"))
    traj, top, itp = ref_fixture
    assert [sha(traj), sha(top), sha(itp)] == list(g["sha"])
    r = O.run(traj, top, itp, 2)
    assert r["n_chromo"] == 7 and r["n_groups"] == 350
    assert np.array_equal(r["groups"], g["groups"])
    assert np.array_equal(r["dE_f32"].view(np.uint32), g["dE_f32"].view(np.uint32))
    txt = O.format_time_rows(r["dE_f32"].reshape(-1, 357))
    assert txt == str(g["time_text"])
    # the survey's smoke values: per-site totals of frame 0 in cm^-1
    tot = O.kcal_to_cm(r["dE_f32"][0]).astype(np.float64).sum(-1)
    assert np.allclose(tot, [383.37, 132.194, -175.118, -276.294, -347.642, -104.302, -76.3561], rtol=6e-6)  # printed with 6 digits


def test_oracle_against_live_reference(oracle_mod, small_inputs, tmp_path):
    """Bit-exact against the unmodified reference binaries, when they were built here."""
    O = oracle_mod
    if not O.have_ref():
        pytest.skip("oracle/_ref not built")
    for tag in ("tiny", "permol", "nosolv"):
        traj, top, itp, frames = small_inputs(tag)
        r = O.run(traj, top, itp, frames)
        h = O.run_ref_harness(traj, top, itp, frames, str(tmp_path / ("h_" + tag)))
        assert h["frames"] == r["frames"] == frames
        assert np.array_equal(r["groups"], h["groups"]) and r["keys"] == h["keys"]
        assert np.array_equal(r["q"].view(np.uint32), h["q"].view(np.uint32))
        assert np.array_equal(r["dE_f32"].view(np.uint32), h["dE_f32"].view(np.uint32))
        e, z = scaled_err(h["dE_f32"], r["dE_f64"], r["scale"])
        assert e < 1e-6


def test_oracle_cov_matches_numpy(oracle_mod):
    O = oracle_mod
    rng = np.random.default_rng(5)
    X = (rng.normal(size=(40, 3, 17)) + 5.0).astype(np.float32)
    for ddof in (0, 1):
        mean, cov = O.cov_f64(X, ddof)
        for k in range(3):
            ref = np.cov(X[:, k, :].astype(np.float64), rowvar=False, ddof=ddof)
            assert np.allclose(cov[k], ref, rtol=1e-12, atol=1e-15)
            assert np.allclose(mean[k], X[:, k, :].astype(np.float64).mean(0), rtol=1e-14)


def test_oracle_mtx_compat_is_minus_mean_products(oracle_mod):
    """The reference's .mtx is ~ -<x_i><x_j>, not a covariance (reference src/FMOparse.cpp:42,193,202)."""
    O = oracle_mod
    rng = np.random.default_rng(6)
    X = (rng.normal(size=(10, 2, 9)) + 3.0).astype(np.float32)
    m = O.mtx_compat_f32(X)
    mean = X.astype(np.float64).mean(0)
    assert np.allclose(m, -np.einsum("ki,kj->kij", mean, mean), rtol=1e-5, atol=1e-6)


def test_oracle_parse_errors(oracle_mod, tmp_path):
    O = oracle_mod
    bad = tmp_path / "bad.itp"
    bad.write_text("[ atoms ]\n 1 T 1 BCX MG 1 0.7 24.3 ;\n 2 T 1 BCX CHA 2 ;\n")     # 5 columns: neither 8 nor 11
    with pytest.raises(O.InputFileMalformed):
        O.parse_topology(str(bad), [], [], [], True)
    ws = tmp_path / "ws.itp"
    ws.write_text("[ atoms ]\n 1 T 1 BCX MG 1 0.7 24.3 ;\n   \n")                       # whitespace-only line throws
    with pytest.raises(O.InputFileMalformed):
        O.parse_topology(str(ws), [], [], [], True)
    mass = tmp_path / "mass.itp"
    mass.write_text("[ atoms ]\n 1 T 1 BCX MG 1 0.7 24.3 T 0.6 25.0 ;\n")               # mB != mA
    with pytest.raises(O.InputFileMalformed):
        O.parse_topology(str(mass), [], [], [], True)
    dup = tmp_path / "dup.top"
    dup.write_text("[ atoms ]\n 1 X 1 ALA N 1 0.1 14.0\n 1 X 1 ALA N 1 0.2 14.0\n")     # conflicting duplicate: Minor
    names, masses, charges = [], [], []
    with pytest.raises(O.InputFileMinorMalformed):
        O.parse_topology(str(dup), names, masses, charges, False)
    assert names == ["ALA,N1"] and charges[0] == np.float32(0.1)                         # first entry wins
    with pytest.raises(O.InputFileMalformed):
        O.parse_topology(str(tmp_path / "missing.top"), [], [], [], False)
    gro = tmp_path / "x.gro"
    gro.write_text("stray line\n")
    with pytest.raises(O.ReadError):
        O.read_frame(str(gro), None)
    gro.write_text("title t= 0.0\nnot-a-number\n")
    with pytest.raises(O.InputFileMalformed):
        O.read_frame(str(gro), None)


def test_division_free_decode_is_exact(oracle_mod):
    """K1 converts "%8.3f" text as N * r + (N - q * 1000) * r (two FMAs) instead of dividing: bit-identical to the IEEE
    quotient for EVERY integer N < 2^24 (a "%8.3f" field holds at most 7 digits: N < 10^7), and to the reference's
    (float)atof(text) (reference src/grom2atomFMO.cpp:448-450)."""
    L = oracle_mod.lib()
    assert L.cgo_div1000_mismatches(0, 1 << 24) == 0
    assert L.cgo_div1000_vs_atof_mismatches(0, 10_000_000, 7) == 0

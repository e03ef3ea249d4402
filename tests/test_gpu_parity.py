"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C-ABI (cgromcorrfmo_b200/capi.py ->
libcgromcorr.so); the oracle (oracle/) is only the checker.

Tolerances (BASELINE.json north_star, made precise in SURVEY.md section 8c):
  * group ids, columns, frame counts, decoded coordinates: bit-exact;
  * dE: |dE_gpu - dE_exact| <= 1e-6 * max(|dE_exact|, sum over pairs of |term|) per entry, dE_exact = the float64
    restatement fed the same float32 inputs; per-site totals within 1e-6 of the same scale;
  * covariance: 1e-9 relative to max(|C_ij|, sqrt(C_ii C_jj)) against float64 on the identical dE matrix.
"""
import os

import numpy as np
import pytest

from conftest import scaled_err, sha

pytestmark = pytest.mark.gpu

DE_TOL = 1e-6
COV_TOL = 1e-9
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "the gpu tests need a CUDA device"
    return torch


@pytest.fixture(scope="module")
def capi_mod():
    from cgromcorrfmo_b200 import capi
    capi.load()
    return capi


def _device_text(torch, path):
    text = np.fromfile(path, dtype=np.uint8)
    padded = np.concatenate([text, np.zeros(((-text.size) % 16) + 16, np.uint8)])
    return torch.from_numpy(padded).cuda(), text.size


@pytest.mark.parametrize("tag", ["tiny", "permol", "nosolv"])
def test_dE_against_oracle(tag, small_inputs, oracle_mod, capi_mod):
    O, capi = oracle_mod, capi_mod
    traj, top, itp, frames = small_inputs(tag)
    ref = O.run(traj, top, itp, frames)
    with capi.Context(top, itp, 0) as c:
        c.open(traj)
        assert np.array_equal(c.groups(), ref["groups"])
        dE, done, rc = c.run(0, frames)
        assert rc == capi.CGC_OK and done == frames == ref["frames"]
        assert dE.shape == ref["dE_f64"].shape
        e, z = scaled_err(dE, ref["dE_f64"], ref["scale"])
        assert e <= DE_TOL, "max scaled error %g" % e
        assert z == 0.0                                    # the own-site column (and any empty column) stays exactly 0
        tot_err = np.abs(dE.astype(np.float64).sum(-1) - ref["dE_f64"].sum(-1))
        assert np.all(tot_err <= DE_TOL * np.maximum(np.abs(ref["dE_f64"].sum(-1)), ref["scale"].sum(-1)))
        # not worse than the reference's own float32 noise by more than a small factor
        e_ref, _ = scaled_err(ref["dE_f32"], ref["dE_f64"], ref["scale"])
        assert e <= max(4 * e_ref, 2e-7)
        assert c.decode_errors() == 0
        assert c.launch_count >= 3


def test_dE_against_reference_golden(small_inputs, capi_mod, oracle_mod):
    """Against raw float32 rows dumped by the unmodified reference (tests/golden/golden7.npz).  Both sides carry their own
    float32 rounding relative to exact arithmetic (the reference's is <= ~2e-7 of the cancellation scale, SURVEY.md
    section 8c), so the bound is 1e-6 + that, on the scale sum|term| which the oracle supplies."""
    capi, O = capi_mod, oracle_mod
    g = np.load(os.path.join(GOLD, "golden7.npz"))
    traj, top, itp, frames = small_inputs("golden7")
    if [sha(traj), sha(top), sha(itp)] != list(g["sha"]):
        pytest.skip("synthetic generator produced different bytes than when the golden file was made")
    scale = O.run(traj, top, itp, frames)["scale"]
    with capi.Context(top, itp, 0) as c:
        c.open(traj)
        assert np.array_equal(c.groups(), g["groups"])
        dE, done, _ = c.run(0, frames)
    ref = g["dE_f32"].astype(np.float64)
    scale = np.maximum(scale, np.abs(ref))
    assert np.all(np.abs(dE - ref) <= 1.2e-6 * scale)
    assert np.all(dE[scale == 0] == 0)
    # the .time text the stock CLI wrote for sites 1..7: 6 printed digits (half a unit in the 6th place) on top of that
    rows = np.array([[float(v) for v in ln.split(",")] for ln in str(g["time_text"]).strip().split("\n")])
    ours = (dE[:, :7].reshape(-1, dE.shape[-1]).astype(np.float64) * 349.7)
    sc = scale[:, :7].reshape(-1, dE.shape[-1]) * 349.7
    assert np.all(np.abs(ours - rows) <= 1.2e-6 * sc + 6e-6 * np.abs(rows))


def test_reference_fixture_on_gpu(ref_fixture, oracle_mod, capi_mod):
    """The reference's own test inputs: 6383 atoms, 7 BCL, 2 frames (reference src/test/)."""
    O, capi = oracle_mod, capi_mod
    traj, top, itp = ref_fixture
    ref = O.run(traj, top, itp, 2)
    with capi.Context(top, itp, 0) as c:
        c.open(traj)
        assert c.dims() == (6383, 7, 350, 357)
        dE, done, rc = c.run(0, 2)
    e, z = scaled_err(dE, ref["dE_f64"], ref["scale"])
    assert e <= DE_TOL and z == 0.0
    tot = (dE[0].astype(np.float64) * 349.7).sum(-1)      # BASELINE.md section 2 smoke values (6 printed digits)
    smoke = np.array([383.37, 132.194, -175.118, -276.294, -347.642, -104.302, -76.3561])
    assert np.all(np.abs(tot - smoke) <= 1.2e-6 * 349.7 * ref["scale"][0].sum(-1) + 6e-6 * np.abs(smoke))


@pytest.mark.parametrize("tag", ["tiny", "permol"])
def test_decode_bit_exact(tag, small_inputs, oracle_mod, capi_mod, torch_cuda):
    """K1 against (float)atof of the tokens (reference src/grom2atomFMO.cpp:448-450), every frame, bit for bit."""
    O, capi, torch = oracle_mod, capi_mod, torch_cuda
    traj, top, itp, frames = small_inputs(tag)
    with capi.Context(top, itp, 0) as c:
        c.open(traj)
        d_text, nbytes = _device_text(torch, traj)
        offs = [c.frame_atoms_offset(f) for f in range(frames)]
        xyz = torch.empty((frames, c.n_atoms, 3), dtype=torch.float32, device="cuda")
        c.decode_device(d_text.data_ptr(), nbytes, offs, frames, xyz.data_ptr())
        torch.cuda.synchronize()
        assert c.decode_errors() == 0
        got = xyz.cpu().numpy()
    pos = None
    for f in range(frames):
        pos, Xi, _, _, _ = O.read_frame(traj, pos)
        assert np.array_equal(got[f].reshape(-1).view(np.uint32), Xi.view(np.uint32))


def test_decode_negative_and_edge_values(tmp_path, capi_mod, torch_cuda, small_inputs):
    """Negative coordinates, -0.000, the largest field value, every digit position."""
    capi, torch = capi_mod, torch_cuda
    _, top, itp, _ = small_inputs("tiny")
    from cgromcorrfmo_b200 import synth
    names = synth.bcl_atom_names()
    vals = [-0.0, -0.001, 0.0, 0.001, -999.999, 9999.999, 1234.567, -12.345, 0.5, 8.415, 16777.215 % 10000, 0.009]
    rng = np.random.default_rng(3)
    lines = []
    for i in range(140):
        x, y, z = (vals[(i + k) % len(vals)] if i < 36 else float(np.round(rng.uniform(-999, 9999), 3)) for k in range(3))
        lines.append("%5d%-5s%5s%5d%8.3f%8.3f%8.3f\n" % (1, "BCL", names[i], i + 1, x, y, z))
    text = "edge t= 0.0\n  140\n" + "".join(lines) + "   1.0   1.0   1.0\n"
    p = tmp_path / "edge.gro"
    p.write_text(text)
    with capi.Context(top, itp, 0) as c:
        c.open(str(p))
        d_text, nbytes = _device_text(torch, str(p))
        xyz = torch.empty((1, 140, 3), dtype=torch.float32, device="cuda")
        c.decode_device(d_text.data_ptr(), nbytes, [c.frame_atoms_offset(0)], 1, xyz.data_ptr())
        torch.cuda.synchronize()
        got = xyz.cpu().numpy().reshape(-1)
    expect = np.array([np.float32(float(ln[20 + 8 * k: 28 + 8 * k])) for ln in lines for k in range(3)], np.float32)
    assert np.array_equal(got.view(np.uint32), expect.view(np.uint32))      # includes the sign bit of -0.000


def test_decode_flags_malformed_text(tmp_path, capi_mod, small_inputs):
    """A coordinate field damaged in a LATER frame is caught on the device and surfaces as InputFileMalformed."""
    capi = capi_mod
    traj, top, itp, frames = small_inputs("tiny")
    data = bytearray(open(traj, "rb").read())
    with capi.Context(top, itp, 0) as c:
        c.open(traj)
        off = c.frame_atoms_offset(1)
    data[off + 45 * 10 + 23] = ord("x")
    bad = tmp_path / "bad.gro"
    bad.write_bytes(bytes(data))
    with capi.Context(top, itp, 0) as c:
        c.open(str(bad))
        with pytest.raises(capi.CgcError) as e:
            c.run(0, frames)
        assert e.value.code == capi.CGC_ERR_INPUT_MALFORMED


def test_device_and_host_paths_agree_and_frames_are_independent(small_inputs, capi_mod, torch_cuda):
    capi, torch = capi_mod, torch_cuda
    traj, top, itp, frames = small_inputs("tiny")
    with capi.Context(top, itp, 0) as c:
        c.open(traj)
        full, done, _ = c.run(0, frames)
        d_text, nbytes = _device_text(torch, traj)
        offs = [c.frame_atoms_offset(f) for f in range(frames)]
        dEd = torch.empty((frames, c.n_sites, c.num_tot), dtype=torch.float32, device="cuda")
        c.run_device(d_text.data_ptr(), nbytes, offs, frames, dEd.data_ptr())
        torch.cuda.synchronize()
        dev = dEd.cpu().numpy()
        # a frame's row does not depend on which batch it travels in (sharding across ranks relies on this)
        c.set_option("batch_frames", 1)
        one_by_one, _, _ = c.run(0, frames)
        tail, done_tail, _ = c.run(1, frames - 1)
    s = np.abs(full).max()
    assert np.abs(dev - full).max() <= 1e-6 * s
    assert np.abs(one_by_one - full).max() <= 1e-6 * s
    assert done_tail == frames - 1 and np.abs(tail - full[1:]).max() <= 1e-6 * s


def test_eof_is_clean(small_inputs, capi_mod):
    """Asking for more frames than the file holds returns CGC_EOF with the frames that exist
    (the reference segfaults, reference README.md:31)."""
    capi = capi_mod
    traj, top, itp, frames = small_inputs("tiny")
    with capi.Context(top, itp, 0) as c:
        c.open(traj)
        dE, done, rc = c.run(0, frames + 5)
        assert rc == capi.CGC_EOF and done == frames and dE.shape[0] == frames
        dE2, done2, rc2 = c.run(frames, 2)
        assert rc2 == capi.CGC_EOF and done2 == 0


def test_sites_option_matches_reference_seven(small_inputs, oracle_mod, capi_mod):
    """--sites 7 reproduces the stock program's site loop (reference src/FMOparse.cpp:133) on an 8-BCL system."""
    O, capi = oracle_mod, capi_mod
    traj, top, itp, frames = small_inputs("tiny")
    ref = O.run(traj, top, itp, frames, n_sites=7)
    with capi.Context(top, itp, 0) as c:
        c.open(traj)
        c.set_option("sites", 7)
        dE, done, _ = c.run(0, frames)
    assert dE.shape[1] == 7
    e, z = scaled_err(dE, ref["dE_f64"], ref["scale"])
    assert e <= DE_TOL and z == 0.0


@pytest.mark.parametrize("tag", ["tiny", "permol"])
def test_covariance(tag, small_inputs, oracle_mod, capi_mod):
    O, capi = oracle_mod, capi_mod
    traj, top, itp, frames = small_inputs(tag)
    with capi.Context(top, itp, 0) as c:
        c.open(traj)
        c.set_option("batch_frames", 2)                   # more than one batch: sums carry across launches
        dE, done, _ = c.run(0, frames)
        for ddof in (0, 1):
            if frames - ddof <= 0:
                continue
            mean, cov = c.cov_finalize(ddof)
            m64, c64 = O.cov_f64(dE, ddof)
            scale = np.maximum(np.abs(c64), np.sqrt(np.einsum("sii,sjj->sij", c64, c64)))
            ok = scale > 0
            assert np.all(np.abs(cov - c64)[ok] <= COV_TOL * scale[ok])
            assert np.all(cov[~ok] == 0) if (~ok).any() else True
            assert np.allclose(mean, m64, rtol=1e-12, atol=1e-300)
            assert np.array_equal(cov, np.swapaxes(cov, 1, 2))          # exactly symmetric
        part = c.cov_partials()
        assert part[0] == frames


def test_covariance_long_random_rows(capi_mod, torch_cuda, small_inputs):
    """K3 alone on many frames of synthetic rows with a large mean (the cancellation the shift is there for)."""
    capi, torch = capi_mod, torch_cuda
    traj, top, itp, frames = small_inputs("tiny")
    with capi.Context(top, itp, 0) as c:
        c.open(traj)
        S, G = c.n_sites, c.num_tot
        rng = np.random.default_rng(11)
        T = 3000
        X = (rng.normal(size=(T, S, G)) * 0.05 + rng.normal(size=(1, S, G)) * 50.0).astype(np.float32)
        d = torch.from_numpy(X).cuda()
        c.cov_reset()
        for t0 in range(0, T, 700):
            n = min(700, T - t0)
            c.cov_accumulate_device(d[t0:t0 + n].contiguous().data_ptr(), n)
        torch.cuda.synchronize()
        mean, cov = c.cov_finalize(1)
    X64 = X.astype(np.float64)
    for k in range(S):
        ref = np.cov(X64[:, k, :], rowvar=False, ddof=1)      # depca's np.cov (sidechain_corr.py:67)
        scale = np.maximum(np.abs(ref), np.sqrt(np.outer(np.diag(ref), np.diag(ref))))
        assert np.all(np.abs(cov[k] - ref) <= COV_TOL * scale)
        assert np.allclose(mean[k], X64[:, k, :].mean(0), rtol=1e-12)


def test_writers_match_reference_format(small_inputs, oracle_mod, capi_mod, tmp_path):
    """.time and the compat .mtx are byte-identical to the reference's formatting of the same float32 rows."""
    O, capi = oracle_mod, capi_mod
    traj, top, itp, frames = small_inputs("tiny")
    with capi.Context(top, itp, 0) as c:
        c.open(traj)
        dE, done, _ = c.run(0, frames)
        with pytest.raises(capi.CgcError):                    # the literal .mtx needs its float32 sums kept while running
            c.write_mtx(str(tmp_path / "o_compat.mtx"), compat=True)
    with capi.Context(top, itp, 0) as c:
        c.open(traj)
        c.set_option("mtx_compat", 1)
        c.set_option("batch_frames", 2)                       # the sums carry across launches, in frame order
        dE, done, _ = c.run(0, frames)
        capi.write_time(str(tmp_path / "o.time"), dE)
        assert open(tmp_path / "o.time").read() == O.format_time_rows(dE.reshape(-1, dE.shape[-1]))
        c.write_mtx(str(tmp_path / "o_compat.mtx"), compat=True)
        assert open(tmp_path / "o_compat.mtx").read() == O.format_mtx(O.mtx_compat_f32(dE))
        c.write_mtx(str(tmp_path / "o.mtx"), compat=False)
        _, c64 = O.cov_f64(dE, 0)
        assert open(tmp_path / "o.mtx").read() == O.format_mtx(c64.astype(np.float32))


def test_cli_drop_in(small_inputs, oracle_mod, capi_mod, tmp_path):
    """The CLI driver with the reference's five positionals: exit codes, .time/.mtx shapes, parity with the stock binary
    (when oracle/_ref travelled to this box) to the 6 printed digits."""
    import subprocess
    from cgromcorrfmo_b200 import build
    O = oracle_mod
    traj, top, itp, frames = small_inputs("golden7")
    exe = build.CLI
    assert os.path.exists(exe)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "Order of args" in r.stdout                      # reference src/FMOparse.cpp:325-329
    out = str(tmp_path / "out")
    r = subprocess.run([exe, traj, top, itp, out, str(frames), "--mtx-compat"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rows = [ln.split(",") for ln in open(out + ".time").read().strip().split("\n")]
    assert len(rows) == frames * 7 and all(len(x) == len(rows[0]) for x in rows)
    G = len(rows[0])
    mtx = open(out + ".mtx").read().strip().split("\n")
    assert len(mtx) == 7 * G and all(len(x.split(",")) == G for x in mtx)
    g = np.load(os.path.join(GOLD, "golden7.npz"))
    if [sha(traj), sha(top), sha(itp)] == list(g["sha"]):
        scale = O.run(traj, top, itp, frames)["scale"][:, :7].reshape(-1, G) * 349.7
        ref_rows = np.array([[float(v) for v in ln.split(",")] for ln in str(g["time_text"]).strip().split("\n")])
        ours = np.array(rows, dtype=np.float64)
        # both files print 6 significant digits
        assert np.all(np.abs(ours - ref_rows) <= 1.2e-6 * scale + 1.2e-5 * np.abs(ref_rows))
        ref_m = np.array([[float(v) for v in ln.split(",")] for ln in str(g["mtx_text"]).strip().split("\n")])
        ours_m = np.array([[float(v) for v in ln.split(",")] for ln in mtx])
        assert np.all(np.abs(ours_m - ref_m) <= 1e-4 * np.abs(ref_m).max())
    # more frames than the file holds: clean stop, exit 0, .time holds what exists
    r = subprocess.run([exe, traj, top, itp, out + "2", str(frames + 4)], capture_output=True, text=True)
    assert r.returncode == 0 and "ended after" in r.stderr
    assert len(open(out + "2.time").read().strip().split("\n")) == frames * 7
    # malformed input: the reference dies with SIGABRT (134) on InputFileMalformed
    r = subprocess.run([exe, traj, str(tmp_path / "missing.top"), itp, out + "3", "1"], capture_output=True, text=True)
    assert r.returncode == 134 and "InputFileMalformed" in r.stderr


def test_full_size_properties(capi_mod, torch_cuda):
    """BASELINE.json's C2 shape (100k atoms, 24 sites): properties that need no CPU oracle.
      (a) per-site totals equal an independent float64 evaluation of the same double sum in torch on the GPU;
      (b) splitting one group into two leaves every other column unchanged and the two halves add up to the original."""
    capi, torch = capi_mod, torch_cuda
    from cgromcorrfmo_b200 import synth
    import tempfile
    s = synth.named_system("C2")
    prefix = synth._static_prefix(s)
    with tempfile.TemporaryDirectory() as d:
        top, itp = os.path.join(d, "t.top"), os.path.join(d, "b.itp")
        synth.write_top(top, s)
        synth.write_itp(itp, s)
        blob = synth.frame_bytes(s, prefix, 0) + synth.frame_bytes(s, prefix, 1)
        with capi.Context(top, itp, 0) as c:
            c.open_memory(blob)
            n_atoms, n_chromo, n_groups, G = c.dims()
            assert (n_atoms, n_chromo) == (100000, 24)
            dE, done, _ = c.run(0, 2)
            assert done == 2
            groups = c.groups()
            x = torch.tensor(synth.frame_milli(s, 0), dtype=torch.float64, device="cuda") / 1000.0
            x = x.to(torch.float32).to(torch.float64)     # the float32 values the decoder produces
            tot = np.zeros(n_chromo)
            absum = np.zeros(n_chromo)
            for site in range(1, n_chromo + 1):
                q = torch.tensor(c.site_charges(site), dtype=torch.float64, device="cuda")
                sel = torch.tensor(groups == -site, device="cuda")
                dq = q[sel]
                nz = dq != 0
                xc, dq = x[sel][nz], dq[nz]
                env = ~sel
                r = torch.cdist(x[env], xc)
                terms = (float(np.float32(33.2)) * q[env])[:, None] * dq[None, :] / r
                tot[site - 1] = terms.sum().item()
                absum[site - 1] = terms.abs().sum().item()
            ours = dE[0].astype(np.float64).sum(-1)
            assert np.all(np.abs(ours - tot) <= DE_TOL * np.maximum(np.abs(tot), absum))
        # (b) re-number the second half of one protein residue: its column splits in two
        a0 = int(np.nonzero(groups == 5)[0][0])
        n5 = int((groups == 5).sum())
        assert n5 >= 4
        resnr2 = s.resnr.copy()
        resnr2[a0 + n5 // 2: a0 + n5] = 99999                # a number no neighbour uses
        import dataclasses
        s2 = dataclasses.replace(s, resnr=resnr2)
        top2 = os.path.join(d, "t2.top")
        synth.write_top(top2, s2)
        blob2 = synth.frame_bytes(s2, synth._static_prefix(s2), 0)
        with capi.Context(top2, itp, 0) as c2:
            c2.open_memory(blob2)
            assert c2.num_tot == G + 1
            dE2, _, _ = c2.run(0, 1)
        col = n_chromo + 5 - 1
        merged = np.concatenate([dE2[0][:, :col], (dE2[0][:, col] + dE2[0][:, col + 1])[:, None], dE2[0][:, col + 2:]], axis=1)
        yard = np.abs(dE[0]).max()
        assert np.abs(merged - dE[0]).max() <= 2e-6 * yard


def test_more_chromophores_than_groups(tmp_path, oracle_mod, capi_mod):
    """The reference's outer loop runs over 1..nGroups (reference src/grom2atomFMO.cpp:669), so with nChromo > nGroups
    the columns of the last chromophores stay 0.  Kept, checked against the oracle."""
    O, capi = oracle_mod, capi_mod
    from cgromcorrfmo_b200 import synth
    s = synth.make_system(3 * 140 + 9, 3, 21, "none")          # one 9-atom residue + 3 BCL
    traj, top, itp = synth.write_inputs(str(tmp_path), s, 2)
    ref = O.run(traj, top, itp, 2)
    assert ref["n_chromo"] == 3 and ref["n_groups"] < 3
    with capi.Context(top, itp, 0) as c:
        c.open(traj)
        dE, done, _ = c.run(0, 2)
    e, z = scaled_err(dE, ref["dE_f64"], ref["scale"])
    assert e <= DE_TOL and z == 0.0
    assert np.all(dE[:, :, ref["n_groups"]:3] == 0)


def test_zero_frames_and_reopen(small_inputs, capi_mod):
    capi = capi_mod
    traj, top, itp, frames = small_inputs("tiny")
    traj2, top2, itp2, frames2 = small_inputs("nosolv")
    with capi.Context(top, itp, 0) as c:
        c.open(traj)
        dE, done, rc = c.run(0, 0)
        assert rc == capi.CGC_OK and done == 0 and dE.shape[0] == 0
        a, _, _ = c.run(0, frames)
        c.open(traj)                                          # re-opening resets the Correlate sums and options
        b, _, _ = c.run(0, frames)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))   # grid-rounded partial sums: bit-identical from run to run
        assert c.cov_partials()[0] == frames
    with capi.Context(top2, itp2, 0) as c:
        c.open(traj2)
        assert c.run(0, 1)[1] == 1


def test_wide_coordinate_fields(tmp_path, capi_mod, torch_cuda, small_inputs):
    """%10.5f coordinates (GROMACS writes them with trjconv -ndec 5): the decoder takes the double-precision division
    path; still bit-exact with (float)atof."""
    capi, torch = capi_mod, torch_cuda
    _, top, itp, _ = small_inputs("tiny")
    from cgromcorrfmo_b200 import synth
    names = synth.bcl_atom_names()
    rng = np.random.default_rng(9)
    lines = []
    for i in range(140):
        x, y, z = (float(np.round(rng.uniform(-99, 999), 5)) for _ in range(3))
        lines.append("%5d%-5s%5s%5d%10.5f%10.5f%10.5f\n" % (1, "BCL", names[i], i + 1, x, y, z))
    p = tmp_path / "wide.gro"
    p.write_text("wide t= 0.0\n  140\n" + "".join(lines) + "   1.0   1.0   1.0\n")
    with capi.Context(top, itp, 0) as c:
        c.open(str(p))
        assert c.layout() == dict(line_bytes=51, x_col=20, field_width=10, decimals=5)
        d_text, nbytes = _device_text(torch, str(p))
        xyz = torch.empty((1, 140, 3), dtype=torch.float32, device="cuda")
        c.decode_device(d_text.data_ptr(), nbytes, [c.frame_atoms_offset(0)], 1, xyz.data_ptr())
        torch.cuda.synchronize()
        got = xyz.cpu().numpy().reshape(-1)
    expect = np.array([np.float32(float(ln[20 + 10 * k: 30 + 10 * k])) for ln in lines for k in range(3)], np.float32)
    assert np.array_equal(got.view(np.uint32), expect.view(np.uint32))


def test_two_rank_sharding_emulated_on_one_gpu(small_inputs, oracle_mod, capi_mod):
    """Frame-range sharding as the multi-GPU driver does it (cgromcorrfmo_b200/run.py), with both "ranks" as separate
    contexts on one GPU: each streams its range, the partial sums are added element-wise with the shift block restored
    (what dist.allreduce_partials does over NCCL), and the merged sums finalise to the covariance of ALL rows.
    Rank 1 does not start at frame 0, so it evaluates frame 0 only to obtain the common shift."""
    O, capi = oracle_mod, capi_mod
    from cgromcorrfmo_b200 import dist as cdist
    traj, top, itp, frames = small_inputs("tiny")
    parts, rows = [], []
    for rank in range(2):
        lo, n = cdist.frame_range(rank, 2, frames)
        with capi.Context(top, itp, 0) as c:
            c.open(traj)
            dE, done, _ = c.run(lo, n)
            assert done == n
            rows.append(dE)
            parts.append(c.cov_partials())
            S, G = c.n_sites, c.num_tot
    lay = cdist.partials_layout(S, G)
    assert np.array_equal(parts[0][lay["shift"]:lay["sum"]], parts[1][lay["shift"]:lay["sum"]])   # same shift on both
    allrows = np.concatenate(rows)
    assert np.array_equal(parts[0][lay["shift"]:lay["sum"]], allrows[0].astype(np.float64).reshape(-1))  # = row of frame 0
    merged = parts[0] + parts[1]
    merged[lay["shift"]:lay["sum"]] = parts[0][lay["shift"]:lay["sum"]]
    n = merged[0]
    assert n == frames
    s = merged[lay["sum"]:lay["sumsq"]].reshape(S, G)
    ss = merged[lay["sumsq"]:].reshape(S, G, G)
    iu = np.triu_indices(G)
    cov = (ss - np.einsum("si,sj->sij", s, s) / n) / n
    _, c64 = O.cov_f64(allrows, 0)
    scale = np.maximum(np.abs(c64), np.sqrt(np.einsum("sii,sjj->sij", c64, c64)))
    for k in range(S):
        # the device stores the upper block-triangle (64x64 blocks); G < 64 here, so the whole matrix is one block
        ok = scale[k] > 0
        assert np.all(np.abs(cov[k] - c64[k])[ok] <= COV_TOL * scale[k][ok])


def test_device_entry_points_reject_bad_arguments(small_inputs, capi_mod, torch_cuda):
    capi, torch = capi_mod, torch_cuda
    traj, top, itp, frames = small_inputs("tiny")
    with capi.Context(top, itp, 0) as c:
        d = torch.zeros(1024, dtype=torch.uint8, device="cuda")
        out = torch.zeros(16, dtype=torch.float32, device="cuda")
        with pytest.raises(capi.CgcError) as e:           # no trajectory opened yet
            c.run_device(d.data_ptr(), 1024, [0], 1, out.data_ptr())
        assert e.value.code == capi.CGC_ERR_STATE
        c.open(traj)
        with pytest.raises(capi.CgcError) as e:           # text pointer must be 16-byte aligned (bulk TMA source)
            c.run_device(d.data_ptr() + 4, 1000, [0], 1, out.data_ptr())
        assert e.value.code == capi.CGC_ERR_ARG
        with pytest.raises(capi.CgcError) as e:
            c.run_device(d.data_ptr(), 1024, [0], 0, out.data_ptr())
        assert e.value.code == capi.CGC_ERR_ARG
        with pytest.raises(capi.CgcError) as e:
            c.run(-1, 2)
        assert e.value.code == capi.CGC_ERR_ARG


def test_kernel_timing_hook(small_inputs, capi_mod):
    """Option "profile": CUDA events around every hot-path launch, as bench.py uses for the per-kernel rooflines."""
    capi = capi_mod
    traj, top, itp, frames = small_inputs("tiny")
    with capi.Context(top, itp, 0) as c:
        c.open(traj)
        c.set_option("profile", 1)
        n0 = c.launch_count
        c.run(0, frames)
        kt = c.kernel_times(reset=True)
        assert kt["decode"][1] == 1 and kt["cdc"][1] == 1 and kt["finish"][1] == 1 and kt["cov"][1] == 1
        assert all(ms > 0 for ms, _ in kt.values())
        assert c.launch_count - n0 >= 5
        assert c.kernel_times()["cdc"] == (0.0, 0)
        ms, copies = c.copy_times(reset=True)              # the H2D copy of the batch's text, on the copy stream
        assert copies == 1 and ms > 0 and c.copy_times() == (0.0, 0)


def test_c1_shape_against_the_stock_reference_program(tmp_path, oracle_mod, capi_mod):
    """BASELINE.json configs[0]: the monomer shape (30 000 atoms, 7 BCL) that the UNMODIFIED reference program can run as
    is.  Its .time text (7 sites, 6 printed digits) against the CLI driver's, three frames; group ids bit-exact."""
    import subprocess
    from cgromcorrfmo_b200 import build, synth
    O = oracle_mod
    if not O.have_ref():
        pytest.skip("oracle/_ref (the compiled reference) is not on this box")
    s = synth.named_system("C1")
    frames = 3
    traj, top, itp = synth.write_inputs(str(tmp_path / "c1"), s, frames)
    assert O.run_ref_cli(traj, top, itp, str(tmp_path / "ref"), frames, timeout=900) == 0
    r = subprocess.run([build.CLI, traj, top, itp, str(tmp_path / "ours"), str(frames)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ref_rows = np.array([[float(v) for v in ln.split(",")] for ln in open(tmp_path / "ref.time").read().strip().split("\n")])
    our_rows = np.array([[float(v) for v in ln.split(",")] for ln in open(tmp_path / "ours.time").read().strip().split("\n")])
    assert ref_rows.shape == our_rows.shape == (frames * 7, ref_rows.shape[1])
    # the cancellation scale of every entry, from the oracle (frame 0 only: one frame of 30k atoms x 7 sites is seconds)
    one = str(tmp_path / "one.gro")
    first = open(traj, "rb").read()
    open(one, "wb").write(first[: len(first) // frames])
    o = O.run(one, top, itp, 1)
    with capi_mod.Context(top, itp, 0) as c:
        c.open(traj)
        assert np.array_equal(c.groups(), o["groups"])
    scale = o["scale"][0] * 349.7
    # both files print 6 significant digits; the reference's own float32 noise is ~1e-7 of the scale
    tol = 1.3e-6 * np.tile(scale, (frames, 1)) * 1.5 + 1.2e-5 * np.abs(ref_rows)
    assert np.all(np.abs(our_rows - ref_rows) <= tol)
    assert np.all(our_rows[np.tile(scale, (frames, 1)) == 0] == 0)


@pytest.mark.parametrize("n_dq", [1, 2, 9, 31, 100, 140])
def test_other_dq_atom_counts(n_dq, tmp_path, oracle_mod, capi_mod):
    """The pair kernel's main loop is instantiated for 42 site atoms per trip (the 44 dQ != 0 atoms of the stock bcx.itp)
    and for 7 otherwise, with the site blocks padded by neutral entries: excited-state tables with 1 .. 140 non-zero dQ
    atoms, several frames and a site count that does not fill the last chunk."""
    O, capi = oracle_mod, capi_mod
    from cgromcorrfmo_b200 import synth
    s = synth.make_system(4000, 11, 300 + n_dq, "shared")
    has_b = np.zeros(140, dtype=bool)
    has_b[np.random.default_rng(n_dq).permutation(140)[:n_dq]] = True
    s.bcl_has_b = has_b
    s.bcl_qB = np.where(has_b, np.round(s.bcl_qA + 0.05 + 0.01 * np.arange(140), 3), s.bcl_qA)
    traj, top, itp = synth.write_inputs(str(tmp_path), s, 3)
    ref = O.run(traj, top, itp, 3)
    with capi.Context(top, itp, 0) as c:
        c.open(traj)
        assert int(c.get_option("site_cap")) >= max(2, n_dq)
        dE, done, _ = c.run(0, 3)
        assert done == 3
    e, z = scaled_err(dE, ref["dE_f64"], ref["scale"])
    assert e <= DE_TOL and z == 0.0, (n_dq, e, z)


def test_cli_npy_sidecar(small_inputs, capi_mod, tmp_path):
    """--npy: the rows of csv_out.time as raw float32 (frames, sites, numTot), for the pyPCA scripts."""
    import subprocess
    from cgromcorrfmo_b200 import build, pypca
    traj, top, itp, frames = small_inputs("golden7")
    out = str(tmp_path / "o")
    r = subprocess.run([build.CLI, traj, top, itp, out, str(frames + 2), "--npy", "--sites", "7"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    E = np.load(out + ".time.npy")
    rows = np.array([[float(v) for v in ln.split(",")] for ln in open(out + ".time").read().strip().split("\n")])
    assert E.dtype == np.float32 and E.shape == (frames, 7, rows.shape[1])
    cm = pypca.load_E_t_ia_npy(out + ".time.npy")
    assert np.allclose(cm.reshape(-1, rows.shape[1]), rows, rtol=6e-6, atol=0)      # the text has 6 significant digits
    with capi_mod.Context(top, itp, 0) as c:
        c.open(traj)
        c.set_option("sites", 7)
        dE, _, _ = c.run(0, frames)
    d = np.abs(E.astype(np.float64) - dE.astype(np.float64))
    assert np.all(d <= 1e-6 * np.maximum(np.abs(dE), 1e-3))


@pytest.mark.parametrize("seed", range(8))
def test_random_shapes_against_oracle(seed, tmp_path, oracle_mod, capi_mod):
    """Random small systems: atom counts that are not a multiple of the 2048-atom tile or the 256-line decode chunk, 1..13
    chromophores, 1..9 frames, every batch size — the pair kernel's unit ranges then start and end in the middle of a
    (frame, tile) and its chunk size varies with the launch."""
    O, capi = oracle_mod, capi_mod
    from cgromcorrfmo_b200 import synth
    rng = np.random.default_rng(1000 + seed)
    n_bcl = int(rng.choice([1, 2, 3, 5, 8, 13]))
    n_atoms = int(n_bcl * 140 * 1.3) + int(rng.integers(50, 6000))
    frames = int(rng.integers(1, 10))
    solvent = str(rng.choice(["shared", "per_molecule", "none"]))
    s = synth.make_system(n_atoms, n_bcl, 500 + seed, solvent)
    traj, top, itp = synth.write_inputs(str(tmp_path), s, frames, velocities=bool(rng.integers(0, 2)))
    ref = O.run(traj, top, itp, frames)
    with capi.Context(top, itp, 0) as c:
        c.open(traj)
        assert np.array_equal(c.groups(), ref["groups"])
        c.set_option("batch_frames", int(rng.integers(1, frames + 2)))
        first = int(rng.integers(0, frames))
        dE, done, rc = c.run(first, frames)                       # over-asks unless first == 0
        assert done == frames - first
        e, z = scaled_err(dE, ref["dE_f64"][first:], ref["scale"][first:])
        assert e <= DE_TOL and z == 0.0, (seed, n_atoms, n_bcl, frames, solvent, e, z)
        sites = int(rng.integers(1, n_bcl + 1))
    with capi.Context(top, itp, 0) as c:
        c.open(traj)
        c.set_option("sites", sites)
        dE, done, rc = c.run(0, frames)
        e, z = scaled_err(dE, ref["dE_f64"][:, :sites], ref["scale"][:, :sites])
        assert dE.shape[1] == sites and e <= DE_TOL and z == 0.0
        if done > 1 and c.cov_size():
            mean, cov = c.cov_finalize(0)
            _, c64 = O.cov_f64(dE, 0)
            cs = np.maximum(np.abs(c64), np.sqrt(np.einsum("sii,sjj->sij", c64, c64)))
            assert np.all(np.abs(cov - c64)[cs > 0] <= COV_TOL * cs[cs > 0])


def test_python_driver_two_ranks_on_one_gpu(small_inputs, capi_mod, tmp_path):
    """python -m cgromcorrfmo_b200.run under torchrun with two ranks (gloo, both on cuda:0): each rank streams its frame
    range and writes its .time part with the device formatter, the Correlate sums are all-reduced once, rank 0 joins the
    parts and writes the .mtx.  Same files as one process."""
    import subprocess
    import sys
    from cgromcorrfmo_b200 import build
    traj, top, itp, frames = small_inputs("tiny")
    one, two = str(tmp_path / "one"), str(tmp_path / "two")
    r = subprocess.run([build.CLI, traj, top, itp, one, str(frames)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    env = dict(os.environ, CGC_DIST_BACKEND="gloo", PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", "-m", "cgromcorrfmo_b200.run", traj, top, itp, two, str(frames),
                        "--device", "0"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    a = np.array([[float(v) for v in ln.split(",")] for ln in open(one + ".time").read().strip().split("\n")])
    b = np.array([[float(v) for v in ln.split(",")] for ln in open(two + ".time").read().strip().split("\n")])
    assert a.shape == b.shape and np.allclose(a, b, rtol=3e-5, atol=1e-4)
    assert not os.path.exists(two + ".time.part0")
    ma = np.array([[float(v) for v in ln.split(",")] for ln in open(one + ".mtx").read().strip().split("\n")])
    mb = np.array([[float(v) for v in ln.split(",")] for ln in open(two + ".mtx").read().strip().split("\n")])
    assert ma.shape == mb.shape and np.all(np.abs(ma - mb) <= 2e-5 * np.abs(ma).max() + 1e-12)


def test_decode_every_8_3_value(small_inputs, capi_mod, torch_cuda):
    """K1 on EVERY value a "%8.3f" field can hold — 0.000 .. 9999.999 and -0.001 .. -999.999, eleven million fields —
    against (float)atof of the same text (reference src/grom2atomFMO.cpp:448-450): the correctly rounded double N/1000
    rounded to float32."""
    capi, torch = capi_mod, torch_cuda
    from cgromcorrfmo_b200 import synth
    from conftest import SMALL_SYSTEMS
    traj, top, itp, _ = small_inputs("tiny")
    cfg = SMALL_SYSTEMS["tiny"]
    s = synth.make_system(cfg["n_atoms"], cfg["n_bcl"], cfg["seed"], cfg["solvent"])
    n = s.n_atoms
    milli = np.concatenate([np.arange(0, 10_000_000, dtype=np.int64), -np.arange(1, 1_000_000, dtype=np.int64)])
    per_frame = 3 * n
    frames = (milli.size + per_frame - 1) // per_frame
    milli = np.concatenate([milli, np.zeros(frames * per_frame - milli.size, np.int64)]).reshape(frames, n, 3)
    prefix = synth._static_prefix(s)
    head = ("all values t= 0.0\n%5d\n" % n).encode()
    tail = b"   1.0   1.0   1.0\n"
    lines = np.empty((frames, n, 45), dtype=np.uint8)
    lines[:, :, :20] = prefix
    lines[:, :, 20:44] = synth._fmt_fixed(milli, 8, 3).reshape(frames, n, 24)
    lines[:, :, 44] = ord("\n")
    blob = b"".join(head + lines[f].tobytes() + tail for f in range(frames))
    expect = (milli.astype(np.float64) / 1000.0).astype(np.float32)
    text = np.frombuffer(blob, dtype=np.uint8)
    d_text = torch.from_numpy(np.concatenate([text, np.zeros(64, np.uint8)])).cuda()
    with capi.Context(top, itp, 0) as c:
        c.open_memory(blob)
        assert c.frame_count() == frames
        got = np.empty((frames, n, 3), np.float32)
        step = 256
        c.set_option("batch_frames", step)
        xyz = torch.empty((step, n, 3), dtype=torch.float32, device="cuda")
        for f0 in range(0, frames, step):
            k = min(step, frames - f0)
            offs = [c.frame_atoms_offset(f) for f in range(f0, f0 + k)]
            c.decode_device(d_text.data_ptr(), text.size, offs, k, xyz.data_ptr())
            torch.cuda.synchronize()
            got[f0:f0 + k] = xyz[:k].cpu().numpy()
        assert c.decode_errors() == 0
    assert np.array_equal(got.view(np.uint32), expect.view(np.uint32))


def test_cli_checkpoint_and_resume(small_inputs, tmp_path):
    """--checkpoint / --resume: a run stopped after some frames and continued gives the files of an uninterrupted run (the
    reference loses its Correlate sums when it is killed; only the .time prefix survives)."""
    import subprocess
    from cgromcorrfmo_b200 import build
    traj, top, itp, frames = small_inputs("tiny")
    full, part, ck = str(tmp_path / "full"), str(tmp_path / "part"), str(tmp_path / "ck.bin")
    r = subprocess.run([build.CLI, traj, top, itp, full, str(frames), "--npy"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([build.CLI, traj, top, itp, part, "1", "--npy", "--checkpoint", ck], capture_output=True, text=True)
    assert r.returncode == 0 and os.path.exists(ck), r.stderr
    # garbage appended after the checkpoint (a run killed in the middle of a chunk) is cut off again on resume
    with open(part + ".time", "ab") as f:
        f.write(b"1,2,3\n")
    r = subprocess.run([build.CLI, traj, top, itp, part, str(frames), "--npy", "--resume", ck, "--checkpoint", ck],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    load = lambda p: np.array([[float(v) for v in ln.split(",")] for ln in open(p).read().strip().split("\n")])
    a, b = load(full + ".time"), load(part + ".time")
    assert a.shape == b.shape and np.allclose(a, b, rtol=3e-5, atol=1e-4)
    ma, mb = load(full + ".mtx"), load(part + ".mtx")
    assert ma.shape == mb.shape and np.all(np.abs(ma - mb) <= 2e-5 * np.abs(ma).max() + 1e-12)
    ea, eb = np.load(full + ".time.npy"), np.load(part + ".time.npy")
    assert ea.shape == eb.shape == (frames, ea.shape[1], ea.shape[2])
    assert np.allclose(ea, eb, rtol=1e-5, atol=1e-6)
    # a checkpoint of another run is refused
    r = subprocess.run([build.CLI, traj, top, itp, part, str(frames), "--resume", ck, "--sites", "3"], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot resume" in r.stderr


@pytest.mark.gpu
def test_cli_sigint_stops_after_the_current_chunk(small_inputs, tmp_path):
    """SIGINT: the reference finishes the frame in hand, writes the .mtx of the frames seen and exits 0
    (src/FMOparse.cpp:24-29,62,68); so does the CLI, chunk-wise.  The rows written are whole and are a prefix of a full run."""
    import signal
    import subprocess
    import time
    from cgromcorrfmo_b200 import build
    traj, top, itp, frames = small_inputs("tiny")
    blob = open(traj, "rb").read()
    reps = 40
    long_traj = str(tmp_path / "long.gro")
    with open(long_traj, "wb") as f:
        for _ in range(reps):
            f.write(blob)
    total = frames * reps
    full, part = str(tmp_path / "full"), str(tmp_path / "part")
    r = subprocess.run([build.CLI, long_traj, top, itp, full, str(total), "--batch-frames", "1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    env = dict(os.environ, CGC_CLI_CHUNK_PAUSE_MS="400")
    p = subprocess.Popen([build.CLI, long_traj, top, itp, part, str(total), "--batch-frames", "1"], env=env,
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    deadline = time.time() + 60
    while time.time() < deadline and not (os.path.exists(part + ".time") and os.path.getsize(part + ".time") > 0):
        time.sleep(0.02)
    p.send_signal(signal.SIGINT)
    out, err = p.communicate(timeout=60)
    assert p.returncode == 0, err
    assert "Received cancellation" in err
    want = open(full + ".time", "rb").read()
    got = open(part + ".time", "rb").read()
    assert 0 < len(got) < len(want) and want.startswith(got) and got.endswith(b"\n")
    n_rows = got.count(b"\n")
    assert n_rows % 8 == 0 and n_rows // 8 < total          # whole frames (8 chromophores), fewer than asked for
    assert os.path.exists(part + ".mtx")

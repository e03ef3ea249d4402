"""GROMACS .trr input.  The reference has no .trr reader (its authors convert to .gro text with
trjconv first, reference scripts/2014-09-trr2dEcsv.sh:22), so parity is anchored on the .gro path: a .trr holding the
same numbers as the "%8.3f" text must give the same groups, frame count and — on the GPU — bit-identical dE rows."""
import os
import subprocess

import numpy as np
import pytest

from cgromcorrfmo_b200 import capi, synth
from conftest import SMALL_SYSTEMS, scaled_err

VARIANTS = {
    "f32_box": dict(double=False, velocities=False, forces=False, box=True),
    "f64_box_v_f": dict(double=True, velocities=True, forces=True, box=True),
    "f32_nobox_v": dict(double=False, velocities=True, forces=False, box=False),
}


def _system(tag):
    cfg = SMALL_SYSTEMS[tag]
    return synth.make_system(cfg["n_atoms"], cfg["n_bcl"], cfg["seed"], cfg["solvent"]), cfg["frames"]


@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_trr_front_end_on_cpu(variant, small_inputs, oracle_mod, tmp_path):
    O = oracle_mod
    traj, top, itp, frames = small_inputs("tiny")
    s, _ = _system("tiny")
    trr = str(tmp_path / "t.trr")
    synth.write_trr(trr, s, frames, **VARIANTS[variant])
    fr = O.read_trr(trr)
    assert len(fr) == frames
    # the oracle's .trr reader sees exactly the float32 coordinates the text path parses
    pos = None
    for t in range(frames):
        pos, Xi, keys, groups, _ = O.read_frame(traj, pos)
        assert np.array_equal(fr[t]["x"].reshape(-1).view(np.uint32), Xi.view(np.uint32))
    with capi.Context(top, itp, device=-1) as c:
        c.open_trr(trr, traj)
        assert c.frame_count() == frames
        assert np.array_equal(c.groups(), groups)
        assert c.keys() == keys
        rb = 8 if VARIANTS[variant]["double"] else 4
        assert c.layout()["line_bytes"] == 3 * rb
        raw = open(trr, "rb").read()
        for t in range(frames):
            off = c.frame_atoms_offset(t)
            x = np.frombuffer(raw, ">f8" if rb == 8 else ">f4", 3 * s.n_atoms, off).astype(np.float32)
            assert np.array_equal(x.view(np.uint32), fr[t]["x"].reshape(-1).view(np.uint32))
        with pytest.raises(capi.CgcError) as e:                       # no CPU fallback here either
            c.run(0, 1)
        assert e.value.code == capi.CGC_ERR_CUDA


def test_trr_errors_on_cpu(small_inputs, tmp_path):
    traj, top, itp, frames = small_inputs("tiny")
    s, _ = _system("tiny")
    good = synth.trr_frame_bytes(s, 0)
    with capi.Context(top, itp, device=-1) as c:
        bad = bytearray(good)
        bad[3] ^= 0x01                                                # magic number
        p = tmp_path / "bad_magic.trr"
        p.write_bytes(bytes(bad))
        with pytest.raises(capi.CgcError) as e:
            c.open_trr(str(p), traj)
        assert e.value.code == capi.CGC_ERR_INPUT_MALFORMED
        other = synth.make_system(1900, 8, 42, "shared")              # different atom count than the structure file
        p = tmp_path / "other.trr"
        p.write_bytes(synth.trr_frame_bytes(other, 0))
        with pytest.raises(capi.CgcError) as e:
            c.open_trr(str(p), traj)
        assert e.value.code == capi.CGC_ERR_INPUT_MALFORMED
        p = tmp_path / "cut.trr"                                      # second frame cut short: one complete frame
        p.write_bytes(good + synth.trr_frame_bytes(s, 1)[:5000])
        c.open_trr(str(p), traj)
        assert c.frame_count() == 1
        with pytest.raises(capi.CgcError) as e:
            c.open_trr(str(tmp_path / "missing.trr"), traj)
        assert e.value.code == capi.CGC_ERR_READ


@pytest.mark.gpu
@pytest.mark.parametrize("variant", sorted(VARIANTS))
@pytest.mark.parametrize("tag", ["tiny", "permol"])
def test_trr_dE_identical_to_gro_path(tag, variant, small_inputs, oracle_mod, tmp_path):
    O = oracle_mod
    traj, top, itp, frames = small_inputs(tag)
    s, _ = _system(tag)
    trr = str(tmp_path / "t.trr")
    synth.write_trr(trr, s, frames, **VARIANTS[variant])
    with capi.Context(top, itp, 0) as c:
        c.open(traj)
        dE_gro, done, rc = c.run(0, frames)
        assert done == frames
        mean_g, cov_g = c.cov_finalize(0)
    with capi.Context(top, itp, 0) as c:
        c.open_trr(trr, traj)
        dE_trr, done, rc = c.run(0, frames + 3)                       # over-asking stops cleanly
        assert rc == capi.CGC_EOF and done == frames
        dE_trr = dE_trr[:frames]
        mean_t, cov_t = c.cov_finalize(0)
        # same coordinates, same kernels, exact bin adds: the rows are bit-identical to the .gro path's
        ref = O.run_trr(trr, traj, top, itp, frames)
        e, z = scaled_err(dE_trr, ref["dE_f64"], ref["scale"])
        assert e <= 1e-6 and z == 0.0
        assert np.array_equal(dE_trr.view(np.uint32), dE_gro.view(np.uint32))
        assert np.allclose(cov_t, cov_g, rtol=1e-6, atol=0)
    # one frame from the middle of the file, through memory
    raw = np.fromfile(trr, dtype=np.uint8)
    with capi.Context(top, itp, 0) as c:
        c.open_memory_trr(raw, traj)
        one, done, rc = c.run(1, 1)
        assert done == 1
        assert np.array_equal(one[0].view(np.uint32), dE_gro[1].view(np.uint32))


@pytest.mark.gpu
def test_cli_reads_trr(small_inputs, tmp_path):
    from cgromcorrfmo_b200 import build
    traj, top, itp, frames = small_inputs("golden7")
    s, _ = _system("golden7")
    trr = str(tmp_path / "t.trr")
    synth.write_trr(trr, s, frames, double=False, velocities=True)
    out_g, out_t = str(tmp_path / "g"), str(tmp_path / "t")
    r = subprocess.run([build.CLI, traj, top, itp, out_g, str(frames)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([build.CLI, trr, top, itp, out_t, str(frames), "--structure", traj], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    a = np.array([[float(v) for v in ln.split(",")] for ln in open(out_g + ".time").read().strip().split("\n")])
    b = np.array([[float(v) for v in ln.split(",")] for ln in open(out_t + ".time").read().strip().split("\n")])
    assert a.shape == b.shape
    assert np.allclose(a, b, rtol=2e-5, atol=1e-4)                    # 6 printed digits
    r = subprocess.run([build.CLI, trr, top, itp, out_t, str(frames)], capture_output=True, text=True)
    assert r.returncode == 1 and "--structure" in r.stderr


# ----------------------------------------------------------------------------------------------------------------------
# Files NOT written by this repo's writer: tests/golden/trr_{single,double}.trr, assembled field by field with the standard
# library's XDR encoder from the GROMACS trrio layout (tests/golden/make_trr_fixture.py states the provenance: no gmx build
# exists offline, so this is an independent encoder, not gmx output).  Single precision with box + x + v + f (today's
# mdrun), double precision with box + virial + pressure + x + v (GROMACS 3.x/4.0 style), steps 5000.., lambda 0.25.
# ----------------------------------------------------------------------------------------------------------------------
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _trr_fixture(tmp_path):
    g = np.load(os.path.join(GOLD, "trr_fixture.npz"))
    n_atoms, n_bcl, seed = (int(v) for v in g["system"])
    s = synth.make_system(n_atoms, n_bcl, seed, "shared")
    top, itp, gro = str(tmp_path / "topol.top"), str(tmp_path / "bcx.itp"), str(tmp_path / "structure.gro")
    synth.write_top(top, s)
    synth.write_itp(itp, s)
    open(gro, "wb").write(g["structure_gro"].tobytes())
    return g, s, top, itp, gro


@pytest.mark.parametrize("prec", ["single", "double"])
def test_independent_trr_fixture_front_end(prec, oracle_mod, tmp_path):
    """The oracle's reader and the product's host front end on the independently encoded files: frame count, steps, times,
    precision, and the coordinate array found at the offset the header arithmetic gives."""
    O = oracle_mod
    g, s, top, itp, gro = _trr_fixture(tmp_path)
    trr = os.path.join(GOLD, "trr_%s.trr" % prec)
    want = g["x"] if prec == "double" else g["x"].astype(np.float32).astype(np.float64)
    fr = O.read_trr(trr)
    assert [f["step"] for f in fr] == list(g["steps"]) and len(fr) == 3
    assert np.allclose([f["time"] for f in fr], g["times"], rtol=1e-6)
    for k, f in enumerate(fr):
        assert f["natoms"] == s.n_atoms and f["real_bytes"] == (8 if prec == "double" else 4)
        assert np.array_equal(f["x"].view(np.uint32), want[k].astype(np.float32).view(np.uint32))
    raw = open(trr, "rb").read()
    with capi.Context(top, itp, device=-1) as c:
        c.open_trr(trr, gro)
        assert c.frame_count() == 3
        rb = 8 if prec == "double" else 4
        assert c.layout()["line_bytes"] == 3 * rb
        for k in range(3):
            x = np.frombuffer(raw, ">f8" if rb == 8 else ">f4", 3 * s.n_atoms, c.frame_atoms_offset(k)).astype(np.float64)
            assert np.array_equal(x, want[k].reshape(-1))


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["single", "double"])
def test_independent_trr_fixture_on_gpu(prec, oracle_mod, tmp_path):
    """K1's .trr modes on the independently encoded files: decoded coordinates bit-equal to the numbers that were encoded
    (double rounded once to float32), dE rows within tolerance of the float64 restatement fed those coordinates."""
    import torch
    O = oracle_mod
    g, s, top, itp, gro = _trr_fixture(tmp_path)
    trr = os.path.join(GOLD, "trr_%s.trr" % prec)
    want = (g["x"] if prec == "double" else g["x"].astype(np.float32)).astype(np.float32)
    raw = np.fromfile(trr, dtype=np.uint8)
    with capi.Context(top, itp, 0) as c:
        c.open_trr(trr, gro)
        padded = np.concatenate([raw, np.zeros(((-raw.size) % 16) + 16, np.uint8)])
        d_text = torch.from_numpy(padded).cuda()
        xyz = torch.empty((3, s.n_atoms, 3), dtype=torch.float32, device="cuda")
        c.decode_device(d_text.data_ptr(), raw.size, [c.frame_atoms_offset(k) for k in range(3)], 3, xyz.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(xyz.cpu().numpy().view(np.uint32), want.view(np.uint32))
        dE, done, rc = c.run(0, 3)
        assert done == 3
        groups = c.groups()
    names, masses, charges = O.load_tables(top, itp)
    index = {}
    for j, nm in enumerate(names):
        index.setdefault(nm, j)
    _, _, keys, grp, _ = O.read_frame(gro, None)
    assert np.array_equal(groups, grp)
    for k in range(3):
        Xi = want[k].reshape(-1)
        for site in (1, 2):
            q = O.atom_data_lookup(keys, grp, site, names, masses, charges, index)[2]
            r64, sc = O.compute_cdc_f64(site, Xi, q, grp)
            scale = np.maximum(np.abs(r64), sc)
            m = scale > 0
            assert float((np.abs(dE[k, site - 1].astype(np.float64) - r64)[m] / scale[m]).max()) <= 1e-6

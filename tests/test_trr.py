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
        # same coordinates, same kernels.  The FP64 bins are filled by atomics whose order is not fixed, so the last
        # bit of a float32 row may differ between any two runs: compare on the parity scale, far below the tolerance
        ref = O.run_trr(trr, traj, top, itp, frames)
        e, z = scaled_err(dE_trr, ref["dE_f64"], ref["scale"])
        assert e <= 1e-6 and z == 0.0
        d = np.abs(dE_trr.astype(np.float64) - dE_gro.astype(np.float64))
        assert np.all(d <= 1e-7 * np.maximum(np.abs(ref["dE_f64"]), ref["scale"]))
        assert np.allclose(cov_t, cov_g, rtol=1e-6, atol=0)
    # one frame from the middle of the file, through memory
    raw = np.fromfile(trr, dtype=np.uint8)
    with capi.Context(top, itp, 0) as c:
        c.open_memory_trr(raw, traj)
        one, done, rc = c.run(1, 1)
        assert done == 1
        d = np.abs(one[0].astype(np.float64) - dE_gro[1].astype(np.float64))
        assert np.all(d <= 1e-7 * np.maximum(np.abs(ref["dE_f64"][1]), ref["scale"][1]))


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

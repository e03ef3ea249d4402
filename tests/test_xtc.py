"""GROMACS .xtc input ("a binary trr/xtc reader"), decoded on the device (csrc/xtc_kernel.cu).

The reference has no .xtc reader (its authors convert to .gro text with trjconv first, reference
scripts/2014-09-trr2dEcsv.sh:22), and no GROMACS exists offline, so parity is anchored three ways:
  * the oracle's reader (oracle.read_xtc, a restatement of the published xdr3dfcoord routine) against files written by an
    independent pure-Python encoder (tests/golden/xtc_*.xtc, make_xtc_fixture.py) and against the library's C++ writer;
  * the device kernels against the oracle — on the CPU through tests/xtc_emulation.cpp, which runs the kernels' own
    per-lane / per-thread source (csrc/xtc_bits.h) lane by lane, on the GPU (-m gpu) through the C-ABI;
  * the .gro path: an .xtc of precision 1000 holding the integers the "%8.3f" text prints gives bit-identical dE rows.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from cgromcorrfmo_b200 import capi, synth
from conftest import SMALL_SYSTEMS, scaled_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
EMU_SRC = os.path.join(ROOT, "tests", "xtc_emulation.cpp")


def _system(tag):
    cfg = SMALL_SYSTEMS[tag]
    return synth.make_system(cfg["n_atoms"], cfg["n_bcl"], cfg["seed"], cfg["solvent"]), cfg["frames"]


@pytest.fixture(scope="session")
def emu(tmp_path_factory):
    """The kernels' source compiled for the host (g++), loaded with ctypes."""
    d = tmp_path_factory.mktemp("xtcemu")
    so = str(d / "libxtcemu.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall", "-o", so, EMU_SRC])
    lib = ctypes.CDLL(so)
    lib.emu_xtc_decode.restype = ctypes.c_int
    lib.emu_xtc_decode.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                   ctypes.c_int, ctypes.c_void_p]
    lib.emu_value_of_stream_bits.restype = ctypes.c_uint32
    lib.emu_value_of_stream_bits.argtypes = [ctypes.c_uint32, ctypes.c_int]
    lib.emu_divmod64.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint32)]
    lib.emu_divmod32.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32)]
    lib.emu_divmod96.argtypes = [ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint32), ctypes.c_uint32,
                                 ctypes.POINTER(ctypes.c_uint32)]
    lib.emu_to_coord.restype = ctypes.c_float
    lib.emu_to_coord.argtypes = [ctypes.c_int, ctypes.c_float]
    return lib


def frame_offsets(blob: bytes):
    """(coordinate-block offset, end) of every frame of an .xtc held in memory — plain header arithmetic."""
    out, pos = [], 0
    while pos < len(blob):
        assert int.from_bytes(blob[pos:pos + 4], "big") == 1995
        nbytes = int.from_bytes(blob[pos + 88:pos + 92], "big")
        end = pos + 92 + ((nbytes + 3) & ~3)
        out.append((pos + 52, end))
        pos = end
    return out


def emu_decode(lib, blob: bytes, n_atoms: int, scalar: bool = False, max_stream_words: int = 0, n_sites_atoms: int = 37, seed: int = 1,
               text_bytes=None, offsets=None):
    """Run the emulated kernels over every frame of `blob`.  Returns (err flags, xyz [frames, n_atoms, 3] float32, pads ok,
    sites [frames, stride, 4], slot, slot_dq, stats)."""
    offs = offsets if offsets is not None else [o for o, _ in frame_offsets(blob)]
    nf = len(offs)
    n_pad = (n_atoms + 2047) // 2048 * 2048
    rng = np.random.default_rng(seed)
    stride = n_sites_atoms + 3
    slot = np.full(n_pad, -1, np.int32)
    pick = rng.choice(n_atoms, size=min(n_sites_atoms, n_atoms), replace=False)
    slot[pick] = rng.permutation(stride)[:pick.size]
    slot_dq = rng.normal(size=stride).astype(np.float32)
    # the text exactly as long as the file: any read past it is a bug the address sanitizer build would report too
    text = np.frombuffer(blob, np.uint8).copy()
    xyz8 = np.full((nf, n_pad * 3), np.float32(-1.25e30), np.float32)
    sites = np.full((nf, stride, 4), np.float32(3.5e30), np.float32)
    coff = np.asarray(offs, np.int64)
    stats = np.zeros(4, np.int64)
    err = lib.emu_xtc_decode(text.ctypes.data, len(blob) if text_bytes is None else text_bytes, coff.ctypes.data, nf, n_atoms, n_pad,
                             slot.ctypes.data, slot_dq.ctypes.data, stride, xyz8.ctypes.data, sites.ctypes.data, 1 if scalar else 0,
                             max_stream_words, stats.ctypes.data)
    blocks = xyz8.reshape(nf, n_pad // 8, 3, 8)
    xyz = blocks.transpose(0, 1, 3, 2).reshape(nf, n_pad, 3)
    return err, xyz[:, :n_atoms].copy(), xyz[:, n_atoms:].copy(), sites, slot, slot_dq, stats


def expect_coords(ints, precision):
    return (np.asarray(ints, np.float64) / np.float64(np.float32(precision))).astype(np.float32)


def check_against(ints_per_frame, precision, result):
    err, xyz, pads, sites, slot, slot_dq, stats = result
    assert err == 0
    want = np.stack([expect_coords(m, precision) for m in ints_per_frame])
    assert np.array_equal(xyz.view(np.uint32), want.view(np.uint32))
    assert np.all(pads == 0.0)                                        # tile padding zeroed, nothing left unwritten
    n_atoms = want.shape[1]
    used = np.nonzero(slot[:n_atoms] >= 0)[0]
    for f in range(want.shape[0]):
        got = sites[f][slot[used]]
        assert np.array_equal(got[:, :3].view(np.uint32), (-want[f][used]).view(np.uint32))
        assert np.array_equal(got[:, 3], slot_dq[slot[used]])
        untouched = np.setdiff1d(np.arange(sites.shape[1]), slot[used])
        assert np.all(sites[f][untouched] == np.float32(3.5e30))
    return stats


# ----------------------------------------------------------------------------------------------------------------------
# the oracle's reader and the host front end against the independently encoded files
# ----------------------------------------------------------------------------------------------------------------------
def _fixture(tmp_path):
    g = np.load(os.path.join(GOLD, "xtc_fixture.npz"))
    n_atoms, n_bcl, seed = (int(v) for v in g["system"])
    s = synth.make_system(n_atoms, n_bcl, seed, "shared")
    top, itp, gro = str(tmp_path / "topol.top"), str(tmp_path / "bcx.itp"), str(tmp_path / "structure.gro")
    synth.write_top(top, s)
    synth.write_itp(itp, s)
    open(gro, "wb").write(g["structure_gro"].tobytes())
    return g, s, top, itp, gro


def test_oracle_reader_on_the_independent_fixture(oracle_mod):
    O = oracle_mod
    g = np.load(os.path.join(GOLD, "xtc_fixture.npz"))
    fr = O.read_xtc(os.path.join(GOLD, "xtc_water.xtc"))
    assert [f["step"] for f in fr] == list(g["water_steps"]) and np.allclose([f["time"] for f in fr], g["water_times"])
    for k, f in enumerate(fr):
        assert f["precision"] == 1000.0 and np.array_equal(f["ints"], g["water_ints"][k])
        assert np.array_equal(f["x"].view(np.uint32), expect_coords(g["water_ints"][k], 1000.0).view(np.uint32))
    fr = O.read_xtc(os.path.join(GOLD, "xtc_wide.xtc"))
    assert [f["step"] for f in fr] == list(g["wide_steps"])
    for k, f in enumerate(fr):
        assert f["precision"] == 100.0 and np.array_equal(f["ints"], g["wide_ints"][k])


def test_library_writer_against_the_oracle_reader(oracle_mod, tmp_path):
    """cgc_xtc_encode (C++) -> oracle.read_xtc (Python): the integers that went in, for shapes that take every branch of
    the format: isolated atoms, bonded runs, water triples, a range above 2^24 (separate fields), packed numbers above
    64 bits, small differences wider than 32 bits."""
    O = oracle_mod
    rng = np.random.default_rng(5)
    cases = {}
    s, _ = _system("tiny")
    cases["synthetic"] = (synth.frame_milli(s, 0), 1000.0)
    chain = np.cumsum(rng.integers(-60, 61, (700, 3)), axis=0) + 40000
    cases["chain"] = (chain, 1000.0)
    o = rng.integers(0, 30000, (300, 3))
    water = np.stack([o, o + rng.integers(-70, 71, (300, 3)), o + rng.integers(-70, 71, (300, 3))], 1).reshape(-1, 3)
    cases["water"] = (water, 1000.0)
    wide = rng.integers(-20_000_000, 20_000_000, (150, 3))
    wide[5] = wide[4] + [1, -1, 2]
    cases["separate_fields"] = (wide, 100.0)
    big = rng.integers(0, 16_000_000, (120, 3))                       # sizes just under 2^24: the product needs 72 bits
    big[0] = 0
    big[1] = 16_000_000
    big[9] = big[8] + [7, 7, -7]
    cases["packed_72_bits"] = (big, 1000.0)
    far = np.cumsum(rng.choice([-1, 1], (200, 3)) * rng.integers(820, 1100, (200, 3)), axis=0)   # every step > 1625 in |.|_1
    cases["small_above_32_bits"] = (far, 1000.0)
    for name, (ints, prec) in cases.items():
        blob = capi.xtc_encode(ints, step=3, time_ps=1.5, box=np.eye(3) * 9.0, precision=prec)
        p = tmp_path / (name + ".xtc")
        p.write_bytes(blob + blob)
        fr = O.read_xtc(str(p))
        assert len(fr) == 2 and fr[1]["step"] == 3 and fr[0]["time"] == 1.5 and fr[0]["precision"] == prec, name
        assert np.array_equal(fr[0]["ints"], ints) and np.array_equal(fr[1]["ints"], ints), name
    with pytest.raises(capi.CgcError):
        capi.xtc_encode(np.zeros((9, 3), np.int32))                   # frames of <= 9 atoms are stored uncompressed


def test_xtc_front_end_on_cpu(oracle_mod, tmp_path):
    O = oracle_mod
    g, s, top, itp, gro = _fixture(tmp_path)
    path = os.path.join(GOLD, "xtc_water.xtc")
    blob = open(path, "rb").read()
    offs = frame_offsets(blob)
    _, _, keys, groups, _ = O.read_frame(gro, None)
    with capi.Context(top, itp, device=-1) as c:
        c.open_xtc(path, gro)
        assert c.frame_count() == 3 and c.n_atoms == s.n_atoms
        assert np.array_equal(c.groups(), groups) and c.keys() == keys
        assert [c.frame_atoms_offset(k) for k in range(3)] == [o for o, _ in offs]
        assert c.frame_atoms_offset(3) == -1
        with pytest.raises(capi.CgcError) as e:                       # no CPU fallback here either
            c.run(0, 1)
        assert e.value.code == capi.CGC_ERR_CUDA
        c.open_memory_xtc(blob, gro)
        assert c.frame_count() == 3


def test_xtc_errors_on_cpu(small_inputs, tmp_path):
    traj, top, itp, frames = small_inputs("tiny")
    s, _ = _system("tiny")
    good = synth.xtc_frame_bytes(s, 0)
    with capi.Context(top, itp, device=-1) as c:
        bad = bytearray(good)
        bad[3] ^= 0x01                                                # magic number
        p = tmp_path / "bad_magic.xtc"
        p.write_bytes(bytes(bad))
        with pytest.raises(capi.CgcError) as e:
            c.open_xtc(str(p), traj)
        assert e.value.code == capi.CGC_ERR_INPUT_MALFORMED
        bad = bytearray(good)
        bad[84:88] = (5).to_bytes(4, "big")                           # table index below the first usable entry
        p = tmp_path / "bad_idx.xtc"
        p.write_bytes(bytes(bad))
        with pytest.raises(capi.CgcError) as e:
            c.open_xtc(str(p), traj)
        assert e.value.code == capi.CGC_ERR_INPUT_MALFORMED
        other = synth.make_system(1900, 8, 42, "shared")              # different atom count than the structure file
        p = tmp_path / "other.xtc"
        p.write_bytes(synth.xtc_frame_bytes(other, 0))
        with pytest.raises(capi.CgcError) as e:
            c.open_xtc(str(p), traj)
        assert e.value.code == capi.CGC_ERR_INPUT_MALFORMED
        p = tmp_path / "cut.xtc"                                      # second frame cut short: one complete frame
        p.write_bytes(good + synth.xtc_frame_bytes(s, 1)[:3000])
        c.open_xtc(str(p), traj)
        assert c.frame_count() == 1
        with pytest.raises(capi.CgcError) as e:
            c.open_xtc(str(tmp_path / "missing.xtc"), traj)
        assert e.value.code == capi.CGC_ERR_READ


# ----------------------------------------------------------------------------------------------------------------------
# the kernels' own source, run lane by lane on the host
# ----------------------------------------------------------------------------------------------------------------------
def test_emulated_kernels_on_the_independent_fixture(emu):
    g = np.load(os.path.join(GOLD, "xtc_fixture.npz"))
    for name, key, prec in (("xtc_water.xtc", "water_ints", 1000.0), ("xtc_wide.xtc", "wide_ints", 100.0)):
        blob = open(os.path.join(GOLD, name), "rb").read()
        n_atoms = g[key].shape[1]
        for scalar in (False, True):
            check_against(g[key], prec, emu_decode(emu, blob, n_atoms, scalar=scalar))


@pytest.mark.parametrize("tag", ["tiny", "permol", "nosolv"])
def test_emulated_kernels_on_synthetic_systems(tag, emu):
    s, frames = _system(tag)
    blob = b"".join(synth.xtc_frame_bytes(s, t) for t in range(frames))
    ints = [synth.frame_milli(s, t) for t in range(frames)]
    st_warp = check_against(ints, 1000.0, emu_decode(emu, blob, s.n_atoms))
    st_scalar = check_against(ints, 1000.0, emu_decode(emu, blob, s.n_atoms, scalar=True))
    check_against(ints, 1000.0, emu_decode(emu, blob, s.n_atoms, max_stream_words=64))   # stream read from "global memory"
    # the speculative walk confirms many groups per step where the flags are clear
    assert st_warp[0] * 4 < st_scalar[0]
    assert st_warp[3] == st_scalar[3]                                 # same checkpoints either way


def _shape_cases():
    rng = np.random.default_rng(11)
    out = {}
    # several 4096-atom windows, groups straddling the window and the 16-atom boundaries, an atom count that is no multiple of 16
    n = 3 * 4096 + 1237
    pts = rng.integers(0, 60000, (n, 3))
    k = 0
    while k + 9 < n:                                                  # bonded runs of random length between isolated atoms
        ln = int(rng.integers(1, 10))
        pts[k + 1:k + ln] = pts[k] + np.cumsum(rng.integers(-40, 41, (ln - 1, 3)), axis=0)
        k += ln + int(rng.integers(0, 4))
    out["windows"] = (pts, 1000.0)
    o = rng.integers(0, 30000, (3000, 3))
    water = np.stack([o, o + rng.integers(-95, 96, (3000, 3)), o + rng.integers(-95, 96, (3000, 3))], 1).reshape(-1, 3)
    out["water"] = (water, 1000.0)
    out["exact_window"] = (rng.integers(-5000, 5000, (4096, 3)), 1000.0)     # no tile padding at all
    out["window_plus_one"] = (rng.integers(-5000, 5000, (4097, 3)), 500.0)
    tail = rng.integers(0, 9000, (4100, 3))
    tail[4090:] = tail[4089] + np.cumsum(rng.integers(-3, 4, (10, 3)), axis=0)   # the last run starts in window 0 and ends in window 1
    out["run_across_the_last_window"] = (tail, 1000.0)
    wide = rng.integers(-20_000_000, 20_000_000, (900, 3))
    wide[50:58] = wide[49] + np.cumsum(rng.integers(-5, 6, (8, 3)), axis=0)
    out["separate_fields"] = (wide, 100.0)
    big = rng.integers(0, 16_000_000, (700, 3))
    big[0] = 0
    big[1] = 16_000_000
    big[100:105] = big[99] + np.cumsum(rng.integers(-9, 10, (5, 3)), axis=0)
    out["packed_72_bits"] = (big, 1000.0)
    far = np.cumsum(rng.choice([-1, 1], (600, 3)) * rng.integers(820, 1100, (600, 3)), axis=0)
    out["small_above_32_bits"] = (far, 10000.0)
    out["ten_atoms"] = (rng.integers(0, 500, (10, 3)), 1000.0)
    out["negative"] = (rng.integers(-9_000_000, -8_990_000, (333, 3)), 1000.0)
    flat = rng.integers(0, 4000, (500, 3))
    flat[:, 2] = 77                                                   # a coordinate range of one value
    out["flat"] = (flat, 1000.0)
    return out


@pytest.mark.parametrize("case", sorted(_shape_cases()))
def test_emulated_kernels_on_stream_shapes(case, emu, oracle_mod, tmp_path):
    """Every branch of the format and of the kernels' bookkeeping, against the oracle's reader on the same bytes."""
    O = oracle_mod
    ints, prec = _shape_cases()[case]
    blob = capi.xtc_encode(ints, precision=prec) + capi.xtc_encode(ints[::-1].copy(), step=1, precision=prec)
    p = tmp_path / "c.xtc"
    p.write_bytes(blob)
    fr = O.read_xtc(str(p))
    assert np.array_equal(fr[0]["ints"], ints) and np.array_equal(fr[1]["ints"], ints[::-1])
    for scalar, words in ((False, 0), (True, 0), (False, 32)):
        res = emu_decode(emu, blob, ints.shape[0], scalar=scalar, max_stream_words=words)
        check_against([ints, ints[::-1]], prec, res)
        assert np.array_equal(res[1][0].view(np.uint32), fr[0]["x"].view(np.uint32))


def test_emulated_kernels_flag_damaged_streams(emu):
    """Damage never escapes: a wrong atom count, a header that lies, offsets that do not fit and random bit flips all end
    in the error flag or in a (wrong) result — not in a read or write outside the buffers (the loops are bounded, the word
    indices clamped).  tests/xtc_emulation.cpp is also built with -fsanitize=address,undefined for this (next test)."""
    s, _ = _system("tiny")
    good = synth.xtc_frame_bytes(s, 0)
    assert emu_decode(emu, good, s.n_atoms)[0] == 0
    assert emu_decode(emu, good, s.n_atoms + 1)[0] == 8               # the caller's atom count differs
    assert emu_decode(emu, good, s.n_atoms, text_bytes=len(good) - 8)[0] == 8        # stream passes the end of the text
    assert emu_decode(emu, good, s.n_atoms, offsets=[54])[0] == 8     # unaligned block
    assert emu_decode(emu, good, s.n_atoms, offsets=[len(good) - 16])[0] == 8
    rng = np.random.default_rng(3)
    flagged = 0
    for trial in range(300):
        b = bytearray(good)
        for _ in range(int(rng.integers(1, 4))):
            pos = int(rng.integers(56, len(b)))
            b[pos] ^= 1 << int(rng.integers(8))
        for scalar in (False, True):
            flagged += emu_decode(emu, bytes(b), s.n_atoms, scalar=scalar)[0] != 0
    assert flagged > 10                                               # a flipped coordinate bit is just another coordinate; damaged run headers are caught


def test_emulated_kernels_under_the_address_sanitizer(tmp_path):
    """The same source with -fsanitize=address,undefined: intact and damaged streams, every buffer allocated at its exact
    size.  (compute-sanitizer is closed on the GPU pool; this is the memcheck of K1x's index arithmetic.)"""
    exe = str(tmp_path / "xtcemu_asan")
    r = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
                        "-DXTC_EMU_MAIN", "-o", exe, EMU_SRC], capture_output=True, text=True)
    if r.returncode != 0 and "asan" in (r.stderr or "").lower():
        pytest.skip("no address sanitizer runtime in this toolchain")
    assert r.returncode == 0, r.stderr
    s, _ = _system("permol")
    rng = np.random.default_rng(17)
    files = []
    for k, blob in enumerate([synth.xtc_frame_bytes(s, 0) + synth.xtc_frame_bytes(s, 1),
                              capi.xtc_encode(rng.integers(-20_000_000, 20_000_000, (5000, 3)), precision=100.0),
                              open(os.path.join(GOLD, "xtc_water.xtc"), "rb").read()]):
        p = tmp_path / ("f%d.xtc" % k)
        p.write_bytes(blob)
        files.append(str(p))
    for path in files:
        r = subprocess.run([exe, path, "400"], capture_output=True, text=True)
        assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-4000:])
        assert "clean runs ok" in r.stdout


def test_stream_arithmetic_pieces(emu):
    rng = np.random.default_rng(2)
    # bytes least significant first, the last one with the bits that are left
    for rem in range(1, 33):
        for _ in range(50):
            v = int(rng.integers(0, 1 << rem, dtype=np.uint64))
            bits, left, val = "", rem, v
            while left > 8:
                bits += format(val & 0xFF, "08b")
                val >>= 8
                left -= 8
            bits += format(val, "0%db" % left)
            y = int(bits.ljust(32, "0"), 2)
            assert emu.emu_value_of_stream_bits(y, rem) == v
    q64, r32, q32 = ctypes.c_uint64(), ctypes.c_uint32(), ctypes.c_uint32()
    ds = [1, 2, 3, 7, 8, 10, 255, 256, 4096, 5060, 65535, 65536, 16777215, 16777216] + [int(d) for d in rng.integers(1, 1 << 24, 40)]
    ns = [0, 1, (1 << 64) - 1, (1 << 63), (1 << 32) - 1, 1 << 32] + [int(n) for n in rng.integers(0, 1 << 64, 60, dtype=np.uint64)]
    for d in ds:
        for n in ns + [d - 1, d, d + 1, 3 * d - 1, ((1 << 64) - 1) // d * d, ((1 << 64) - 1) // d * d - 1]:
            n &= (1 << 64) - 1
            emu.emu_divmod64(n, d, ctypes.byref(q64), ctypes.byref(r32))
            assert (q64.value, r32.value) == divmod(n, d)
            n3 = n & 0xFFFFFFFF
            emu.emu_divmod32(n3, d, ctypes.byref(q32), ctypes.byref(r32))
            assert (q32.value, r32.value) == divmod(n3, d)
            hi = int(rng.integers(0, 1 << 32))
            lo, h = ctypes.c_uint64(n), ctypes.c_uint32(hi)
            emu.emu_divmod96(ctypes.byref(lo), ctypes.byref(h), d, ctypes.byref(r32))
            big = (hi << 64) | n
            assert ((h.value << 64) | lo.value, r32.value) == divmod(big, d)
    # k / precision rounded once to float32
    ks = np.concatenate([np.arange(-3000, 3000), rng.integers(-(1 << 24), 1 << 24, 4000), rng.integers(-(1 << 31), 1 << 31, 500),
                         [(1 << 24) - 1, 1 << 24, (1 << 24) + 1, -(1 << 31), (1 << 31) - 1]])
    for prec in (1000.0, 100.0, 10000.0, 500.0, 123.0):
        want = (ks.astype(np.float64) / np.float64(np.float32(prec))).astype(np.float32)
        got = np.array([emu.emu_to_coord(int(k), prec) for k in ks], np.float32)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), prec


# ----------------------------------------------------------------------------------------------------------------------
# on the GPU, through the C-ABI
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("scalar_index", [False, True])
@pytest.mark.parametrize("tag", ["tiny", "permol"])
def test_xtc_dE_identical_to_gro_path(tag, scalar_index, small_inputs, oracle_mod, tmp_path, monkeypatch):
    O = oracle_mod
    traj, top, itp, frames = small_inputs(tag)
    s, _ = _system(tag)
    xtc = str(tmp_path / "t.xtc")
    synth.write_xtc(xtc, s, frames)
    if scalar_index:
        monkeypatch.setenv("CGC_XTC_SCALAR_INDEX", "1")
    with capi.Context(top, itp, 0) as c:
        c.open(traj)
        dE_gro, done, rc = c.run(0, frames)
        assert done == frames
        mean_g, cov_g = c.cov_finalize(0)
    with capi.Context(top, itp, 0) as c:
        c.open_xtc(xtc, traj)
        dE_xtc, done, rc = c.run(0, frames + 3)                       # over-asking stops cleanly
        assert rc == capi.CGC_EOF and done == frames
        dE_xtc = dE_xtc[:frames]
        mean_x, cov_x = c.cov_finalize(0)
        ref = O.run_xtc(xtc, traj, top, itp, frames)
        e, z = scaled_err(dE_xtc, ref["dE_f64"], ref["scale"])
        assert e <= 1e-6 and z == 0.0
        # same integers -> same float32 coordinates -> same kernels, exact bin adds: bit-identical rows
        assert np.array_equal(dE_xtc.view(np.uint32), dE_gro.view(np.uint32))
        assert np.allclose(cov_x, cov_g, rtol=1e-6, atol=0)
    raw = np.fromfile(xtc, dtype=np.uint8)
    with capi.Context(top, itp, 0) as c:                              # one frame from the middle of the file, through memory
        c.open_memory_xtc(raw, traj)
        one, done, rc = c.run(1, 1)
        assert done == 1
        assert np.array_equal(one[0].view(np.uint32), dE_gro[1].view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("name,key,prec", [("xtc_water.xtc", "water_ints", 1000.0)])
def test_independent_xtc_fixture_on_gpu(name, key, prec, oracle_mod, tmp_path):
    """K1x on the independently encoded file: decoded coordinates bit-equal to the integers that were encoded, dE rows within
    tolerance of the float64 restatement fed those coordinates."""
    import torch
    O = oracle_mod
    g, s, top, itp, gro = _fixture(tmp_path)
    path = os.path.join(GOLD, name)
    want = np.stack([expect_coords(m, prec) for m in g[key]])
    raw = np.fromfile(path, dtype=np.uint8)
    nf = want.shape[0]
    with capi.Context(top, itp, 0) as c:
        c.open_xtc(path, gro)
        padded = np.concatenate([raw, np.zeros(((-raw.size) % 16) + 16, np.uint8)])
        d_text = torch.from_numpy(padded).cuda()
        xyz = torch.empty((nf, s.n_atoms, 3), dtype=torch.float32, device="cuda")
        c.decode_device(d_text.data_ptr(), raw.size, [c.frame_atoms_offset(k) for k in range(nf)], nf, xyz.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(xyz.cpu().numpy().view(np.uint32), want.view(np.uint32))
        dE, done, rc = c.run(0, nf)
        assert done == nf
        groups = c.groups()
    names, masses, charges = O.load_tables(top, itp)
    index = {}
    for j, nm in enumerate(names):
        index.setdefault(nm, j)
    _, _, keys, grp, _ = O.read_frame(gro, None)
    assert np.array_equal(groups, grp)
    for k in range(nf):
        Xi = want[k].reshape(-1)
        for site in (1, 2):
            q = O.atom_data_lookup(keys, grp, site, names, masses, charges, index)[2]
            r64, sc = O.compute_cdc_f64(site, Xi, q, grp)
            scale = np.maximum(np.abs(r64), sc)
            m = scale > 0
            assert float((np.abs(dE[k, site - 1].astype(np.float64) - r64)[m] / scale[m]).max()) <= 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(_shape_cases()))
def test_stream_shapes_on_gpu(case, tmp_path):
    """The stream shapes of the CPU emulation test through the real kernels (cgc_decode_device), against the integers that
    were encoded.  The structure file is a synthetic system of the same atom count (only the coordinates matter here)."""
    import torch
    ints, prec = _shape_cases()[case]
    n = ints.shape[0]
    if n < 300:
        pytest.skip("needs a system with at least one chromophore")
    s = synth.make_system(n, 1, 5, "shared")
    traj, top, itp = synth.write_inputs(str(tmp_path), s, 1)
    blob = capi.xtc_encode(ints, precision=prec) + capi.xtc_encode(ints[::-1].copy(), step=1, precision=prec)
    want = np.stack([expect_coords(ints, prec), expect_coords(ints[::-1], prec)])
    raw = np.frombuffer(blob, np.uint8)
    for scalar in ("0", "1"):
        os.environ["CGC_XTC_SCALAR_INDEX"] = scalar
        try:
            with capi.Context(top, itp, 0) as c:
                c.open_memory_xtc(raw, traj)
                assert c.frame_count() == 2
                padded = np.concatenate([raw, np.zeros(((-raw.size) % 16) + 16, np.uint8)])
                d_text = torch.from_numpy(padded).cuda()
                xyz = torch.empty((2, n, 3), dtype=torch.float32, device="cuda")
                c.decode_device(d_text.data_ptr(), raw.size, [c.frame_atoms_offset(k) for k in range(2)], 2, xyz.data_ptr())
                torch.cuda.synchronize()
                assert c.decode_errors() == 0
                assert np.array_equal(xyz.cpu().numpy().view(np.uint32), want.view(np.uint32))
        finally:
            os.environ.pop("CGC_XTC_SCALAR_INDEX", None)


@pytest.mark.gpu
def test_damaged_xtc_raises_on_gpu(small_inputs, tmp_path):
    traj, top, itp, frames = small_inputs("tiny")
    s, _ = _system("tiny")
    good = synth.xtc_frame_bytes(s, 0)
    b = bytearray(good + synth.xtc_frame_bytes(s, 1))
    # a run length the stream cannot hold: set the first flag bit's five bits to the maximum over and over
    for pos in range(92 + 40, 92 + 400):
        b[pos] = 0xFF
    p = tmp_path / "damaged.xtc"
    p.write_bytes(bytes(b))
    with capi.Context(top, itp, 0) as c:
        c.open_xtc(str(p), traj)
        with pytest.raises(capi.CgcError) as e:
            c.run(0, 2)
        assert e.value.code == capi.CGC_ERR_INPUT_MALFORMED
        c.open_xtc(str(tmp_path / "damaged.xtc"), traj)               # the context stays usable
    p2 = tmp_path / "good.xtc"
    p2.write_bytes(good)
    with capi.Context(top, itp, 0) as c:
        c.open_xtc(str(p2), traj)
        dE, done, rc = c.run(0, 1)
        assert done == 1


@pytest.mark.gpu
def test_cli_reads_xtc(small_inputs, tmp_path):
    from cgromcorrfmo_b200 import build
    traj, top, itp, frames = small_inputs("golden7")
    s, _ = _system("golden7")
    xtc = str(tmp_path / "t.xtc")
    synth.write_xtc(xtc, s, frames)
    out_g, out_x = str(tmp_path / "g"), str(tmp_path / "x")
    r = subprocess.run([build.CLI, traj, top, itp, out_g, str(frames)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([build.CLI, xtc, top, itp, out_x, str(frames), "--structure", traj], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(out_g + ".time", "rb").read() == open(out_x + ".time", "rb").read()   # bit-identical rows -> identical text
    a = np.array([[float(v) for v in ln.split(",")] for ln in open(out_g + ".mtx").read().strip().split("\n")])
    b = np.array([[float(v) for v in ln.split(",")] for ln in open(out_x + ".mtx").read().strip().split("\n")])
    assert a.shape == b.shape and np.allclose(a, b, rtol=1e-5, atol=1e-12)
    r = subprocess.run([build.CLI, xtc, top, itp, out_x, str(frames)], capture_output=True, text=True)
    assert r.returncode == 1 and "--structure" in r.stderr


@pytest.mark.gpu
def test_python_driver_reads_xtc(small_inputs, tmp_path):
    """python -m cgromcorrfmo_b200.run with --structure: the .xtc run writes the same .time text as the .gro run."""
    from cgromcorrfmo_b200 import run
    traj, top, itp, frames = small_inputs("tiny")
    s, _ = _system("tiny")
    xtc = str(tmp_path / "t.xtc")
    synth.write_xtc(xtc, s, frames)
    out_g, out_x = str(tmp_path / "g"), str(tmp_path / "x")
    assert run.traj2dEcsv(traj, top, itp, out_g, frames) == frames
    assert run.traj2dEcsv(xtc, top, itp, out_x, frames + 5, structure=traj) == frames
    assert open(out_g + ".time", "rb").read() == open(out_x + ".time", "rb").read()
    with pytest.raises(capi.CgcError):
        run.traj2dEcsv(xtc, top, itp, out_x, frames)                  # no structure file


def _random_stream(rng):
    """A random system: stretches of isolated atoms, bonded chains, small clusters and exact repeats, at a random scale."""
    n = int(rng.choice([10, 11, 15, 16, 17, 31, 33, 100, 255, 256, 257, 1000, 4095, 4096, 4097, 4111, 4112, 4113, 8192, 8200,
                        int(rng.integers(10, 9000))]))
    scale = int(rng.choice([50, 2000, 60000, 3_000_000, 17_000_000, 400_000_000]))
    pts = np.empty((n, 3), np.int64)
    k = 0
    while k < n:
        kind = int(rng.integers(0, 5))
        ln = min(n - k, int(rng.integers(1, 40)))
        base = rng.integers(-scale, scale, 3)
        if kind == 0:                                                 # isolated atoms
            pts[k:k + ln] = rng.integers(-scale, scale, (ln, 3))
        elif kind == 1:                                               # a bonded chain
            step = int(rng.choice([3, 40, 150, 900, 5000]))
            pts[k:k + ln] = base + np.cumsum(rng.integers(-step, step + 1, (ln, 3)), axis=0)
        elif kind == 2:                                               # a cluster around one point
            r = int(rng.choice([2, 30, 120, 700]))
            pts[k:k + ln] = base + rng.integers(-r, r + 1, (ln, 3))
        elif kind == 3:                                               # the same point again and again
            pts[k:k + ln] = base
        else:                                                         # triples: one atom and two close companions
            for j in range(k, k + ln):
                pts[j] = base + rng.integers(-110, 111, 3) if (j - k) % 3 else rng.integers(-scale, scale, 3)
                if (j - k) % 3 == 0:
                    base = pts[j]
        k += ln
    pts = np.clip(pts, -(1 << 31) + 1000, (1 << 31) - 1000)
    prec = float(rng.choice([1000.0, 1000.0, 100.0, 10000.0, 10.0]))
    return pts, prec


def test_emulated_kernels_on_random_streams(emu, oracle_mod, tmp_path):
    """A hundred and fifty random systems through the library's writer, the oracle's reader and the emulated kernels (both index walks,
    staged and unstaged stream): the integers that went in come out, bit for bit, every time."""
    O = oracle_mod
    rng = np.random.default_rng(20241020)
    for case in range(150):
        ints, prec = _random_stream(rng)
        blob = capi.xtc_encode(ints, step=case, precision=prec)
        if case % 6 == 0:                                             # the reader in Python is slow: a sample of the cases
            p = tmp_path / "r.xtc"
            p.write_bytes(blob)
            fr = O.read_xtc(str(p))
            assert len(fr) == 1 and np.array_equal(fr[0]["ints"], ints), case
        for scalar, words in ((False, 0), (True, 0), (False, 48)):
            check_against([ints], prec, emu_decode(emu, blob, ints.shape[0], scalar=scalar, max_stream_words=words, seed=case))


@pytest.mark.parametrize("tag", ["tiny", "permol"])
def test_gro2xtc_converter(tag, small_inputs, oracle_mod, tmp_path):
    """python -m cgromcorrfmo_b200.gro2xtc: the .xtc holds exactly the integers the .gro text prints (45- and 69-byte lines),
    with the frames' time stamps and box; the oracle's reader and the library's front end agree on the result."""
    from cgromcorrfmo_b200 import gro2xtc
    O = oracle_mod
    traj, top, itp, frames = small_inputs(tag)
    s, _ = _system(tag)
    xtc = str(tmp_path / "conv.xtc")
    assert gro2xtc.convert(traj, xtc) == frames
    fr = O.read_xtc(xtc)
    assert len(fr) == frames
    for t, f in enumerate(fr):
        assert np.array_equal(f["ints"], synth.frame_milli(s, t))
        assert f["step"] == t and abs(f["time"] - 0.2 * t) < 1e-6 and f["precision"] == 1000.0
    assert gro2xtc.convert(traj, str(tmp_path / "p100.xtc"), precision=100.0) == frames
    f100 = O.read_xtc(str(tmp_path / "p100.xtc"))[0]
    assert np.array_equal(f100["ints"], np.round(synth.frame_milli(s, 0) / 10.0).astype(np.int64))
    with capi.Context(top, itp, device=-1) as c:
        c.open_xtc(xtc, traj)
        assert c.frame_count() == frames


def test_reference_fixture_as_xtc(ref_fixture, emu, oracle_mod, tmp_path):
    """The reference's own test trajectory (src/test/test.gro: a real MD structure of the FMO protein, 6383 atoms, 69-byte
    lines) converted with gro2xtc: the oracle's .xtc reader and the emulated kernels return, bit for bit, the float32
    coordinates the reference's reader (ReadOneTimeGrom2Atom, reference src/grom2atomFMO.cpp:448-450) takes from the text.
    Real bonded structure is what the stream looks like in practice: ~7.6 atoms per group, nine groups in ten flagged."""
    from cgromcorrfmo_b200 import gro2xtc
    O = oracle_mod
    gro, top, itp = ref_fixture
    xtc = str(tmp_path / "ref.xtc")
    frames = gro2xtc.convert(gro, xtc)
    assert frames == 2
    fr = O.read_xtc(xtc)
    pos = None
    want = []
    for t in range(frames):
        pos, Xi, keys, groups, _ = O.read_frame(gro, pos)
        assert np.array_equal(fr[t]["x"].reshape(-1).view(np.uint32), Xi.view(np.uint32))
        want.append(Xi.reshape(-1, 3))
    blob = open(xtc, "rb").read()
    for scalar in (False, True):
        err, xyz, pads, sites, slot, slot_dq, stats = emu_decode(emu, blob, fr[0]["natoms"], scalar=scalar)
        assert err == 0 and np.all(pads == 0.0)
        assert np.array_equal(xyz.view(np.uint32), np.stack(want).view(np.uint32))
    with capi.Context(top, itp, device=-1) as c:
        c.open_xtc(xtc, gro)
        assert c.frame_count() == frames and c.n_atoms == fr[0]["natoms"]
        assert np.array_equal(c.groups(), groups)


def test_index_walk_steps_on_rigid_water(emu):
    """What the speculative walk is for: in rigid bulk water (TIP3P geometry, random orientations) the writer settles on one
    table index, every molecule is one clear-flag group of three, and the walk confirms up to 128 groups per step."""
    rng = np.random.default_rng(4)
    nw = 4000
    q = rng.normal(size=(nw, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    a, b, c, d = q.T
    R = np.stack([np.stack([a*a+b*b-c*c-d*d, 2*(b*c-a*d), 2*(b*d+a*c)], 1), np.stack([2*(b*c+a*d), a*a-b*b+c*c-d*d, 2*(c*d-a*b)], 1),
                  np.stack([2*(b*d-a*c), 2*(c*d+a*b), a*a-b*b-c*c+d*d], 1)], 1)
    th = np.deg2rad(104.52) / 2
    h = np.array([[np.sin(th), np.cos(th), 0.0], [-np.sin(th), np.cos(th), 0.0]]) * 95.72
    O_ = rng.uniform(0, 6000, (nw, 3))
    m = np.round(np.stack([O_, O_ + R @ h[0], O_ + R @ h[1]], 1).reshape(-1, 3)).astype(np.int64)
    blob = capi.xtc_encode(m)
    st_warp = check_against([m], 1000.0, emu_decode(emu, blob, m.shape[0]))
    st_scalar = check_against([m], 1000.0, emu_decode(emu, blob, m.shape[0], scalar=True))
    assert st_scalar[0] <= nw + 50                                    # one group per molecule
    assert st_warp[0] * 20 < st_scalar[0]                             # and dozens of them per step of the speculative walk
    assert len(blob) / m.shape[0] < 3.7                               # bytes per atom

"""CPU tests of the product's host side through the C-ABI: the library loads and exports every symbol the header
declares, the text front end agrees with the oracle bit for bit, errors map to the reference's exception classes, and
compute entry points refuse to run without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

from cgromcorrfmo_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "cgromcorr.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(cgc_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = capi.load()
    for name in sorted(declared):
        assert hasattr(lib, name), "libcgromcorr.so does not export " + name
    assert declared == set(capi.SIGNATURES), "capi.SIGNATURES and include/cgromcorr.h disagree"
    assert b"sm_100a" in lib.cgc_version()


@pytest.mark.parametrize("tag", ["tiny", "permol", "nosolv", "golden7"])
def test_front_end_matches_oracle(tag, small_inputs, oracle_mod):
    O = oracle_mod
    traj, top, itp, frames = small_inputs(tag)
    with capi.Context(top, itp, device=-1) as c:
        names, masses, charges = c.table()
        on, om, oq = O.load_tables(top, itp)
        assert names == on
        assert np.array_equal(masses.view(np.uint32), om.view(np.uint32))
        assert np.array_equal(charges.view(np.uint32), oq.view(np.uint32))
        c.open(traj)
        fr = O.read_frame(traj, None)
        assert np.array_equal(c.groups(), fr[3])                      # bit-exact group ids
        assert c.keys() == fr[2]
        n_atoms, n_chromo, n_groups, num_tot = c.dims()
        assert n_groups == int(fr[3].max()) and n_chromo == -int(fr[3].min()) and num_tot == n_groups + n_chromo
        assert c.frame_count() == frames                              # bit-exact frame count
        for site in (1, n_chromo):
            _, _, q, _ = O.atom_data_lookup(fr[2], fr[3], site, on, om, oq)
            assert np.array_equal(c.site_charges(site).view(np.uint32), q.view(np.uint32))
        cols = c.columns()
        g = fr[3]
        expect = np.where(g < 0, -g - 1, n_chromo + g - 1)
        assert np.array_equal(cols, expect)
        # frame offsets: each frame's atom block starts right after its two header lines
        text = open(traj, "rb").read()
        pos = None
        for f in range(frames):
            start = 0 if pos is None else pos
            hdr_end = text.index(b"\n", text.index(b"\n", start) + 1) + 1
            assert c.frame_atoms_offset(f) == hdr_end
            pos = O.read_frame(traj, pos)[0]
        assert c.frame_atoms_offset(frames) == -1


def test_front_end_reference_fixture(ref_fixture, oracle_mod):
    O = oracle_mod
    traj, top, itp = ref_fixture
    with capi.Context(top, itp, device=-1) as c:
        c.open(traj)
        assert c.dims() == (6383, 7, 350, 357)
        assert c.layout() == dict(line_bytes=69, x_col=20, field_width=8, decimals=3)
        fr = O.read_frame(traj, None)
        assert np.array_equal(c.groups(), fr[3]) and c.keys() == fr[2]
        assert c.frame_count() == 2


def test_no_cpu_fallback(small_inputs):
    traj, top, itp, frames = small_inputs("tiny")
    with capi.Context(top, itp, device=-1) as c:
        c.open(traj)
        with pytest.raises(capi.CgcError) as e:
            c.run(0, 1)
        assert e.value.code == capi.CGC_ERR_CUDA
        with pytest.raises(capi.CgcError) as e:
            c.cov_finalize(0)
        assert e.value.code == capi.CGC_ERR_CUDA


def _write(p, s):
    p.write_text(s)
    return str(p)


def test_error_mapping(tmp_path, small_inputs):
    traj, top, itp, frames = small_inputs("tiny")
    # InputFileMalformed: FEP line with a wrong column count (reference src/grom2atomFMO.cpp:149-154)
    bad = _write(tmp_path / "bad.itp", "[ atoms ]\n 1 T 1 BCX MG 1 0.7 24.3 ;\n 2 T 1 BCX CHA 2 ;\n")
    with pytest.raises(capi.CgcError) as e:
        capi.Context(top, bad, device=-1)
    assert e.value.code == capi.CGC_ERR_INPUT_MALFORMED and "8 or 11" in e.value.message
    # InputFileMalformed: topology cannot be opened (:93-95)
    with pytest.raises(capi.CgcError) as e:
        capi.Context(str(tmp_path / "nope.top"), itp, device=-1)
    assert e.value.code == capi.CGC_ERR_INPUT_MALFORMED
    # InputFileMinorMalformed is swallowed like main() does; first entry wins
    dup = _write(tmp_path / "dup.top", "[ atoms ]\n 1 X 1 ALA N 1 0.1 14.0\n 1 X 1 ALA N 1 0.2 14.0\n")
    with capi.Context(dup, itp, device=-1) as c:
        assert c.topology_warnings == 1
        names, masses, charges = c.table()
        assert names[-1] == "ALA,N1" and charges[-1] == np.float32(0.1)
    with capi.Context(top, itp, device=-1) as c:
        # ReadError: open failed (:495-497) / stray line before any header (:486-488)
        with pytest.raises(capi.CgcError) as e:
            c.open(str(tmp_path / "nope.gro"))
        assert e.value.code == capi.CGC_ERR_READ
        with pytest.raises(capi.CgcError) as e:
            c.open(_write(tmp_path / "stray.gro", "hello\n 1\n"))
        assert e.value.code == capi.CGC_ERR_READ
        # InputFileMalformed: count line is not an integer (:376-381)
        with pytest.raises(capi.CgcError) as e:
            c.open(_write(tmp_path / "cnt.gro", "x t= 0\nabc\n"))
        assert e.value.code == capi.CGC_ERR_INPUT_MALFORMED
        # InputFileMalformed: an atom whose key is not in the table (:594-596)
        line = "%5d%-5s%5s%5d%8.3f%8.3f%8.3f\n" % (1, "XXX", "QQ", 1, 1.0, 2.0, 3.0)
        with pytest.raises(capi.CgcError) as e:
            c.open(_write(tmp_path / "key.gro", "x t= 0\n    1\n" + line + "   1.0 1.0 1.0\n"))
        assert e.value.code == capi.CGC_ERR_INPUT_MALFORMED and "XXX,QQ1" in e.value.message
        # Unsupported: whitespace-separated but not fixed-column
        with pytest.raises(capi.CgcError) as e:
            c.open(_write(tmp_path / "free.gro", "x t= 0\n    2\n1BCL MG 1 1.0 2.0 3.0\n1BCL N1 2 1.5 2.5 3.5\n 1 1 1\n"))
        assert e.value.code == capi.CGC_ERR_UNSUPPORTED
        # options need an opened trajectory; bad values are rejected
        with pytest.raises(capi.CgcError) as e:
            c.set_option("sites", 3)
        assert e.value.code == capi.CGC_ERR_STATE
        c.open(traj)
        with pytest.raises(capi.CgcError):
            c.set_option("sites", 99)      # the reference asserts the site exists (src/grom2atomFMO.cpp:632)
        with pytest.raises(capi.CgcError):
            c.set_option("nonsense", 1)


def test_include_and_truncated_trajectory(tmp_path, small_inputs):
    from cgromcorrfmo_b200 import synth
    traj, top, itp, frames = small_inputs("tiny")
    # #include is followed relative to the including file (reference src/grom2atomFMO.cpp:121-125)
    body = open(top).read()
    _write(tmp_path / "atoms.itp", body)
    main_top = _write(tmp_path / "main.top", '; wrapper\n#include "atoms.itp"\n')
    with capi.Context(top, itp, device=-1) as a, capi.Context(main_top, itp, device=-1) as b:
        assert a.table()[0] == b.table()[0]
    # a trajectory cut in the middle of its last frame: the complete frames are counted, no crash
    data = open(traj, "rb").read()
    cut = _write(tmp_path / "cut.gro", "")
    open(cut, "wb").write(data[: len(data) - 1000])
    with capi.Context(top, itp, device=-1) as c:
        c.open(cut)
        assert c.frame_count() == frames - 1


def test_fixed_column_extension_five_digit_serials(tmp_path, small_inputs):
    """Atom serials >= 10000 touch the atom-name column; the reference's tokeniser aborts on such files
    (SURVEY.md section 8c).  The product falls back to strict fixed columns and derives the same groups."""
    traj, top, itp, frames = small_inputs("tiny")
    lines = open(traj).read().split("\n")
    out = []
    for ln in lines:
        if len(ln) == 44 and ln[20:28].strip().replace(".", "").replace("-", "").isdigit():
            ln = ln[:15] + "%5d" % (10000 + int(ln[15:20])) + ln[20:]
        out.append(ln)
    wide = _write(tmp_path / "wide.gro", "\n".join(out))
    with capi.Context(top, itp, device=-1) as a, capi.Context(top, itp, device=-1) as b:
        a.open(traj)
        b.open(wide)
        assert np.array_equal(a.groups(), b.groups()) and a.keys() == b.keys()


def test_time_writer_is_byte_identical_to_printf_g(tmp_path):
    """cgc_write_time formats with a fast "%g" (csrc/fmt_g.h); it must print exactly what std::ostream << float /
    printf("%g") prints (reference src/FMOparse.cpp:151-165).  Python's '%g' is the same correctly-rounded conversion."""
    rng = np.random.default_rng(2)
    n = 400_000
    x = (rng.choice([-1.0, 1.0], n) * 10.0 ** rng.uniform(-14, 7, n) * rng.uniform(0.1, 1, n)).astype(np.float32)
    special = np.array([0.0, -0.0, 1.0, 0.5, 123456.5 / 349.7, 1e-5, 9.999995e-5, 999999.5, 0.1, 1e5, 1e6, 12345.65,
                        2.5e-7, 1e-38, 1e30, 2.8596e-3, 100000.0 / 349.7, 1000000.0 / 349.7, np.inf, -np.inf, np.nan],
                       np.float32)
    x[:special.size] = special
    # values whose 7th digit is exactly 5 (rounding ties of the decimal expansion) and exact powers of ten
    ties = np.array([(k * 10 + 5) / 64.0 for k in range(100000, 100400)], np.float32)
    x[100:100 + ties.size] = ties / np.float32(349.7)
    x = x.reshape(-1, 500)
    path = str(tmp_path / "w.time")
    capi.write_time(path, x)
    with np.errstate(over="ignore", invalid="ignore"):
        cm = (x.astype(np.float64) * 349.7).astype(np.float32)
    expect = "".join(",".join("%g" % float(v) for v in row) + "\n" for row in cm)
    assert open(path).read() == expect
    capi.write_time(path, x[:2], truncate=False)          # append mode (the reference re-opens in append mode per row)
    assert open(path).read() == expect + expect[: len("".join(expect.splitlines(True)[:2]))]

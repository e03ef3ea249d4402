"""`.time` / `.mtx` text back into numbers on the device (K8, cgc_parse_text) against the reference's reader
(scripts/2014-06-csv2hdf5.py:56-61: float() of every token), restated in oracle.csv_rows.  Bit-exact: the arrays are compared
as raw 64-bit patterns."""
import os

import numpy as np
import pytest

from cgromcorrfmo_b200 import capi, pypca

GOLDEN7 = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden7.npz")


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def test_oracle_reader_on_the_reference_programs_own_text(oracle_mod):
    """the stock CLI's .time and .mtx text (tests/golden/golden7.npz) through the restated reader: shape and a few values."""
    g = np.load(GOLDEN7)
    t = oracle_mod.csv_rows(str(g["time_text"]))
    assert t.shape == (3 * 7, t.shape[1]) and t.dtype == np.float64
    first = str(g["time_text"]).split("\n")[0].split(",")
    assert t[0, 1] == float(first[1]) and t[0, -1] == float(first[-1])
    m = oracle_mod.csv_rows(str(g["mtx_text"]))
    assert m.shape == (7 * t.shape[1], t.shape[1])


def test_argument_checks_need_no_device(small_inputs):
    _, top, itp, _ = small_inputs("tiny")
    with capi.Context(top, itp, device=-1) as c:
        assert c.parse_text(b"").shape == (0, 0)
        assert c.parse_text(b" \n\n").shape == (0, 0)
        with pytest.raises(capi.CgcError) as e:      # no CPU fallback: a table-only context cannot parse
            c.parse_text(b"1,2\n3,4\n")
        assert e.value.code == capi.CGC_ERR_CUDA


@pytest.fixture(scope="module")
def gpu_ctx(small_inputs):
    _, top, itp, _ = small_inputs("tiny")
    with capi.Context(top, itp, device=0) as c:
        yield c


@pytest.mark.gpu
def test_gpu_reads_the_reference_programs_text(gpu_ctx, oracle_mod):
    g = np.load(GOLDEN7)
    for key in ("time_text", "mtx_text"):
        text = str(g[key]).encode()
        want = oracle_mod.csv_rows(text)
        got = gpu_ctx.parse_text(text)
        assert got.shape == want.shape
        assert np.array_equal(_bits(got), _bits(want))
        assert np.array_equal(_bits(gpu_ctx.parse_text(text, num_col=want.shape[1])), _bits(want))


@pytest.mark.gpu
@pytest.mark.parametrize("ncol,rows", [(1, 1), (1, 300), (7, 1), (357, 21), (1099, 48), (33, 4097)])
def test_gpu_round_trip_of_written_rows(gpu_ctx, oracle_mod, tmp_path, ncol, rows):
    """rows over 24 decades (fixed and exponent notation, zeros, negative zero, integers) written by the product's `.time`
    writer, read back by the device: identical to float() of every token, through the file-based shim as well."""
    rng = np.random.default_rng(ncol * 10007 + rows)
    x = (rng.normal(size=(rows, ncol)) * 10.0 ** rng.integers(-12, 12, size=(rows, ncol))).astype(np.float32)
    x.flat[:: 13] = 0.0
    x.flat[5:: 29] = -0.0
    x.flat[3:: 31] = np.round(x.flat[3:: 31] * 1e3)
    path = str(tmp_path / "rt.time")
    capi.write_time(path, x / np.float32(349.7))
    text = open(path, "rb").read()
    want = oracle_mod.csv_rows(text)
    got = gpu_ctx.parse_text(text)
    assert got.shape == (rows, ncol)
    assert np.array_equal(_bits(got), _bits(want))
    if rows % 7 == 0:
        E = pypca.load_E_t_ia_csv(gpu_ctx, path, Nsites=7, dtype=np.float64)
        assert E.shape == (rows // 7, 7, ncol) and np.array_equal(_bits(E.reshape(rows, ncol)), _bits(want))
        assert pypca.load_E_t_ia_csv(gpu_ctx, path).dtype == np.float32
        with pytest.raises(ValueError, match="cannot reshape"):
            pypca.load_E_t_ia_csv(gpu_ctx, path, frames_per_file=rows // 7 + 1)


@pytest.mark.gpu
def test_gpu_token_forms_and_host_tokens(gpu_ctx, oracle_mod):
    """every token shape: signs, leading zeros, bare exponents, 15 digits, exponents at the edge of the exact range; and the
    ones the device leaves to strtod (nan, inf, 17 digits, 1e300, denormals, spaces, CR LF) — same bits either way."""
    device_tokens = ["0", "-0", "+5", "007", "1.", ".5", "-.25", "1e5", "1E5", "1e+05", "1e-05", "123456789012345",
                     "0.000000000000000000001", "1e22", "1e-22", "9.99999e+21", "123456e-22", "0e999", "0.0e-999",
                     "2.22507385850720e-3", "4.35e1", "1.17549e-10", "3.40282e+15", "-1.23457e-05", "999999", "1e0"]
    host_tokens = ["nan", "-nan", "inf", "-inf", "12345678901234567", "1e300", "4.9e-324", "1e23", "1e-23", " 7.5", "1.5 ",
                   "0x10", "1_0"]
    for toks in (device_tokens, device_tokens + host_tokens[:9]):
        rows = [toks[i:] + toks[:i] for i in range(len(toks))]      # every token in every column position
        text = ("\n".join(",".join(r) for r in rows) + "\n").encode()
        want = np.array([[float(t) for t in r] for r in rows])
        got = gpu_ctx.parse_text(text)
        assert got.shape == want.shape
        assert np.array_equal(_bits(got)[~np.isnan(want)], _bits(want)[~np.isnan(want)])
        assert np.array_equal(np.isnan(got), np.isnan(want))
    # CR LF line ends and padded tokens go to the host and come back like float()
    got = gpu_ctx.parse_text(b" 1.5 , 2\r\n3,4\r\n")
    assert np.array_equal(got, [[1.5, 2.0], [3.0, 4.0]])
    # no final newline, surrounding blank lines: f.strip() semantics
    assert np.array_equal(gpu_ctx.parse_text(b"\n\n1,2\n3,4"), [[1.0, 2.0], [3.0, 4.0]])
    assert np.array_equal(gpu_ctx.parse_text(b"5"), [[5.0]])


@pytest.mark.gpu
def test_gpu_malformed_text_is_an_error(gpu_ctx):
    for bad, what in ((b"1,2,3\n4,5\n6,7,8\n", "row 1 has 2 values"), (b"1,2\n3,4,5\n", "row 1 has 3 values"),
                      (b"1,2\n\n3,4\n", "row 1"), (b"1,abc\n", "could not convert"), (b"1,,2\n", "could not convert"),
                      (b"1,2\n3,0x10\n", "could not convert"), (b"1,1_0\n", "could not convert")):
        with pytest.raises(capi.CgcError, match=what) as e:
            gpu_ctx.parse_text(bad)
        assert e.value.code == capi.CGC_ERR_INPUT_MALFORMED
    with pytest.raises(capi.CgcError, match="row 0 has 2 values"):
        gpu_ctx.parse_text(b"1,2\n3,4\n", num_col=3)


@pytest.mark.gpu
def test_gpu_chunked_like_whole(gpu_ctx, oracle_mod, monkeypatch):
    """the text cut into many device chunks (at newlines) gives the same array; tokens straddle every span and tile edge."""
    rng = np.random.default_rng(99)
    x = (rng.normal(size=(600, 41)) * 10.0 ** rng.integers(-6, 6, size=(600, 41))).astype(np.float32)
    text = "".join(",".join("%g" % v for v in r) + "\n" for r in x).encode()
    want = oracle_mod.csv_rows(text)
    whole = gpu_ctx.parse_text(text)
    monkeypatch.setenv("CGC_PARSE_CHUNK_BYTES", "5000")
    cut = gpu_ctx.parse_text(text)
    monkeypatch.delenv("CGC_PARSE_CHUNK_BYTES")
    assert np.array_equal(_bits(whole), _bits(want)) and np.array_equal(_bits(cut), _bits(want))
    bad = text.replace(b"\n", b",1\n", 1)
    monkeypatch.setenv("CGC_PARSE_CHUNK_BYTES", "5000")
    with pytest.raises(capi.CgcError, match="row 1 has 41 values, expected 42"):
        gpu_ctx.parse_text(bad)

"""K5, the `.time` text formatted on the device (csrc/format_kernel.cu), against printf's "%g" — which is what the
reference's `stream << float` prints (reference src/FMOparse.cpp:158-165) — and against the host writer."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _g(v: np.float32) -> str:
    return "%g" % float(v)          # CPython's %g is correctly rounded (round-half-even), like glibc's


def _expect(a: np.ndarray) -> bytes:
    return "".join(",".join(_g(v) for v in row) + "\n" for row in a).encode()


@pytest.fixture(scope="module")
def ctx(small_inputs):
    from cgromcorrfmo_b200 import capi
    traj, top, itp, frames = small_inputs("tiny")
    c = capi.Context(top, itp, 0)
    c.open(traj)
    yield c
    c.close()


def test_device_g_matches_printf_on_random_values(ctx):
    rng = np.random.default_rng(5)
    parts = [
        rng.normal(size=20000) * 10.0 ** rng.integers(-12, 12, 20000),          # all decades
        rng.normal(size=20000) * 100.0,                                         # the magnitudes dE has
        np.round(rng.normal(size=5000) * 1000) / 1000.0,                        # short decimals (trailing zeros stripped)
        rng.integers(-2000000, 2000000, 5000).astype(np.float64),               # integers around the %g switch at 1e6
        np.array([0.0, -0.0, 1.0, -1.0, 100000.0, 999999.0, 999999.5, 1000000.0, 1e-4, 9.99999e-5, 1e-5, 0.1, 0.5,
                  123456.0, 1234567.0, 0.000123456, 1e15, 9.9999994e15, 1.00000001e-16, 3.4e-16]),
    ]
    v = np.concatenate(parts).astype(np.float32)
    v = v[np.isfinite(v) & ((np.abs(v) == 0) | ((np.abs(v) >= 1e-16) & (np.abs(v) < 1e16)))]
    v = v[: (v.size // 50) * 50].reshape(-1, 50)
    text, flag = ctx.format_text_device(v)
    assert flag == 0
    assert text == _expect(v)


def test_device_g_ties_are_broken_exactly(ctx):
    """Seven-digit values that are exactly representable sit exactly on a rounding tie of the sixth digit: half to even.
    Their float32 neighbours sit 1 ulp off the tie: the double product a * 10^k cannot tell, the 128-bit compare can."""
    ties = []
    for base in (1234565, 1234575, 1000005, 9999995, 2500005, 7654325):
        for scale in (1.0, 0.5, 0.25, 8.0, 1024.0, 2.0 ** -10, 2.0 ** -20):
            x = np.float32(base * scale)
            ties += [x, np.nextafter(x, np.float32(np.inf)), np.nextafter(x, np.float32(-np.inf)), -x]
    # decimal ties that are NOT exactly representable: 0.1234565 etc. (one side or the other, decided by the exact value)
    ties += [np.float32(s) for s in ("0.1234565", "12.34565", "1234.565", "1.234565e-7", "1.234575e9", "0.9999995", "99999.95")]
    v = np.array(ties, dtype=np.float32)
    v = np.concatenate([v, np.zeros((-v.size) % 8, np.float32)]).reshape(-1, 8)
    text, flag = ctx.format_text_device(v)
    assert flag == 0
    assert text == _expect(v)


def test_unit_conversion_and_unsupported_values(ctx):
    rng = np.random.default_rng(6)
    kcal = (rng.normal(size=(7, 33)) * 0.3).astype(np.float32)
    text, flag = ctx.format_text_device(kcal, to_cm=True)
    cm = (kcal.astype(np.float64) * 349.7).astype(np.float32)                  # reference src/FMOparse.cpp:151
    assert flag == 0 and text == _expect(cm)
    bad = np.array([[1.0, np.nan, 2.0], [np.inf, 1e-30, 1e20]], dtype=np.float32)
    text, flag = ctx.format_text_device(bad)
    assert flag != 0                                                             # the caller formats such a batch on the host
    assert text.split(b"\n")[0].split(b",")[0] == b"1"


@pytest.mark.parametrize("tag", ["tiny", "permol"])
def test_run_write_time_equals_host_writer(tag, small_inputs, tmp_path):
    """cgc_run_write_time: the file written from the device text equals cgc_write_time on the rows of the same run, and
    appending across calls, frame ranges and an over-asked end behaves like cgc_run."""
    from cgromcorrfmo_b200 import capi
    traj, top, itp, frames = small_inputs(tag)
    dev, host = str(tmp_path / "dev.time"), str(tmp_path / "host.time")
    with capi.Context(top, itp, 0) as c:
        c.open(traj)
        c.set_option("batch_frames", 2)
        dE1, done1, rc1 = c.run_write_time(0, 1, dev, truncate=True, want_dE=True)
        dE2, done2, rc2 = c.run_write_time(1, frames + 5, dev, truncate=False, want_dE=True)
        assert done1 == 1 and done2 == frames - 1 and rc2 == capi.CGC_EOF
        rows = np.concatenate([dE1, dE2])
        capi.write_time(host, rows, truncate=True)
        assert open(dev, "rb").read() == open(host, "rb").read()
        # file only (no rows back)
        none, done, _ = c.run_write_time(0, frames, dev, truncate=True, want_dE=False)
        assert none is None and done == frames
        assert len(open(dev, "rb").read().strip().split(b"\n")) == frames * c.n_sites


def test_cli_device_and_host_format_agree(small_inputs, tmp_path):
    from cgromcorrfmo_b200 import build
    traj, top, itp, frames = small_inputs("golden7")
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    r = subprocess.run([build.CLI, traj, top, itp, a, str(frames)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([build.CLI, traj, top, itp, b, str(frames), "--host-format"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ta = np.array([[float(v) for v in ln.split(",")] for ln in open(a + ".time").read().strip().split("\n")])
    tb = np.array([[float(v) for v in ln.split(",")] for ln in open(b + ".time").read().strip().split("\n")])
    assert ta.shape == tb.shape and np.allclose(ta, tb, rtol=3e-5, atol=1e-4)     # two runs: last-bit differences in the rows


@pytest.mark.parametrize("exponent", [0, 19, -10])
def test_device_g_exhaustive_binade(exponent, ctx, oracle_mod):
    """EVERY float32 of a binade (2^23 values, both signs alternating) against glibc's printf("%g"): [1, 2);
    [2^19, 2^20), which holds the switch to exponent notation at 1e6 and 2^19 exact rounding ties (x.5 with seven digits);
    [2^-10, 2^-9), negative-exponent scaling."""
    bits = (np.uint32((127 + exponent) << 23) + np.arange(1 << 23, dtype=np.uint32))
    v = bits.view(np.float32).copy()
    v[1::2] *= -1
    v = v.reshape(-1, 1024)
    want = oracle_mod.format_g_rows(v)
    got, flag = ctx.format_text_device(v)
    assert flag == 0
    assert len(got) == len(want)
    assert got == want

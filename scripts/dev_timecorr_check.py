"""Developer run: K7 (cgc_time_corr) at the scripts' default size — device time, FP64 rate, and numpy beside it.

    python scripts/dev_timecorr_check.py [n] [corr_len]

Algorithmic FLOP of the linear sums: 2 * sum_{tau < L} (n - tau) per pair.  The CPU legs are the reference's own
formulations on the box's host cores: HDFStoreCt_v3's FFT (sidechain_corr.py:239-240) and, on a sample of the lags, the
np.inner loop of HDFStoreCt_v2 (:216-217)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgromcorrfmo_b200 import capi, synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        s = synth.make_system(1500, 7, 7, "shared")
        _, top, itp = synth.write_inputs(d, s, 1)
        rng = np.random.default_rng(1)
        series = rng.normal(size=(16, n))
        series -= series.mean(axis=1, keepdims=True)
        flop_pair = 2.0 * (n * L - L * (L - 1) / 2.0)
        with capi.Context(top, itp, 0) as c:
            c.time_corr(series[:2], [0], [1], min(L, 1000))          # warm-up (context, first launch)
            for n_pairs in (1, 8, 64, 256):
                a = np.arange(n_pairs) % 16
                b = (np.arange(n_pairs) // 16 + a + 1) % 16
                for circ in (False, True):
                    best = None
                    for _ in range(3):
                        t0 = time.perf_counter()
                        out, ms = c.time_corr(series, a, b, L, circular=circ, want_ms=True)
                        wall = time.perf_counter() - t0
                        best = ms if best is None else min(best, ms)
                    print("pairs %4d circular %d: kernels %.3f ms = %.2f TFLOP/s (algorithmic, linear part), call %.1f ms wall"
                          % (n_pairs, circ, best, flop_pair * n_pairs / best / 1e9, wall * 1e3))
            # numpy beside it
            f1, f2 = series[0], series[1]
            t0 = time.perf_counter()
            v3 = np.real(np.fft.ifft(np.fft.fft(f2) * np.conj(np.fft.fft(f1)))[0:L]) / (n - np.arange(L))
            t_fft = time.perf_counter() - t0
            lags = np.arange(0, L, max(1, L // 200))
            t0 = time.perf_counter()
            v2 = np.array([np.inner(f1[0:n - tc], f2[tc:n]) / (n - tc) for tc in lags])
            t_v2 = (time.perf_counter() - t0) * L / len(lags)
            g_lin = c.time_corr(series, [0], [1], L)[0]
            g_circ = c.time_corr(series, [0], [1], L, circular=True)[0]
            scale = np.sqrt(np.dot(f1, f1) * np.dot(f2, f2)) / (n - np.arange(L))
            print("numpy per pair: HDFStoreCt_v3 (FFT) %.1f ms, HDFStoreCt_v2 (np.inner loop, extrapolated from %d lags) %.0f ms"
                  % (t_fft * 1e3, len(lags), t_v2 * 1e3))
            print("max |device - numpy| / Cauchy-Schwarz bound: circular %.2e, linear (sampled lags) %.2e"
                  % (np.max(np.abs(g_circ - v3) / scale), np.max(np.abs(g_lin[lags] - v2) / scale[lags])))


if __name__ == "__main__":
    main()

"""Developer run: K8 (cgc_parse_text) on a C2-shaped `.time` text — device time, GB/s of text, and the reference's reader
(scripts/2014-06-csv2hdf5.py:56-61) beside it on a sample.   python scripts/dev_parse_check.py [frames]"""
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgromcorrfmo_b200 import capi, synth  # noqa: E402


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    S, G = 24, 1099
    rng = np.random.default_rng(3)
    x = (rng.normal(size=(frames * S, G)) * 10.0 ** rng.integers(-3, 2, size=(frames * S, G))).astype(np.float32)
    with tempfile.TemporaryDirectory() as d:
        s = synth.make_system(1500, 7, 7, "shared")
        _, top, itp = synth.write_inputs(d, s, 1)
        path = os.path.join(d, "c2.time")
        capi.write_time(path, x)
        text = np.fromfile(path, dtype=np.uint8)
        print("text %.1f MB, %d rows x %d values (%.2f bytes per value)" % (text.size / 1e6, frames * S, G, text.size / x.size))
        with capi.Context(top, itp, 0) as c:
            c.parse_text(text[: 1 << 20].tobytes().rsplit(b"\n", 1)[0])      # warm-up
            best_ms, best_wall = 1e9, 1e9
            for _ in range(3):
                t0 = time.perf_counter()
                got, ms = c.parse_text(text, want_ms=True)
                best_wall = min(best_wall, time.perf_counter() - t0)
                best_ms = min(best_ms, ms)
            alg = 2.0 * text.size + 8.0 * got.size     # the text is read by both passes, every value written once
            print("kernels %.3f ms: %.1f GB/s of text, %.0f GB/s algorithmic traffic (2 x text + 8 B/value); call %.1f ms wall "
                  "(H2D of the text from pageable memory + D2H of the values into fresh pages)"
                  % (best_ms, text.size / best_ms / 1e6, alg / best_ms / 1e6, best_wall * 1e3))
        sample_rows = 24 * 8
        sample = bytes(text[: 0]) + b"\n".join(text.tobytes().split(b"\n", sample_rows)[:sample_rows]) + b"\n"
        t0 = time.perf_counter()
        f = sample.decode().strip()
        ref = np.array([[float(v) for v in l.split(",")] for l in f.split("\n")])
        t_ref = time.perf_counter() - t0
        print("reference reader (float() per token) on %d rows: %.3f s -> %.1f s for the whole text on one core"
              % (sample_rows, t_ref, t_ref * frames * S / sample_rows))
        assert np.array_equal(ref.view(np.uint64), got[:sample_rows].view(np.uint64))
        print("bit-identical on the sample")


if __name__ == "__main__":
    main()

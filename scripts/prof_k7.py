"""One K7 call for ncu: 64 pairs at the scripts' default size (n 200 000, corr_len 20 000).  python scripts/prof_k7.py [pairs]"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgromcorrfmo_b200 import capi, synth  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
with tempfile.TemporaryDirectory() as d:
    s = synth.make_system(1500, 7, 7, "shared")
    _, top, itp = synth.write_inputs(d, s, 1)
    series = np.random.default_rng(1).normal(size=(16, 200000))
    a = np.arange(pairs) % 16
    b = (np.arange(pairs) // 16 + a + 1) % 16
    with capi.Context(top, itp, 0) as c:
        for _ in range(2):
            out, ms = c.time_corr(series, a, b, 20000, want_ms=True)
        print("K7 %d pairs: %.3f ms" % (pairs, ms))

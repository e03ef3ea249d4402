"""One K5 call (device "%g" formatter) and one K8 call (device parser) on a C2-shaped block of rows, for ncu:
4096 rows x 1099 values.   python scripts/prof_k5_k8.py"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgromcorrfmo_b200 import capi, synth  # noqa: E402

rng = np.random.default_rng(3)
x = (rng.normal(size=(4096, 1099)) * 10.0 ** rng.integers(-3, 2, size=(4096, 1099))).astype(np.float32)
with tempfile.TemporaryDirectory() as d:
    s = synth.make_system(1500, 7, 7, "shared")
    _, top, itp = synth.write_inputs(d, s, 1)
    with capi.Context(top, itp, 0) as c:
        for _ in range(2):
            text, needs_host = c.format_text_device(x, to_cm=True)
            back, ms = c.parse_text(text, want_ms=True)
        ref = (x.astype(np.float64) * 349.7).astype(np.float32)
        print("K5: %d bytes of text, needs_host %s; K8: %.3f ms; round trip within %%g precision: %s"
              % (len(text), needs_host, ms, bool(np.allclose(back, ref, rtol=1e-5, atol=0))))

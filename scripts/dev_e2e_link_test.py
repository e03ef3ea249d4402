"""Developer experiment: what slows the H2D copies of cgc_run down relative to back-to-back copies of the same buffer?"""
import os, sys, time, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from cgromcorrfmo_b200 import capi
F = 256
with tempfile.TemporaryDirectory() as d:
    s, top, itp, blob, frame_bytes = bench.make_inputs("C2", d, F)
    text = np.frombuffer(blob, dtype=np.uint8); n_text = text.size
    host = torch.empty(n_text + 64, dtype=torch.uint8).pin_memory(); host[:n_text] = torch.from_numpy(text.copy()); host[n_text:] = 0
    d_text = host.to("cuda")
    alone, rates = bench.calibrate_h2d(torch, None, 1, host, d_text)
    print("calibration: best single copy %.1f GB/s, sustained %.1f GB/s" % (alone, rates[0]), flush=True)
    out = np.empty((F, 24, 1099), np.float32)
    for label, batch, rows, cov in (("default (33-frame batches, rows out, K3)", 33, True, True), ("no rows out", 33, False, True), ("no K3", 33, True, False),
                                    ("no rows, no K3", 33, False, False), ("64-frame batches", 64, True, True), ("16-frame batches", 16, True, True),
                                    ("128-frame batches", 128, True, True)):
        with capi.Context(top, itp, 0) as c:
            c.open_memory((host.data_ptr(), n_text))
            c.set_option("batch_frames", batch); c.set_option("profile", 1)
            if not cov: c.set_option("covariance", 0)
            for _ in range(3): c.run(0, F, want_dE=rows, out=out if rows else None)
            c.copy_times(reset=True); c.kernel_times(reset=True)
            t0 = time.perf_counter()
            for _ in range(20): c.run(0, F, want_dE=rows, out=out if rows else None)
            dt = time.perf_counter() - t0
            ms, n = c.copy_times(reset=True)
            kt = c.kernel_times(reset=True)
            print("%-42s %6.0f frames/s | link busy %.2f, %.1f GB/s while copying (%d copies) | K2 %.1f ms per step" % (
                label, 20 * F / dt, ms / (dt * 1e3), 20 * n_text / ms / 1e6, n, kt["cdc"][0] / 20), flush=True)

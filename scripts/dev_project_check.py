"""Developer check: K4 mode projection at the C5 shape (24 sites x 1099 groups, all 1099 modes) — device time through
cgc_kernel-style CUDA events around the whole call, against numpy float64 on one site."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cgromcorrfmo_b200 import capi
import bench

T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = "/tmp/dev_c5"; os.makedirs(d, exist_ok=True)
s, top, itp, blob, fb = bench.make_c2(d, 1)
with capi.Context(top, itp, 0) as c:
    c.open_memory(blob)
    S, G = c.n_sites, c.num_tot
    rng = np.random.default_rng(3)
    X = (rng.normal(size=(T, S, G)) * 0.05 + 20.0).astype(np.float32)
    mean = X.mean(axis=0, dtype=np.float64)
    W = rng.normal(size=(S, G, G))
    c.project(X[:64], mean, W)                                    # warm-up (allocations, first launch)
    t0 = time.time()
    out = c.project(X, mean, W)
    dt = time.time() - t0
    flop = 2.0 * T * S * G * G
    print("K4 projection: T %d S %d G %d M %d : %.3f s wall incl. H2D of rows+modes and D2H (%.1f MB in, %.1f MB out) -> %.2f TFLOP/s end to end"
          % (T, S, G, G, dt, (X.nbytes + W.nbytes) / 1e6, out.nbytes / 1e6, flop / dt / 1e12), flush=True)
    t0 = time.time()
    ref = (X[:, 3, :].astype(np.float64) - mean[3]) @ W[3].T
    dn = time.time() - t0
    sc = np.abs(X[:, 3, :].astype(np.float64) - mean[3]) @ np.abs(W[3]).T
    print("numpy float64 one site: %.3f s (x%d sites = %.1f s); max |diff|/scale %.3g" % (dn, S, dn * S, (np.abs(out[:, 3, :] - ref) / sc).max()))

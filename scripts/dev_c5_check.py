"""Developer check: the C5 shape (BASELINE.json configs[4]) — per-site covariance of 24 sites x ~1000 groups over 1e5
frames through K3 alone (FP64 DMMA), rows synthesised on the device; parity against float64 numpy on a site."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cgromcorrfmo_b200 import capi, synth
import bench

T = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
d = "/tmp/dev_c5"; os.makedirs(d, exist_ok=True)
s, top, itp, blob, fb = bench.make_c2(d, 1)
with capi.Context(top, itp, 0) as c:
    c.open_memory(blob)
    S, G = c.n_sites, c.num_tot
    print("S", S, "G", G, "frames", T, "batch", B, flush=True)
    g = torch.Generator(device="cuda").manual_seed(5)
    mean = torch.randn((1, S, G), device="cuda", generator=g) * 20.0
    c.set_option("profile", 1)
    c.cov_reset()
    keep = []
    torch.cuda.synchronize()
    t0 = time.time()
    for t in range(0, T, B):
        n = min(B, T - t)
        X = (mean + 0.05 * torch.randn((n, S, G), device="cuda", generator=g)).to(torch.float32).contiguous()
        c.cov_accumulate_device(X.data_ptr(), n)
        keep.append(X[:, 3, :].double().cpu())      # site 3 for the check
    torch.cuda.synchronize()
    wall = time.time() - t0
    kt = c.kernel_times(reset=True)["cov"]
    flop = S * T * G * (G + 1)
    print("K3 total %.1f ms over %d launches -> %.2f TFLOP/s (S*T*G*(G+1)); wall incl. row synthesis %.2f s" % (kt[0], kt[1], flop / kt[0] / 1e9, wall), flush=True)
    mean_g, cov_g = c.cov_finalize(1)
X3 = torch.cat(keep).numpy()
ref = np.cov(X3, rowvar=False)
sc = np.maximum(np.abs(ref), np.sqrt(np.outer(np.diag(ref), np.diag(ref))))
print("site 3: max |cov - np.cov| / scale = %.3g ; mean rel err %.3g" % ((np.abs(cov_g[3] - ref) / sc).max(), np.abs(mean_g[3] - X3.mean(0)).max() / np.abs(X3.mean(0)).max()))

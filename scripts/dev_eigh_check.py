"""Developer check: K6 at the C5 shape (24 matrices of 1099 x 1099) against numpy.linalg.eigh, with timings."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgromcorrfmo_b200 import capi
import bench

S = int(sys.argv[1]) if len(sys.argv) > 1 else 24
d = "/tmp/dev_c5"; os.makedirs(d, exist_ok=True)
s, top, itp, blob, fb = bench.make_c2(d, 1)
rng = np.random.default_rng(9)
G = 1099
A = []
for k in range(S):
    X = rng.normal(size=(3000, G)) * (0.02 + rng.random(G))
    c = np.cov(X, rowvar=False); c[k, :] = 0; c[:, k] = 0
    A.append(c)
A = np.stack(A)
with capi.Context(top, itp, 0) as c:
    c.open_memory(blob)
    c.eigh(A[:1, :64, :64].copy())
    t0 = time.time()
    ev, vt, sweeps = c.eigh(A)
    dt = time.time() - t0
print("K6: %d matrices of %d x %d: %.2f s, %d sweeps" % (S, G, G, dt, sweeps), flush=True)
t0 = time.time()
ref = [np.linalg.eigh(A[k]) for k in range(S)]
dn = time.time() - t0
print("numpy.linalg.eigh: %.2f s on %d cores" % (dn, len(os.sched_getaffinity(0))))
err = max(np.abs(ev[k] - ref[k][0]).max() / np.abs(ref[k][0]).max() for k in range(S))
res = max(np.abs(vt[k].T @ np.diag(ev[k]) @ vt[k] - A[k]).max() / np.abs(A[k]).max() for k in range(min(S, 3)))
print("max eigenvalue error / lambda_max %.3g ; reconstruction error (3 matrices) %.3g" % (err, res))

"""Developer check: one full-size C2 frame (100 000 atoms, 24 sites, 1099 columns) through the GPU path against the
UNMODIFIED reference library (oracle/_ref/ref_harness, all 24 sites, float32 sequential) and against the float64
restatement.  ~70 s of CPU (the reference's O(T^2) topology pass) + a few seconds."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgromcorrfmo_b200 import capi, synth
from oracle import oracle as O

d = "/tmp/dev_c2ref"; os.makedirs(d, exist_ok=True)
s = synth.named_system("C2")
top, itp, traj = d + "/topol.top", d + "/bcx.itp", d + "/one.gro"
synth.write_top(top, s); synth.write_itp(itp, s)
open(traj, "wb").write(synth.frame_bytes(s, synth._static_prefix(s), 0))
with capi.Context(top, itp, 0) as c:
    c.open(traj)
    dE, done, rc = c.run(0, 1)
    groups = c.groups()
print("GPU: dims", dE.shape, flush=True)
t0 = time.time()
ref = O.run_ref_harness(traj, top, itp, 1, d + "/ref")
print("reference harness: %.0f s (topology %.0f s)" % (time.time() - t0, ref["timing"]["topo_s"]), flush=True)
assert np.array_equal(groups, ref["groups"]), "group ids differ"
r32 = ref["dE_f32"][0].astype(np.float64)
# the float64 restatement gives the exact values and the cancellation-aware scale (sum over pairs of |term|)
t0 = time.time()
names, masses, charges = O.load_tables(top, itp)
index = {}
for j, n in enumerate(names):
    index.setdefault(n, j)
_, Xi, keys, grp, _ = O.read_frame(traj, None)
worst_gpu = worst_ref = 0.0
for site in range(1, 25):
    q = O.atom_data_lookup(keys, grp, site, names, masses, charges, index)[2]
    r64, sc = O.compute_cdc_f64(site, Xi, q, grp)
    scale = np.maximum(np.abs(r64), sc); m = scale > 0
    worst_gpu = max(worst_gpu, float((np.abs(dE[0, site - 1].astype(np.float64) - r64)[m] / scale[m]).max()))
    worst_ref = max(worst_ref, float((np.abs(r32[site - 1] - r64)[m] / scale[m]).max()))
    assert np.all(dE[0, site - 1][~m] == 0)
print("float64 restatement: %.0f s" % (time.time() - t0))
print("C2 frame 0, 24 sites x 1099 columns: max scaled error GPU vs exact %.3g ; the unmodified reference (float32) vs exact %.3g"
      % (worst_gpu, worst_ref))
d_gr = np.abs(dE[0].astype(np.float64) - r32)
print("GPU vs the reference's float32 rows directly: max |diff| %.3g kcal/mol, max |row entry| %.3g" % (d_gr.max(), np.abs(r32).max()))

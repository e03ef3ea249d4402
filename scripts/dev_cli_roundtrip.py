"""Developer check: the whole chain on a C2 file — CLI (`.time` text formatted on the device, K5, plus the --npy sidecar),
then the text read back on the device (K8) — and the two compared: every parsed value must be float("%g" % v) of the value
the sidecar holds.   python scripts/dev_cli_roundtrip.py [frames] [dir]"""
import os, subprocess, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgromcorrfmo_b200 import build, capi, pypca, synth

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 512
d = (sys.argv[2] if len(sys.argv) > 2 else "/tmp") + "/dev_rt"; os.makedirs(d, exist_ok=True)
s = synth.named_system("C2")
top, itp = d + "/topol.top", d + "/bcx.itp"
synth.write_top(top, s); synth.write_itp(itp, s)
prefix = synth._static_prefix(s)
distinct = [synth.frame_bytes(s, prefix, f) for f in range(16)]
with open(d + "/traj.gro", "wb") as f:
    for i in range(frames):
        f.write(distinct[i % 16])
t0 = time.time()
r = subprocess.run([build.CLI, d + "/traj.gro", top, itp, d + "/out", str(frames), "--npy", "--timing"], capture_output=True, text=True)
print("CLI rc %d in %.2f s (%.0f frames/s incl. start-up) | %s" % (r.returncode, time.time() - t0, frames / (time.time() - t0), r.stderr.strip()[-400:]), flush=True)
assert r.returncode == 0
kcal = np.load(d + "/out.time.npy", mmap_mode="r")
print("sidecar", kcal.shape, kcal.dtype, "; .time %.1f MB" % (os.path.getsize(d + "/out.time") / 1e6))
with capi.Context(top, itp, 0) as c:
    t0 = time.time()
    E = pypca.load_E_t_ia_csv(c, d + "/out.time", Nsites=kcal.shape[1], dtype=np.float64)
    print("K8: %.1f MB of text -> %s float64 in %.3f s" % (os.path.getsize(d + "/out.time") / 1e6, E.shape, time.time() - t0))
assert E.shape == kcal.shape
cm = (np.asarray(kcal, dtype=np.float64) * 349.7).astype(np.float32)          # src/FMOparse.cpp:151
rel = np.abs(E - cm) / np.maximum(np.abs(cm), 1e-300)
print("all %d values within 6 significant digits of the sidecar: max rel %.2e" % (E.size, rel[cm != 0].max()))
assert rel[cm != 0].max() < 5.1e-6 and np.all(E[cm == 0] == 0)
idx = np.random.default_rng(0).integers(0, E.size, size=200000)
want = np.array([float("%g" % v) for v in cm.reshape(-1)[idx]])
assert np.array_equal(want.view(np.uint64), E.reshape(-1)[idx].view(np.uint64))
print("200000 sampled values bit-identical to float('%g' % value)")

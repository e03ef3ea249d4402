"""Developer check: the CLI driver on a real C2 trajectory file (page cache -> pinned staging -> GPU -> .time/.mtx)."""
import os, subprocess, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgromcorrfmo_b200 import synth, build
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 256
d = "/tmp/dev_cli"; os.makedirs(d, exist_ok=True)
s = synth.named_system("C2")
prefix = synth._static_prefix(s)
synth.write_top(d + "/topol.top", s); synth.write_itp(d + "/bcx.itp", s)
t0 = time.time()
distinct = [synth.frame_bytes(s, prefix, f) for f in range(32)]
with open(d + "/traj.gro", "wb") as f:
    for i in range(frames):
        f.write(distinct[i % 32])
print("wrote %d frames (%.2f GB) in %.1fs" % (frames, os.path.getsize(d + "/traj.gro") / 1e9, time.time() - t0), flush=True)
for extra in (["--no-mtx", "--timing"], ["--timing"], ["--batch-frames", "32", "--timing"]):
    t0 = time.time()
    r = subprocess.run([build.CLI, d + "/traj.gro", d + "/topol.top", d + "/bcx.itp", d + "/out", str(frames)] + extra, capture_output=True, text=True)
    dt = time.time() - t0
    print("CLI", extra, "rc", r.returncode, "%.2fs -> %.0f frames/s" % (dt, frames / dt), "| .time %.1f MB" % (os.path.getsize(d + "/out.time") / 1e6), r.stderr[-600:], flush=True)

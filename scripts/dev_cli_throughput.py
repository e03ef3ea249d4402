"""Developer check: the CLI end to end on a real C2 file (page-cached), .gro and .trr, with its --timing breakdown."""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgromcorrfmo_b200 import build, synth

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 512
d = (sys.argv[2] if len(sys.argv) > 2 else "/tmp") + "/dev_cli"; os.makedirs(d, exist_ok=True)
s = synth.named_system("C2")
top, itp = d + "/topol.top", d + "/bcx.itp"
synth.write_top(top, s); synth.write_itp(itp, s)
prefix = synth._static_prefix(s)
t0 = time.time()
distinct = [synth.frame_bytes(s, prefix, f) for f in range(16)]
with open(d + "/traj.gro", "wb") as f:
    for i in range(frames):
        f.write(distinct[i % 16])
with open(d + "/frame0.gro", "wb") as f:
    f.write(distinct[0])
tdist = [synth.trr_frame_bytes(s, f) for f in range(16)]
with open(d + "/traj.trr", "wb") as f:
    for i in range(frames):
        f.write(tdist[i % 16])
print("inputs written in %.1f s: %.2f GB .gro, %.2f GB .trr" % (time.time() - t0, os.path.getsize(d + "/traj.gro") / 1e9, os.path.getsize(d + "/traj.trr") / 1e9), flush=True)
for name, args in (("gro", [d + "/traj.gro", top, itp, d + "/out_gro", str(frames)]),
                   ("gro --no-mtx", [d + "/traj.gro", top, itp, d + "/out_gro2", str(frames), "--no-mtx"]),
                   ("trr", [d + "/traj.trr", top, itp, d + "/out_trr", str(frames), "--structure", d + "/frame0.gro"])):
    for rep in range(2):
        t0 = time.time()
        r = subprocess.run([build.CLI] + args + ["--timing"], capture_output=True, text=True, env=dict(os.environ, CGC_TRACE="1") if rep == 1 else None)
        dt = time.time() - t0
        print("%-14s rep %d rc %d wall %.2f s = %.0f frames/s | %s" % (name, rep, r.returncode, dt, frames / dt, r.stderr.strip()[-1800:]), flush=True)
print(".time size %.2f GB" % (os.path.getsize(d + "/out_gro.time") / 1e9))

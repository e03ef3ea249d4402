"""Developer check: the CLI end to end on a real C2 file in the page cache (/dev/shm by default), .gro and .trr, with its
--timing breakdown and the host-thread trace of the streaming engine (CGC_TRACE).

    python scripts/dev_cli_throughput.py [frames] [dir] [gpus,gpus,...]
"""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgromcorrfmo_b200 import build, synth

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
d = (sys.argv[2] if len(sys.argv) > 2 else "/dev/shm") + "/dev_cli"; os.makedirs(d, exist_ok=True)
gpu_list = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1]
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
ref_time = None
for g in gpu_list:
    extra = ["--gpus", str(g)] if g > 1 else []
    for name, args in (("gro", [d + "/traj.gro", top, itp, d + "/out_gro", str(frames)]),
                       ("gro --no-mtx", [d + "/traj.gro", top, itp, d + "/out_gro2", str(frames), "--no-mtx"]),
                       ("trr", [d + "/traj.trr", top, itp, d + "/out_trr", str(frames), "--structure", d + "/frame0.gro"])):
        for rep in range(2):
            t0 = time.time()
            try:
                r = subprocess.run([build.CLI] + args + extra + ["--timing"], capture_output=True, text=True, timeout=180,
                                   env=dict(os.environ, CGC_TRACE="1") if rep == 1 else None)
            except subprocess.TimeoutExpired as e:
                print("gpus %d %-14s rep %d TIMED OUT after 180 s | %s" % (g, name, rep, (e.stderr or b"")[-800:]), flush=True)
                continue
            dt = time.time() - t0
            print("gpus %d %-14s rep %d rc %d wall %.2f s = %.0f frames/s | %s" % (g, name, rep, r.returncode, dt, frames / dt, r.stderr.strip()[-2400:]), flush=True)
    import hashlib
    h = hashlib.sha256(open(d + "/out_gro.time", "rb").read()).hexdigest()
    h2 = hashlib.sha256(open(d + "/out_trr.time", "rb").read()).hexdigest()
    if ref_time is None:
        ref_time = h
    print("gpus %d: .time size %.2f GB, sha %s (%s the first run's), .trr run %s" % (
        g, os.path.getsize(d + "/out_gro.time") / 1e9, h[:16], "same as" if h == ref_time else "DIFFERS from",
        "identical text" if h2 == h else "DIFFERENT text"), flush=True)
for f in os.listdir(d):
    os.remove(os.path.join(d, f))

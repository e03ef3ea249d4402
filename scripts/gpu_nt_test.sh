#!/bin/bash
python - <<'PY'
import os, subprocess, sys, time
sys.path.insert(0, '.')
from cgromcorrfmo_b200 import build, synth
frames = 6144
d = "/dev/shm/nt"; os.makedirs(d, exist_ok=True)
s = synth.named_system("C2")
top, itp = d + "/topol.top", d + "/bcx.itp"
synth.write_top(top, s); synth.write_itp(itp, s)
prefix = synth._static_prefix(s)
distinct = [synth.frame_bytes(s, prefix, f) for f in range(16)]
with open(d + "/traj.gro", "wb") as f:
    for i in range(frames): f.write(distinct[i % 16])
for rep in range(3):
    for nt in ("0", "1"):
        for extra in ([], ["--no-mtx"]):
            r = subprocess.run([build.CLI, d + "/traj.gro", top, itp, d + "/out", str(frames), "--timing"] + extra, capture_output=True, text=True,
                               env=dict(os.environ, CGC_TRACE="1", CGC_STAGE_NT=nt), timeout=120)
            tr = [l for l in r.stderr.split("\n") if l.startswith("cgc trace")]
            tm = [l for l in r.stderr.split("\n") if l.startswith("timing")]
            print("NT", nt, " ".join(extra), "|", tr[0][:60] if tr else "", "|", tm[0][tm[0].index("frame loop"):tm[0].index("frame loop")+40] if tm else "", "|", tr[0][tr[0].index("host thread"):][:150] if tr else "", flush=True)
for f in os.listdir(d): os.remove(os.path.join(d, f))
PY

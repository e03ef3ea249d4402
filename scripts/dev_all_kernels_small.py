"""Small end-to-end run that touches every kernel (K1 text + .trr, K2 both loop instantiations, finish, K3, K4,
K5) on tiny inputs."""
import os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgromcorrfmo_b200 import capi, synth

with tempfile.TemporaryDirectory() as d:
    for n_dq in (44, 9):
        s = synth.make_system(3000, 5, 77 + n_dq, "shared")
        if n_dq != 44:
            has_b = np.zeros(140, dtype=bool); has_b[:n_dq] = True
            s.bcl_has_b = has_b
            s.bcl_qB = np.where(has_b, np.round(s.bcl_qA + 0.07, 3), s.bcl_qA)
        sub = os.path.join(d, "c%d" % n_dq); os.makedirs(sub)
        traj, top, itp = synth.write_inputs(sub, s, 3)
        trr = os.path.join(sub, "t.trr"); synth.write_trr(trr, s, 3, double=(n_dq != 44), velocities=True)
        with capi.Context(top, itp, 0) as c:
            c.open(traj)
            c.set_option("batch_frames", 2)
            dE, done, rc = c.run_write_time(0, 3, os.path.join(sub, "o.time"), truncate=True, want_dE=True)
            mean, cov = c.cov_finalize(1)
            w = np.linalg.eigh(cov)[1].transpose(0, 2, 1).copy()
            proj = c.project(dE, mean, w[:, -3:, :])
            print(n_dq, "gro", done, float(np.abs(dE).max()), proj.shape)
        with capi.Context(top, itp, 0) as c:
            c.open_trr(trr, traj)
            dE2, done, rc = c.run(0, 3)
            print(n_dq, "trr", done, float(np.abs(dE2 - dE).max()))

"""Developer diagnostic: the sequence of tests/test_engine.py::test_guard_bands_stay_intact with prints instead of asserts."""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgromcorrfmo_b200 import capi, synth
import torch, tempfile
d = tempfile.mkdtemp()
systems = [(synth.make_system(5000, 7, 43, "per_molecule"), 2, True, 7), (synth.make_system(2000, 8, 42, "shared"), 3, False, 30)]
for k, (s, frames, vel, reps) in enumerate(systems):
    dd = os.path.join(d, "s%d" % k)
    traj, top, itp = synth.write_inputs(dd, s, frames, velocities=vel)
    blob = open(traj, "rb").read() * reps
    path = os.path.join(dd, "g.gro"); open(path, "wb").write(blob)
    total = frames * reps
    for batch in (1, 3, 0):
        for variant in ("plain", "compat+mtx"):
            with capi.Context(top, itp, 0) as c:
                c.open(path)
                if batch: c.set_option("batch_frames", batch)
                if variant != "plain": c.set_option("mtx_compat", 1)
                a, done, rc = c.run_write_time(0, total, path + ".time", want_dE=True)
                if variant != "plain":
                    c.cov_finalize(1)
                    c.write_mtx(path + ".mtx")
                text = np.frombuffer(blob, dtype=np.uint8)
                d_text = torch.from_numpy(np.concatenate([text, np.zeros(((-text.size) % 16) + 16, np.uint8)])).cuda()
                offs = [c.frame_atoms_offset(f) for f in range(total)]
                d_dE = torch.empty((total, c.n_sites, c.num_tot), dtype=torch.float32, device="cuda")
                c.run_device(d_text.data_ptr(), text.size, offs, total, d_dE.data_ptr())
                torch.cuda.synchronize()
                e = d_dE.cpu().numpy()
                c.cov_accumulate_device(d_dE.data_ptr(), total)
                xyz = torch.empty((total, c.dims()[0], 3), dtype=torch.float32, device="cuda")
                c.decode_device(d_text.data_ptr(), text.size, offs, total, xyz.data_ptr())
                torch.cuda.synchronize()
                e2 = d_dE.cpu().numpy()
                bad_frames = [int(f) for f in np.nonzero((a != e).reshape(total, -1).any(1))[0]]
                print("guard", os.environ.get("CGC_GUARD"), "system", k, "batch", batch, variant, "| rows equal:", np.array_equal(a.view(np.uint32), e.view(np.uint32)),
                      "after K3/decode:", np.array_equal(e.view(np.uint32), e2.view(np.uint32)), "max|diff| %.3g" % float(np.abs(a - e).max()),
                      "frames differing", bad_frames[:8], "entries", int((a != e).sum()), "decode errs", c.decode_errors(), flush=True)
                if os.environ.get("CGC_GUARD"): print("   guards", capi.debug_guard_check(), flush=True)

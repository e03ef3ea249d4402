import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgromcorrfmo_b200 import capi, synth
import torch, tempfile
d = tempfile.mkdtemp()
s = synth.make_system(2000, 8, 42, "shared")
traj, top, itp = synth.write_inputs(d, s, 3)
reps = 30
blob = open(traj, "rb").read() * reps
path = os.path.join(d, "g.gro"); open(path, "wb").write(blob)
total = 3 * reps
with capi.Context(top, itp, 0) as c:
    c.open(path)
    c.set_option("batch_frames", 1)
    a0, done, rc = c.run(0, total)
    a, done, rc = c.run_write_time(0, total, path + '.time', want_dE=True)
    print('run vs run_write_time bits equal:', np.array_equal(a0.view(np.uint32), a.view(np.uint32)))
    cols = c.columns(); groups = c.groups()
    text = np.frombuffer(blob, dtype=np.uint8)
    d_text = torch.from_numpy(np.concatenate([text, np.zeros(((-text.size) % 16) + 16, np.uint8)])).cuda()
    offs = [c.frame_atoms_offset(f) for f in range(total)]
    d_dE = torch.empty((total, c.n_sites, c.num_tot), dtype=torch.float32, device="cuda")
    c.run_device(d_text.data_ptr(), text.size, offs, total, d_dE.data_ptr())
    torch.cuda.synchronize()
    e = d_dE.cpu().numpy()
    diff = np.argwhere(a.view(np.uint32) != e.view(np.uint32))
    print("guard", os.environ.get("CGC_GUARD"), "differing entries", len(diff))
    for f, si, g in diff[:12]:
        print("  frame %d site %d col %d: run %r (bits %08x)  device %r (bits %08x); atoms in that column: %d" % (
            f, si, g, float(a[f, si, g]), a.view(np.uint32)[f, si, g], float(e[f, si, g]), e.view(np.uint32)[f, si, g], int((cols == g).sum())))
    print("  columns involved", sorted(set(int(x[2]) for x in diff))[:20], "sites", sorted(set(int(x[1]) for x in diff)), "frames", sorted(set(int(x[0]) for x in diff))[:10])

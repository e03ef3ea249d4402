"""Short device-path run for ncu: C2 shape, F frames, a few launches of K1/K2/finish/K3."""
import os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cgromcorrfmo_b200 import capi
import bench

F = int(sys.argv[1]) if len(sys.argv) > 1 else 24
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
with tempfile.TemporaryDirectory() as d:
    s, top, itp, blob, fb = bench.make_c2(d, F)
    text = np.frombuffer(blob, dtype=np.uint8)
    dev = torch.from_numpy(np.concatenate([text, np.zeros(64, np.uint8)])).cuda()
    with capi.Context(top, itp, 0) as c:
        c.open_memory(blob)
        c.set_option("batch_frames", F)
        offs = [c.frame_atoms_offset(f) for f in range(F)]
        dE = torch.empty((F, c.n_sites, c.num_tot), dtype=torch.float32, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(reps):
            e0.record()
            c.run_device(dev.data_ptr(), text.size, offs, F, dE.data_ptr(), st)
            c.cov_accumulate_device(dE.data_ptr(), F, st)
            e1.record()
            torch.cuda.synchronize()
            print("rep", i, "ms", e0.elapsed_time(e1), flush=True)

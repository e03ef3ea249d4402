"""Developer check on a GPU box: parity of every kernel against the oracle on small inputs + first timings.
Not part of the product; writes gpurun_out/dev_check.log."""
import json, os, sys, time, traceback
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cgromcorrfmo_b200 import capi, synth
from oracle import oracle as O

os.makedirs("gpurun_out", exist_ok=True)
LOG = open("gpurun_out/dev_check.log", "w")
def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True); LOG.write(s + "\n"); LOG.flush()

def scaled_err(dE, ref64, scale):
    d = np.abs(dE.astype(np.float64) - ref64)
    s = np.maximum(np.abs(ref64), scale)
    m = s > 0
    return float((d[m] / s[m]).max()), float(np.abs(dE[~m]).max() if (~m).any() else 0.0)

def check_small(tag, n_atoms, n_bcl, seed, solvent, frames, velocities=False):
    d = f"/tmp/dev_{tag}"
    s = synth.make_system(n_atoms, n_bcl, seed, solvent)
    traj, top, itp = synth.write_inputs(d, s, frames, velocities=velocities)
    ref = O.run(traj, top, itp, frames)
    with capi.Context(top, itp, 0) as c:
        c.open(traj)
        dE, done, rc = c.run(0, frames)
        e, z = scaled_err(dE, ref["dE_f64"], ref["scale"])
        e32, _ = scaled_err(ref["dE_f32"], ref["dE_f64"], ref["scale"])
        log(f"[{tag}] frames {done} rc {rc} dims {c.dims()} max scaled err GPU {e:.3g} (reference f32 itself {e32:.3g}) zero-cols max {z:.3g} decode_err {c.decode_errors()}")
        tot = np.abs(dE.astype(np.float64).sum(-1) - ref["dE_f64"].sum(-1)) / np.abs(ref["dE_f64"].sum(-1))
        log(f"[{tag}] per-site total rel err max {tot.max():.3g}")
        # covariance
        if c.cov_size():
            mean, cov = c.cov_finalize(0)
            m64, c64 = O.cov_f64(dE, 0)
            sc = np.maximum(np.abs(c64), np.sqrt(np.einsum('sii,sjj->sij', c64, c64)))
            ok = sc > 0
            log(f"[{tag}] cov rel err max {(np.abs(cov - c64)[ok] / sc[ok]).max():.3g}  mean rel err {np.abs(mean-m64).max() / np.abs(m64).max():.3g}")
            c.write_mtx(d + "/out.mtx", compat=True)
            mt = O.format_mtx(O.mtx_compat_f32(dE))
            log(f"[{tag}] compat .mtx byte-equal to oracle on same dE: {mt == open(d + '/out.mtx').read()}")
        capi.write_time(d + "/out.time", dE)
        log(f"[{tag}] .time byte-equal to oracle formatting: {O.format_time_rows(dE.reshape(-1, dE.shape[-1])) == open(d + '/out.time').read()}")
        # decode parity (bit-exact) through the device entry point
        text = np.fromfile(traj, dtype=np.uint8)
        pad = (-text.size) % 16
        t = torch.from_numpy(np.concatenate([text, np.zeros(pad + 16, np.uint8)])).cuda()
        offs = [c.frame_atoms_offset(f) for f in range(frames)]
        xyz = torch.empty((frames, c.n_atoms, 3), dtype=torch.float32, device="cuda")
        c.decode_device(t.data_ptr(), text.size, offs, frames, xyz.data_ptr())
        torch.cuda.synchronize()
        x0 = xyz[0].cpu().numpy().reshape(-1)
        log(f"[{tag}] decode frame0 bit-exact vs (float)atof: {np.array_equal(x0.view(np.uint32), ref['x0'].view(np.uint32))}")
        # device path equals host path
        dEd = torch.empty((frames, c.n_sites, c.num_tot), dtype=torch.float32, device="cuda")
        c.run_device(t.data_ptr(), text.size, offs, frames, dEd.data_ptr())
        torch.cuda.synchronize()
        log(f"[{tag}] device path max |diff| vs host path: {np.abs(dEd.cpu().numpy() - dE).max():.3g}")

def timing_c2(frames=16, reps=5):
    log("building C2 system ...")
    t0 = time.time()
    s = synth.named_system("C2")
    prefix = synth._static_prefix(s)
    d = "/tmp/dev_c2"; os.makedirs(d, exist_ok=True)
    top, itp = d + "/topol.top", d + "/bcx.itp"
    synth.write_top(top, s); synth.write_itp(itp, s)
    blob = b"".join(synth.frame_bytes(s, prefix, f) for f in range(frames))
    log(f"C2 generated in {time.time()-t0:.1f}s, {len(blob)/frames/1e6:.2f} MB/frame")
    text = np.frombuffer(blob, dtype=np.uint8)
    host = torch.from_numpy(np.concatenate([text, np.zeros(32, np.uint8)])).pin_memory()
    with capi.Context(top, itp, 0) as c:
        t0 = time.time()
        c.open_memory((host.data_ptr(), text.size))
        log(f"open: {time.time()-t0:.2f}s dims {c.dims()} site_cap {c.get_option('site_cap')}")
        c.set_option("batch_frames", frames)
        S, G = c.n_sites, c.num_tot
        dev = host.cuda()
        offs = [c.frame_atoms_offset(f) for f in range(frames)]
        dEd = torch.empty((frames, S, G), dtype=torch.float32, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(2):
            c.run_device(dev.data_ptr(), text.size, offs, frames, dEd.data_ptr(), st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            c.run_device(dev.data_ptr(), text.size, offs, frames, dEd.data_ptr(), st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        ndq = int(c.get_option("site_cap"))
        pairs = frames * S * ndq * (c.n_atoms - 140)
        log(f"C2 device path: {ms:.3f} ms per {frames} frames -> {frames/ms*1e3:.1f} frames/s, {pairs/ms/1e6:.3f} Gpairs/s (44-conv) roofline 16/clk/SM@1.965GHz=4654 -> {pairs/ms/1e6/4654:.3f}")
        # covariance accumulate timing
        c.cov_accumulate_device(dEd.data_ptr(), frames, st); torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            c.cov_accumulate_device(dEd.data_ptr(), frames, st)
        e1.record(); torch.cuda.synchronize()
        msc = e0.elapsed_time(e1) / reps
        fl = frames * S * G * G * 2
        log(f"C2 cov accumulate: {msc:.3f} ms per {frames} frames, {fl/msc/1e9:.2f} TFLOP/s (full-square count)")
        # end to end from pinned host text
        t0 = time.time()
        dE, done, rc = c.run(0, frames)
        t1 = time.time() - t0
        dE, done, rc = c.run(0, frames)
        t0 = time.time()
        dE, done, rc = c.run(0, frames)
        t2 = time.time() - t0
        log(f"C2 e2e host->host: first {t1:.3f}s, steady {t2*1e3:.2f} ms per {frames} frames -> {frames/t2:.1f} frames/s")
        # parity of frame 0 vs oracle f64 (C, a few sites)
        fr = O.read_frame  # oracle on 1 frame
        open(d + "/one.gro", "wb").write(blob[: len(blob) // frames])
        t0 = time.time()
        ref = O.run(d + "/one.gro", top, itp, 1)
        log(f"oracle 1 frame C2 all sites: {time.time()-t0:.1f}s")
        e, z = scaled_err(dE[:1], ref["dE_f64"], ref["scale"])
        e32, _ = scaled_err(ref["dE_f32"], ref["dE_f64"], ref["scale"])
        log(f"C2 frame0 max scaled err GPU {e:.3g} (reference f32 {e32:.3g})")

for fn, args in [(check_small, ("tiny", 2000, 8, 42, "shared", 3)),
                 (check_small, ("vel", 5000, 7, 43, "per_molecule", 2, True)),
                 (check_small, ("noSolv", 3000, 9, 44, "none", 2)),
                 (timing_c2, ())]:
    try:
        fn(*args)
    except Exception:
        log("FAILED", fn.__name__, args, traceback.format_exc())

#!/usr/bin/env python
"""Benchmark of the traj2dEcsv hot path on B200 (contract: see the repo prompt / DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--frames-per-step F]

Workload (BASELINE.json configs[1], "C2"): synthetic FMO trimer, 100 000 atoms, 24 BCL chromophores, ~1.1k residue
groups, fixed-column .gro text (45 B per atom line), seed 1002.  One STEP = one batch of F frames (default 256) pushed
through the whole hot path: K1 text decode -> K2 CDC pair sums for all 24 sites -> float32 dE rows -> K3 Correlate sums.
The F frames are distinct pre-generated frames; every step re-streams them (a 10 000-frame file would be 45 GB), which
keeps the per-step input (F x 4.5 MB = 1.15 GB at F = 256) far above the 126 MB L2.

  value   frames/s, whole job, text resident in HBM when the timed region starts (cgc_run_device + K3)
  e2e     frames/s through the C-ABI call a user makes (cgc_run) with HOST buffers: pinned text in, H2D copies, kernels,
          D2H of the dE rows — all inside the timed region
  N > 1   one process per GPU (torchrun), every rank streams its own frame range (weak scaling, no data-path
          collective); the Correlate partial sums are combined by ONE NCCL all-reduce, inside the timed region
  --impl reference   the reference's own CPU implementation (oracle/_ref/ref_harness = the unmodified reference library,
          all sites) on the host cores, one process per core over split trajectories as the authors ran it
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "CDC frames/s (FMO trimer ~100k atoms)"
UNIT = "frames/s"
WORKLOAD = "C2"
SM_COUNT = 148
PAIRS_PER_CLK_PER_SM = 16.0   # 8 issue slots per pair (7 FP32 + 1 MUFU.RSQ): MUFU at 16 lanes/clk/SM sets the same ceiling


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=float(d.get("hbm_gbs", 6650.0)), sm_max_mhz=float(d.get("sm_max_mhz", 1965.0)), source="measured")
    return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, source="fallback")


def k2_traffic_per_frame():
    """dram__bytes_read.sum + dram__bytes_write.sum of cdc_kernel per frame of the C2 workload, from the committed
    `ncu --set full` capture (profiles/r01_k2_traffic.json); None when that file is absent."""
    p = os.path.join(ROOT, "profiles", "r01_k2_traffic.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["dram_bytes_per_frame"])
        except Exception:
            return None
    return None


def calibrate_fp64_tensor(torch, dev):
    """Yard-stick for K3 only (MEASURED_PEAKS.json has no FP64 figure): cuBLAS DGEMM 4096^3, best of 5, TFLOP/s."""
    n = 4096
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def time_corr_leg(capi, top, itp, device, fp64_peak):
    """K7 (cgc_time_corr), the last step of the pyPCA chain, outside the timed region: 64 pairs of 200 000-sample series
    over 20 000 lags (the scripts' defaults, scripts/depca/depca/sidechain_corr.py:190), device time of its kernels."""
    n, L, n_pairs = 200000, 20000, 64
    series = np.random.default_rng(7).normal(size=(16, n))
    a = np.arange(n_pairs) % 16
    b = (np.arange(n_pairs) // 16 + a + 1) % 16
    with capi.Context(top, itp, device) as c:
        c.time_corr(series[:2], [0], [1], 1024)
        ms = min(c.time_corr(series, a, b, L, want_ms=True)[1] for _ in range(3))
        launches = c.launch_count
    flop = 2.0 * (n * L - L * (L - 1) / 2.0) * n_pairs
    return {"kernel": "time_corr_kernel (K7, pyPCA time correlation; post-processing, not in the step)", "bound": "fp64 tensor (DMMA)",
            "achieved": flop / ms / 1e9, "peak": fp64_peak, "unit": "TFLOP/s", "frac": (flop / ms / 1e9 / fp64_peak) if fp64_peak else None,
            "peak_source": "cuBLAS DGEMM 4096^3 measured in this run", "algorithmic_flop_per_call": flop, "call_ms": ms,
            "workload": "%d pairs, n %d, corr_len %d, linear sums" % (n_pairs, n, L)}, launches


def calibrate_h2d(torch, host, d_text):
    """Yard-stick for the end-to-end leg: pinned host -> device copy of the same text buffer, best of 5, GB/s."""
    best = 1e9
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        d_text.copy_(host, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return host.numel() / (best * 1e-3) / 1e9


def cores_available() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ----------------------------------------------------------------------------------------------------------------------
# synthetic C2 inputs
# ----------------------------------------------------------------------------------------------------------------------
def make_c2(workdir: str, n_frames: int, first_frame: int = 0):
    """Returns (system, top, itp, text bytes of n_frames frames, frame_bytes)."""
    from cgromcorrfmo_b200 import synth
    s = synth.named_system(WORKLOAD)
    top, itp = os.path.join(workdir, "topol.top"), os.path.join(workdir, "bcx.itp")
    synth.write_top(top, s)
    synth.write_itp(itp, s)
    prefix = synth._static_prefix(s)
    frames = [synth.frame_bytes(s, prefix, first_frame + f) for f in range(n_frames)]
    return s, top, itp, b"".join(frames), len(frames[0])


# ----------------------------------------------------------------------------------------------------------------------
# clocks sampled DURING the timed region
# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t0: float, t1: float):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [ln for (t, ln) in self.lines if t0 - 0.05 <= t <= t1 + 0.15] or [ln for (_, ln) in self.lines]
        sm, mx, reasons = [], [], set()
        for ln in rows:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
# the reference's CPU implementation (oracle/_ref), timed on the host cores
# ----------------------------------------------------------------------------------------------------------------------
def time_reference_cpu(workdir: str, n_frames: int, warm_frames: int, sites_sample: int, n_procs: int):
    """Runs `n_procs` copies of the all-sites harness around the UNMODIFIED reference library, one per core, each on its own
    trajectory file (the authors' parallelisation, reference scripts/2014-06-gro2dEcsv.sh:3).  Only `sites_sample` of the
    24 sites are evaluated per frame (sites are independent and cost the same); the per-frame time is
    read + 24/sites_sample * (lookup + cdc).  Start-up (two ParseTopology passes, O(T^2), ~90 s at T = 1e5) is not part
    of the frame loop and is reported separately."""
    from oracle import oracle as O
    kind = "reference"
    if not O.have_ref():
        return None
    from cgromcorrfmo_b200 import synth
    s = synth.named_system(WORKLOAD)
    top, itp = os.path.join(workdir, "ref_topol.top"), os.path.join(workdir, "ref_bcx.itp")
    synth.write_top(top, s)
    synth.write_itp(itp, s)
    prefix = synth._static_prefix(s)
    distinct = [synth.frame_bytes(s, prefix, f) for f in range(min(4, n_frames))]
    traj = os.path.join(workdir, "ref_traj.gro")
    with open(traj, "wb") as f:
        for i in range(n_frames):
            f.write(distinct[i % len(distinct)])
    harness = os.path.join(O.REF_DIR, "ref_harness")
    t0 = time.time()
    procs = []
    for p in range(n_procs):
        out = os.path.join(workdir, "ref_out_%d" % p)
        procs.append((out, subprocess.Popen([harness, traj, top, itp, out, str(n_frames), str(sites_sample), "0"],
                                            stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)))
    per_proc = []
    for out, pr in procs:
        pr.wait()
        if pr.returncode != 0:
            continue
        rows = [[float(v) for v in ln.split()] for ln in open(out + ".frames.txt").read().strip().split("\n")]
        tm = [float(v) for v in open(out + ".timing.txt").read().split()]
        n_sites = 24
        rows = rows[warm_frames:] if len(rows) > warm_frames else rows
        frame_s = np.mean([r[0] + (r[1] + r[2]) * n_sites / max(1, int(tm[6])) for r in rows])
        per_proc.append((frame_s, tm[5], np.mean([r[0] for r in rows]), np.mean([(r[1] + r[2]) / max(1, int(tm[6])) for r in rows])))
    wall = time.time() - t0
    if not per_proc:
        return None
    fps = float(sum(1.0 / p[0] for p in per_proc))
    return dict(value=fps, unit=UNIT, cores=len(per_proc), kind=kind,
                sample="%d processes x %d frames of C2 (after %d warm-up), %d of 24 sites per frame scaled to 24 "
                       "(read %.2f s + 24 x %.3f s per frame per core); start-up %.0f s per process excluded; wall %.0f s"
                       % (len(per_proc), n_frames - warm_frames, warm_frames, sites_sample, per_proc[0][2], per_proc[0][3],
                          per_proc[0][1], wall),
                frames_per_s_per_core=float(np.mean([1.0 / p[0] for p in per_proc])))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n_procs = cores_available()
    n_frames = max(2, min(args.warmup + args.steps, 48))
    warm = min(args.warmup, n_frames - 1)
    with tempfile.TemporaryDirectory(prefix="cgc_ref_") as d:
        t0 = time.time()
        r = time_reference_cpu(d, n_frames, warm, 3, n_procs)
        wall = time.time() - t0
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref (the compiled reference) is missing on this box"}))
        return 0
    timed = n_frames - warm
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["cores"] / r["value"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2: FMO trimer 100000 atoms, 24 BCL sites, per-residue groups, .gro text 45 B/atom",
                   "step": "one frame per host process (%d processes); %d timed frames per process" % (r["cores"], timed)},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall,
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from cgromcorrfmo_b200 import capi, dist as cdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path is CUDA-only, there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cores = cdist.bind_to_gpu_numa(local) if world > 1 else None   # before any pinned allocation
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    F = args.frames_per_step
    peaks = measured_peaks()

    with tempfile.TemporaryDirectory(prefix="cgc_bench_") as d:
        t0 = time.time()
        # every rank streams its own frame range of the synthetic trajectory: frames [rank*F, (rank+1)*F)
        s, top, itp, blob, frame_bytes = make_c2(d, F, first_frame=rank * F)
        text = np.frombuffer(blob, dtype=np.uint8)
        n_text = text.size
        host = torch.empty(n_text + 64, dtype=torch.uint8).pin_memory()
        host[:n_text] = torch.from_numpy(text.copy())
        host[n_text:] = 0
        del blob
        if rank == 0:
            log("bench: generated %d frames (%.1f MB text) in %.1f s" % (F, n_text / 1e6, time.time() - t0))

        ctx = capi.Context(top, itp, local)
        ctx.open_memory((host.data_ptr(), n_text))
        n_atoms, n_chromo, n_groups, G = ctx.dims()
        S = ctx.n_sites
        ctx.set_option("batch_frames", F)
        ctx.set_option("profile", 1)
        ndq = int(ctx.get_option("site_cap"))
        q1 = ctx.site_charges(1)
        n_c = int((ctx.groups() == -1).sum())
        pairs_per_frame = S * ndq * (n_atoms - n_c)                 # dQ != 0 convention
        pairs_per_frame_140 = S * n_c * (n_atoms - n_c)             # the reference's own count (all 140 site atoms)

        # Correlate partial sums live in a torch tensor so torch.distributed can all-reduce it in place
        cov = torch.zeros(ctx.cov_size(), dtype=torch.float64, device=dev)
        ctx.cov_bind(cov.data_ptr())
        d_text = host.to(dev)
        offs = np.array([ctx.frame_atoms_offset(f) for f in range(F)], dtype=np.int64)
        d_dE = torch.empty((F, S, G), dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream().cuda_stream

        # the shift of the Correlate sums must be the same row on every rank: the dE row of the trajectory's frame 0
        ctx.run_device(d_text.data_ptr(), n_text, offs[:1], 1, d_dE.data_ptr(), stream)
        row0 = d_dE[0].clone()
        if world > 1:
            dist.broadcast(row0, 0)
        ctx.cov_set_shift_device(row0.data_ptr(), stream)

        def step_device():
            ctx.run_device(d_text.data_ptr(), n_text, offs, F, d_dE.data_ptr(), stream)
            ctx.cov_accumulate_device(d_dE.data_ptr(), F, stream)

        def combine():
            if world > 1:
                cdist.allreduce_partials(cov, S, G)                    # ONE collective: n, sum(x-c), sum (x-c)(x-c)^T

        def barrier():
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        # ---------------- device-resident: warm-up, then K timed steps
        for _ in range(max(3, args.warmup)):
            step_device()
        barrier()
        ctx.kernel_times(reset=True)
        launches0 = ctx.launch_count
        sampler = ClockSampler(local)
        sampler.start()
        time.sleep(0.25)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        w0 = time.time()
        e0.record()
        for _ in range(args.steps):
            step_device()
        combine()
        e1.record()
        barrier()
        w1 = time.time()
        ms_total = e0.elapsed_time(e1)
        clocks = sampler.stop(w0, w1)
        ktimes = ctx.kernel_times(reset=True)
        launches = ctx.launch_count - launches0
        if world > 1:
            t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_total = float(t.item())
        value = world * args.steps * F / (ms_total / 1e3)

        if args.device_leg_only:
            ctx.close()
            if rank == 0:
                print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                                  "ms_per_step": ms_total / args.steps, "device_leg_only": True,
                                  "kernel_ms": {k: v[0] for k, v in ktimes.items()}, "gpu_launches": int(launches)}))
            if world > 1:
                dist.destroy_process_group()
            return 0

        # ---------------- end to end through cgc_run: pinned host text in, dE rows out to host memory
        out_host = np.empty((F, S, G), dtype=np.float32)
        ctx.set_option("batch_frames", args.e2e_batch)               # several batches per call: H2D overlaps the kernels
        for _ in range(max(3, args.warmup)):
            ctx.run(0, F, out=out_host)
        barrier()
        ctx.kernel_times(reset=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            ctx.run(0, F, out=out_host)
        combine()
        e1.record()
        barrier()
        e2e_s = max(time.perf_counter() - t0, e0.elapsed_time(e1) / 1e3)
        if world > 1:
            t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        e2e_value = world * args.steps * F / e2e_s
        e2e_ktimes = ctx.kernel_times(reset=True)
        same = float(np.abs(out_host - d_dE.cpu().numpy()).max())
        h2d_peak = calibrate_h2d(torch, host, d_text)
        ctx.close()

        # ---------------- the same frames as a GROMACS .trr (single precision, coordinates + box): 12 B/atom over PCIe
        # instead of 45.  Not the headline (the reference reads .gro text); SURVEY.md section 8f-3.
        from cgromcorrfmo_b200 import synth
        struct_gro = os.path.join(d, "frame0.gro")
        with open(struct_gro, "wb") as fh:
            fh.write(bytes(text[:frame_bytes]))
        trr_blob = b"".join(synth.trr_frame_bytes(s, rank * F + f) for f in range(F))
        n_trr = len(trr_blob)
        host_trr = torch.empty(n_trr + 64, dtype=torch.uint8).pin_memory()
        host_trr[:n_trr] = torch.from_numpy(np.frombuffer(trr_blob, dtype=np.uint8).copy())
        host_trr[n_trr:] = 0
        del trr_blob
        ctx2 = capi.Context(top, itp, local)
        ctx2.open_memory_trr((host_trr.data_ptr(), n_trr), struct_gro)
        ctx2.set_option("batch_frames", args.e2e_batch)
        ctx2.set_option("profile", 1)
        cov2 = torch.zeros(ctx2.cov_size(), dtype=torch.float64, device=dev)
        ctx2.cov_bind(cov2.data_ptr())
        ctx2.cov_set_shift_device(row0.data_ptr(), stream)            # the same shift on every rank, as above
        out_trr = np.empty((F, S, G), dtype=np.float32)
        for _ in range(max(3, args.warmup)):
            ctx2.run(0, F, out=out_trr)
        barrier()
        ctx2.kernel_times(reset=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            ctx2.run(0, F, out=out_trr)
        if world > 1:
            cdist.allreduce_partials(cov2, S, G)
        e1.record()
        barrier()
        trr_s = max(time.perf_counter() - t0, e0.elapsed_time(e1) / 1e3)
        if world > 1:
            t = torch.tensor([trr_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            trr_s = float(t.item())
        trr_ktimes = ctx2.kernel_times(reset=True)
        trr_value = world * args.steps * F / trr_s
        trr_same = float(np.abs(out_trr.astype(np.float64) - out_host.astype(np.float64)).max())
        trr_scale = float(np.abs(out_host).max())
        ctx2.close()
        fp64_peak = calibrate_fp64_tensor(torch, dev) if rank == 0 else None
        k7 = None
        if rank == 0:
            try:
                k7, _ = time_corr_leg(capi, top, itp, local, fp64_peak)
            except Exception as e:  # reported, never required
                log("bench: time-correlation leg failed:", e)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---------------- rooflines (dominant kernel: K2)
    k2_ms, k2_n = ktimes["cdc"]
    k1_ms, k1_n = ktimes["decode"]
    k3_ms, k3_n = ktimes["cov"]
    fin_ms, _ = ktimes["finish"]
    k2_avg_s = (k2_ms / max(1, k2_n)) / 1e3
    pairs_per_launch = F * pairs_per_frame
    achieved = pairs_per_launch / k2_avg_s / 1e9 if k2_avg_s > 0 else 0.0
    peak = PAIRS_PER_CLK_PER_SM * SM_COUNT * peaks["sm_max_mhz"] * 1e6 / 1e9
    k1_bytes = F * n_atoms * (45 + 12)
    k1_avg_s = (k1_ms / max(1, k1_n)) / 1e3
    k3_flop = F * S * G * (G + 1)
    k3_avg_s = (k3_ms / max(1, k3_n)) / 1e3
    kern_total = k1_ms + k2_ms + fin_ms + k3_ms
    roofline = {
        "kernel": "cdc_kernel (K2, CDC pair sums)", "bound": "fp32+mufu issue (16 pairs/clk/SM)",
        "achieved": achieved, "peak": peak, "unit": "Gpairs/s", "frac": achieved / peak if peak else None,
        "traffic": (k2_traffic_per_frame() * F) if k2_traffic_per_frame() else None,
        "traffic_source": "profiles/r01_k2_traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per frame) x frames per launch",
        "algorithmic_bytes_per_launch": F * (n_atoms * 12 + S * ndq * 16),
        "peak_source": "16 pairs/clk/SM x %d SMs x %.0f MHz (sm_max_mhz, %s MEASURED_PEAKS.json)" % (SM_COUNT, peaks["sm_max_mhz"], peaks["source"]),
        "algorithmic_pairs_per_launch": pairs_per_launch, "avg_launch_ms": k2_avg_s * 1e3,
        "share_of_kernel_time": k2_ms / kern_total if kern_total else None,
        "frac_at_sampled_clock": (achieved / (PAIRS_PER_CLK_PER_SM * SM_COUNT * clocks["sm_mhz"] * 1e-3)) if clocks.get("sm_mhz") else None,
        "pairs_per_launch_reference_count_140": F * pairs_per_frame_140,
    }
    others = [
        {"kernel": "decode_kernel (K1)", "bound": "hbm", "achieved": k1_bytes / k1_avg_s / 1e9 if k1_avg_s else 0.0,
         "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": (k1_bytes / k1_avg_s / 1e9 / peaks["hbm_gbs"]) if k1_avg_s else None,
         "algorithmic_bytes_per_launch": k1_bytes, "avg_launch_ms": k1_avg_s * 1e3, "share_of_kernel_time": k1_ms / kern_total if kern_total else None},
        {"kernel": "cov_syrk_kernel+cov_sum_kernel (K3)", "bound": "fp64 tensor (DMMA)", "achieved": k3_flop / k3_avg_s / 1e12 if k3_avg_s else 0.0,
         "peak": fp64_peak, "unit": "TFLOP/s", "frac": (k3_flop / k3_avg_s / 1e12 / fp64_peak) if (k3_avg_s and fp64_peak) else None,
         "peak_source": "cuBLAS DGEMM 4096^3 measured in this run (no FP64 figure in MEASURED_PEAKS.json)",
         "algorithmic_flop_per_launch": k3_flop, "avg_launch_ms": k3_avg_s * 1e3,
         "share_of_kernel_time": k3_ms / kern_total if kern_total else None},
    ]

    if k7:
        others.append(k7)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        with tempfile.TemporaryDirectory(prefix="cgc_cpu_") as d:
            try:
                cpu = time_reference_cpu(d, 3, 1, 3, cores_available())
            except Exception as e:  # the baseline is reported, never required
                log("bench: cpu baseline failed:", e)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2: FMO trimer %d atoms, %d BCL sites x %d dQ atoms, numTot %d, .gro text 45 B/atom, seed 1002"
                               % (n_atoms, S, ndq, G),
                   "frames_per_step": F, "frames_per_step_all_ranks": F * world,
                   "step": "K1 decode + K2 CDC (all sites) + float32 rows + K3 Correlate over one batch of frames",
                   "input_bytes_per_step": int(n_text), "l2": "per-step input exceeds the 126 MB L2 (no flush needed)",
                   "parallelism": "frame-sharded x%d, one all-reduce of the Correlate sums" % world,
                   "numa": ("rank 0 bound to %d cores local to its GPU" % len(numa_cores)) if numa_cores else "unbound",
                   "pairs_per_frame": pairs_per_frame, "pair_convention": "dQ != 0 site atoms only (44 of 140)"},
        "pairs_per_s": value * pairs_per_frame, "pairs_per_s_reference_count_140": value * pairs_per_frame_140,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n_text + 8 * F), "d2h_bytes_per_step": int(F * S * G * 4),
                "api": "cgc_run (C-ABI) on pinned host text, %d-frame batches in a 3-slot ring; dE rows copied back to host" % args.e2e_batch,
                "max_abs_diff_vs_device_path": same, "kernel_ms": {k: v[0] for k, v in e2e_ktimes.items()}, "wall_ms": e2e_s * 1e3,
                "bound": "PCIe host->device copy of the text", "h2d_gbs": e2e_value / world * frame_bytes / 1e9,
                "h2d_peak_gbs": h2d_peak, "frac_of_h2d_peak": (e2e_value / world * frame_bytes / 1e9) / h2d_peak,
                "h2d_peak_source": "pinned cudaMemcpyAsync of the same text buffer, best of 6, measured in this run"},
        "e2e_trr": {"value": trr_value, "unit": UNIT, "h2d_bytes_per_step": int(n_trr + 8 * F), "d2h_bytes_per_step": int(F * S * G * 4),
                    "api": "cgc_open_memory_trr + cgc_run on pinned .trr frames (float32 coordinates + box, %d B/frame), %d-frame batches"
                           % (n_trr // F, args.e2e_batch),
                    "note": "same frames as the headline; not the reference's input format (it reads .gro text)",
                    "max_abs_diff_vs_gro_path": trr_same, "max_abs_dE": trr_scale,
                    "kernel_ms": {k: v[0] for k, v in trr_ktimes.items()}, "wall_ms": trr_s * 1e3},
        "gpu_launches": int(launches),
        "kernel_ms": {k: v[0] for k, v in ktimes.items()},
        "roofline": roofline, "rooflines_other": others, "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=256)
    ap.add_argument("--e2e-batch", type=int, default=32, help="frames per device batch inside cgc_run (e2e leg)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--device-leg-only", action="store_true",
                    help="profiling aid: run only the device-resident leg (the ncu launch list then holds nothing but "
                         "the whole-step launches whose shares bench.py reports in kernel_ms); prints a reduced JSON line")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

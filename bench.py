#!/usr/bin/env python
"""Benchmark of the traj2dEcsv hot path on B200 (contract: see the repo prompt / DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C1|C2|C4|C5] [--frames-per-step F]

Workloads (BASELINE.json configs; C2 is the one the metric is quoted on and the default):
  C1  FMO monomer, 30 000 atoms, 7 BCL, ~320 columns                         (configs[0])
  C2  FMO trimer, 100 000 atoms, 24 BCL, 1099 columns, per-residue groups    (configs[1]; configs[2] = the same at N GPUs)
  C4  membrane-sized trimer, 350 000 atoms, every solvent molecule its own group (~1e5 columns), covariance off (configs[3])
  C5  the Correlate step alone: 24 sites x 1099 columns, rows already computed (configs[4])
All of them synthetic, fixed-column .gro text (45 B per atom line), seeds 1001-1005.  One STEP = one batch of F distinct
pre-generated frames pushed through the whole hot path: K1 text decode -> K2 CDC pair sums for all sites -> float32 dE
rows -> K3 Correlate sums (C5: K3 only).  Every step re-streams the same F frames, whose text (C2: F x 4.5 MB = 1.15 GB at
F = 256) is far larger than the 126 MB L2.

  value     frames/s, whole job, text resident in HBM when the timed region starts (cgc_run_device + K3)
  e2e       frames/s through the C-ABI call a user makes (cgc_run) with HOST buffers: pinned text in, H2D copies, kernels,
            D2H of the dE rows — all inside the timed region
  e2e_xtc   the same frames as a GROMACS .xtc (compressed integers, ~5 B/atom instead of 45), decoded on the device (K1x): the
            leg is bound by the kernels, not by the PCIe link, so it scales with the GPUs; rows bit-identical to the .gro path
  e2e_cli   the drop-in program itself (traj2dEcsv, C++): a real .gro file in /dev/shm (6144 frames = 27.6 GB at N = 1, 1024 per
            GPU at N > 1) -> csv_out.time + csv_out.mtx, whole process and frame loop; at N > 1 ONE process drives the N GPUs (traj2dEcsv --gpus N, ncclAllReduce)
  N > 1     one process per GPU (torchrun), every rank streams its own frame range (weak scaling, no data-path
            collective); the Correlate partial sums are combined by ONE NCCL all-reduce, inside the timed region;
            allreduce_check compares the merged covariance with a single-rank run over the same frames
  --impl reference   the reference's own CPU implementation on the host cores, one process per core over split
            trajectories as the authors ran it (scripts/2014-06-gro2dEcsv.sh:3): C2 = oracle/_ref/ref_harness (the
            unmodified reference library, all 24 sites, nothing extrapolated), C1 = the stock program oracle/_ref/traj2dEcsv_ref
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "frames/s"
SM_COUNT = 148
PAIRS_PER_CLK_PER_SM = 16.0   # 8 issue slots per pair (7 FP32 + 1 MUFU.RSQ): MUFU at 16 lanes/clk/SM sets the same ceiling

# name -> (metric, workload string shared by both arms, default frames per step)
WORKLOADS = {
    "C1": ("CDC frames/s (FMO monomer ~30k atoms)",
           "C1: FMO monomer 30000 atoms, 7 BCL sites x 44 dQ atoms, numTot 320, .gro text 45 B/atom, seed 1001", 512),
    "C2": ("CDC frames/s (FMO trimer ~100k atoms)",
           "C2: FMO trimer 100000 atoms, 24 BCL sites x 44 dQ atoms, numTot 1099, .gro text 45 B/atom, seed 1002", 256),
    "C4": ("CDC frames/s (membrane FMO trimer ~350k atoms)",
           "C4: membrane-sized FMO trimer 350000 atoms, 24 BCL sites x 44 dQ atoms, per-molecule groups (numTot ~1e5), "
           ".gro text 45 B/atom, seed 1004, covariance off", 32),
    "C5": ("covariance frames/s (24 sites x 1099 groups)",
           "C5: Correlate step alone, 24 sites x 1099 columns of float32 dE rows (random, large means), FP64 X^T X per site", 1024),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def shared_config(name: str, F: int, world: int) -> dict:
    """The `config` object: identical in both arms (the driver compares them)."""
    return {"workload": WORKLOADS[name][1], "frames_per_step": F,
            "l2": "per-step input exceeds the 126 MB L2 (no flush needed)",
            "pair_convention": "dQ != 0 site atoms only (44 of 140)"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=float(d.get("hbm_gbs", 6650.0)), sm_max_mhz=float(d.get("sm_max_mhz", 1965.0)), source="measured")
    return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, source="fallback")


def k2_traffic_per_frame():
    """dram__bytes_read.sum + dram__bytes_write.sum of cdc_kernel per frame of the C2 workload, from the committed
    `ncu --set full` capture; None when that file is absent."""
    for name in ("r02_k2_traffic.json", "r01_k2_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            try:
                return float(json.load(open(p))["dram_bytes_per_frame"]), "profiles/" + name
            except Exception:
                pass
    return None, None


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


def calibrate_h2d(torch, dist, world, host, d_text, seconds=0.5):
    """Yard-stick for the end-to-end leg: pinned host -> device copies of the same text buffer.  Returns
    (alone_gbs, [concurrent_gbs of every rank]): best single copy on this rank, and — the ceiling the e2e leg is held
    against — the rate every rank sustains while ALL ranks copy at the same time (barrier-synchronised start, `seconds` of
    back-to-back copies)."""
    nbytes = host.numel()
    best = 1e9
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        d_text.copy_(host, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    alone = nbytes / (best * 1e-3) / 1e9
    reps = max(2, int(seconds * alone * 1e9 / nbytes))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        d_text.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    conc = reps * nbytes / dt / 1e9
    rates = [conc]
    if world > 1:
        t = torch.zeros(world, dtype=torch.float64, device=d_text.device)
        t[dist.get_rank()] = conc
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        rates = [float(v) for v in t.tolist()]
    return alone, rates


def cores_available() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ----------------------------------------------------------------------------------------------------------------------
# synthetic inputs
# ----------------------------------------------------------------------------------------------------------------------
def make_inputs(name: str, workdir: str, n_frames: int, first_frame: int = 0):
    """Returns (system, top, itp, text bytes of n_frames frames, frame_bytes)."""
    from cgromcorrfmo_b200 import synth
    s = synth.named_system("C2" if name == "C5" else name)
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
def time_reference_cpu(name: str, workdir: str, n_frames: int, warm_frames: int, n_procs: int):
    """Runs `n_procs` copies of the UNMODIFIED reference, one per core, each over its own pass of the trajectory file (the
    authors' parallelisation, reference scripts/2014-06-gro2dEcsv.sh:3), EVERY site of every frame evaluated.
      C2: oracle/_ref/ref_harness (the reference library's ReadOneTimeGrom2Atom + AtomDataLookup_v2 + ComputeCDC_v1 for all
          24 sites; the stock program stops at 7).  Start-up (two ParseTopology passes, O(T^2), 60-90 s at T = 1e5) is not
          part of the frame loop and is reported separately.
      C1: the stock program oracle/_ref/traj2dEcsv_ref itself (7 sites = all of C1's), whole process minus a 1-frame run."""
    from oracle import oracle as O
    if not O.have_ref():
        return None
    from cgromcorrfmo_b200 import synth
    s = synth.named_system(name)
    top, itp = os.path.join(workdir, "ref_topol.top"), os.path.join(workdir, "ref_bcx.itp")
    synth.write_top(top, s)
    synth.write_itp(itp, s)
    prefix = synth._static_prefix(s)
    distinct = [synth.frame_bytes(s, prefix, f) for f in range(min(4, n_frames))]
    traj = os.path.join(workdir, "ref_traj.gro")
    with open(traj, "wb") as f:
        for i in range(n_frames):
            f.write(distinct[i % len(distinct)])
    n_sites = int(s.is_bcl.sum()) // 140
    t0 = time.time()
    if name == "C1":
        exe = os.path.join(O.REF_DIR, "traj2dEcsv_ref")

        def run_all(frames):
            tt = time.time()
            procs = [subprocess.Popen(["bash", "-c", "ulimit -s unlimited 2>/dev/null; exec '%s' '%s' '%s' '%s' '%s' %d" % (
                exe, traj, top, itp, os.path.join(workdir, "ref_out_%d" % p), frames)], stdout=subprocess.DEVNULL,
                stderr=subprocess.DEVNULL) for p in range(n_procs)]
            rcs = [pr.wait() for pr in procs]
            return time.time() - tt, rcs
        base_s, rcs0 = run_all(max(1, warm_frames))           # start-up + the warm-up frames
        full_s, rcs1 = run_all(n_frames)
        if any(rcs0) or any(rcs1) or full_s <= base_s:
            return None
        timed = n_frames - max(1, warm_frames)
        fps = n_procs * timed / (full_s - base_s)
        return dict(value=float(fps), unit=UNIT, cores=n_procs, kind="reference",
                    sample="%d processes of the stock traj2dEcsv (unmodified reference, 7 sites = all of C1) x %d frames: wall "
                           "%.1f s minus %.1f s for a %d-frame run (start-up); every process writes its own .time/.mtx"
                           % (n_procs, n_frames, full_s, base_s, max(1, warm_frames)),
                    frames_per_s_per_core=float(fps / n_procs), startup_s=float(base_s))
    harness = os.path.join(O.REF_DIR, "ref_harness")
    procs = []
    for p in range(n_procs):
        out = os.path.join(workdir, "ref_out_%d" % p)
        procs.append((out, subprocess.Popen([harness, traj, top, itp, out, str(n_frames), str(n_sites), "0"],
                                            stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)))
    per_proc = []
    for out, pr in procs:
        pr.wait()
        if pr.returncode != 0:
            continue
        rows = [[float(v) for v in ln.split()] for ln in open(out + ".frames.txt").read().strip().split("\n")]
        tm = [float(v) for v in open(out + ".timing.txt").read().split()]
        if int(tm[6]) != n_sites:
            continue
        rows = rows[warm_frames:] if len(rows) > warm_frames else rows
        frame_s = np.mean([r[0] + r[1] + r[2] for r in rows])            # read + lookup + cdc, all sites, nothing scaled
        per_proc.append((frame_s, tm[5], np.mean([r[0] for r in rows]), np.mean([r[1] + r[2] for r in rows])))
    wall = time.time() - t0
    if not per_proc:
        return None
    fps = float(sum(1.0 / p[0] for p in per_proc))
    return dict(value=fps, unit=UNIT, cores=len(per_proc), kind="reference",
                sample="%d processes x %d frames of %s (after %d warm-up), all %d sites of every frame evaluated by the unmodified "
                       "reference library (read %.2f s + lookup/CDC %.2f s per frame per core; the stock program itself stops at 7 "
                       "sites); start-up %.0f s per process (O(T^2) ParseTopology) excluded; wall %.0f s"
                       % (len(per_proc), n_frames - warm_frames, name, warm_frames, n_sites, per_proc[0][2], per_proc[0][3],
                          per_proc[0][1], wall),
                frames_per_s_per_core=float(np.mean([1.0 / p[0] for p in per_proc])), startup_s=float(per_proc[0][1]))


def time_c5_cpu(F: int, steps: int):
    """C5 on the host: depca's np.cov formulation (scripts/depca/depca/sidechain_corr.py:50-72) as FP64 X^T X per site with
    numpy's threaded BLAS — a port (the reference's own float32 rank-1 loop, src/FMOparse.cpp:168-178, is not callable
    outside main_Cr)."""
    rng = np.random.default_rng(1005)
    X = rng.normal(size=(F, 24, 1099)).astype(np.float32)
    acc = np.zeros((24, 1099, 1099))
    t0 = time.perf_counter()
    for _ in range(steps):
        for k in range(24):
            d = X[:, k, :].astype(np.float64)
            acc[k] += d.T @ d
    dt = time.perf_counter() - t0
    return dict(value=steps * F / dt, unit=UNIT, cores=cores_available(), kind="port",
                sample="%d steps of %d frames: per site float64 X^T X with numpy's threaded BLAS (depca's np.cov formulation)" % (steps, F))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    name = args.workload
    F = args.frames_per_step or WORKLOADS[name][2]
    n_procs = cores_available()
    t0 = time.time()
    if name == "C5":
        r = time_c5_cpu(F, max(1, min(args.steps, 3)))
    elif name == "C4":
        r = None
        why = ("the reference scans nGroups x N group memberships per site call (src/grom2atomFMO.cpp:669-674): ~3.5e10 checks "
               "per site at C4's 1e5 groups, minutes per frame — not run")
    else:
        n_frames = max(2, min(args.warmup + args.steps, 24))
        warm = min(args.warmup, n_frames - 1)
        with tempfile.TemporaryDirectory(prefix="cgc_ref_") as d:
            r = time_reference_cpu(name, d, n_frames, warm, n_procs)
        why = "oracle/_ref (the compiled reference) is missing on this box"
    wall = time.time() - t0
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": why}))
        return 0
    line = {
        "impl": "reference", "metric": WORKLOADS[name][0], "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["cores"] / r["value"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": shared_config(name, F, 1),
        "step_detail": "one frame per host process (%d processes), all sites" % r["cores"],
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "startup_s_per_process": r.get("startup_s"),
        "gpu_launches": 0, "wall_s": wall,
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------------------------------
# e2e_cli: the drop-in program on a real file
# ----------------------------------------------------------------------------------------------------------------------
def write_shared_trajectory(path, blob_text, frame_bytes, frames_per_rank, rank, world, barrier):
    """Every rank writes its share of ONE trajectory file (its own distinct frames, cycled) at the right offset."""
    n_distinct = len(blob_text) // frame_bytes
    if rank == 0:
        with open(path, "wb") as f:
            f.truncate(frame_bytes * frames_per_rank * world)
    barrier()
    fd = os.open(path, os.O_WRONLY)
    try:
        off = rank * frames_per_rank * frame_bytes
        left = frames_per_rank
        while left > 0:
            n = min(left, n_distinct)
            os.pwrite(fd, blob_text[: n * frame_bytes], off)
            off += n * frame_bytes
            left -= n
    finally:
        os.close(fd)
    barrier()


def run_cli(traj, top, itp, out, frames, devices, extra=()):
    from cgromcorrfmo_b200 import build
    multi = ["--devices", ",".join(str(d) for d in devices)] if len(devices) > 1 else ["--device", str(devices[0])]
    cmd = [build.CLI, traj, top, itp, out, str(frames), "--timing"] + multi + list(extra)
    t0 = time.perf_counter()
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    except subprocess.TimeoutExpired:
        return dict(rc=-9, wall_s=time.perf_counter() - t0, in_loop=None, timing="timed out after 240 s")
    wall = time.perf_counter() - t0
    m = re.search(r"in-loop (\d+) frames/s", r.stderr)
    return dict(rc=r.returncode, wall_s=wall, in_loop=float(m.group(1)) if m else None, timing=r.stderr.strip().split("\n")[-1][-900:])


def cli_leg(shm_dir, top, itp, text, frame_bytes, frames_per_rank, rank, world, barrier, expect_rows, S, G, devices):
    """traj2dEcsv on a real .gro file in the page cache (tmpfs).  Rank 0 runs it (N > 1: ONE process, --gpus N) while the
    other ranks stay out of the way; returns the result dict on rank 0."""
    from cgromcorrfmo_b200 import capi
    traj = os.path.join(shm_dir, "cgc_bench_traj.gro")
    marker = traj + ".done"
    if rank == 0 and os.path.exists(marker):
        os.remove(marker)
    write_shared_trajectory(traj, text, frame_bytes, frames_per_rank, rank, world, barrier)
    res = None
    if rank == 0:
        try:
            total = frames_per_rank * world
            out = os.path.join(shm_dir, "cgc_bench_out")
            runs = [run_cli(traj, top, itp, out, total, devices) for _ in range(2)]   # the first run also warms the driver caches
            best = min((r for r in runs if r["rc"] == 0), key=lambda r: r["wall_s"], default=None)
            res = {"runs": runs}
            if best is not None:
                # the rows the CLI printed for rank 0's first frames against cgc_run's rows for the same frames
                n_chk = min(expect_rows.shape[0], frames_per_rank)
                want = os.path.join(shm_dir, "cgc_bench_want.time")
                capi.write_time(want, expect_rows[:n_chk].reshape(n_chk * S, G), truncate=True)
                wb = open(want, "rb").read()
                with open(out + ".time", "rb") as f:
                    gb = f.read(len(wb))
                size = os.path.getsize(out + ".time")
                res.update({
                    "value": total / best["wall_s"], "unit": UNIT, "in_loop_frames_per_s": best["in_loop"], "frames": total,
                    "whole_process_s": best["wall_s"], "gpus_in_one_process": world,
                    "input": "%.2f GB .gro in %s" % (total * frame_bytes / 1e9, shm_dir),
                    "outputs": "csv_out.time %.2f GB + csv_out.mtx, also in %s" % (size / 1e9, shm_dir),
                    "time_text_matches_cgc_run_rows": bool(gb == wb), "time_text_bytes_checked": len(wb),
                    "api": "traj2dEcsv traj_in topology excited_itp csv_out num_frames" + (" --devices %s" % ",".join(str(d) for d in devices) if world > 1 else ""),
                    "timing": best["timing"]})
                os.remove(want)
            for suffix in (".time", ".mtx"):
                if os.path.exists(out + suffix):
                    os.remove(out + suffix)
        finally:
            open(marker, "w").close()
    else:
        # stay off the GPUs and the cores while rank 0's single process drives all N GPUs: no NCCL barrier kernel spinning
        deadline = time.time() + 900
        while not os.path.exists(marker) and time.time() < deadline:
            time.sleep(0.05)
    barrier()
    if rank == 0:
        os.remove(traj)
        os.remove(marker)
    return res


# ----------------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------------
def device_for_rank(local_rank: int, world: int, visible: int) -> int:
    """LOCAL_RANK -> CUDA device.  CGC_BENCH_DEVICES="a,b,c,.." names the devices to use (rank r takes the r-th); otherwise
    identity, except on an 8-GPU box with fewer ranks than GPUs (below).  The end-to-end legs are bound by host->device
    copies, and which GPUs share a host-side bottleneck is a property of the box (scripts/h2d_matrix.cu measures it)."""
    env = os.environ.get("CGC_BENCH_DEVICES", "")
    if env:
        devs = [int(x) for x in env.split(",") if x.strip() != ""]
        if len(devs) >= world and all(0 <= d < visible for d in devs):
            return devs[local_rank]
    if 1 < world < visible and visible == 8:
        # This pool's 8-GPU boxes (profiles/r02_h2d_matrix_8gpu.txt): GPUs 0-3 share one ~115 GB/s host path (28.9 GB/s each
        # when all four copy), GPUs 4-7 keep 55 GB/s each when all four copy.  Fewer ranks than GPUs: take the upper ones.
        return visible - world + local_rank
    return local_rank


def run_ours(args):
    import torch
    import torch.distributed as dist
    from cgromcorrfmo_b200 import capi, dist as cdist, synth

    name = args.workload
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path is CUDA-only, there is no CPU fallback")
    local = device_for_rank(local, world, torch.cuda.device_count())
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cores = cdist.bind_to_gpu_numa(local) if world > 1 else None   # before any pinned allocation
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    F = args.frames_per_step or WORKLOADS[name][2]
    peaks = measured_peaks()
    if name == "C5":
        return run_c5(args, torch, dist, capi, world, rank, local, dev, F, peaks)
    cov_on = name != "C4"

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with tempfile.TemporaryDirectory(prefix="cgc_bench_") as d:
        t0 = time.time()
        # every rank streams its own frame range of the synthetic trajectory: frames [rank*F, (rank+1)*F)
        s, top, itp, blob, frame_bytes = make_inputs(name, d, F, first_frame=rank * F)
        text = np.frombuffer(blob, dtype=np.uint8)
        n_text = text.size
        host = torch.empty(n_text + 64, dtype=torch.uint8).pin_memory()
        host[:n_text] = torch.from_numpy(text.copy())
        host[n_text:] = 0
        if rank == 0:
            log("bench: generated %d frames of %s (%.1f MB text) in %.1f s" % (F, name, n_text / 1e6, time.time() - t0))

        ctx = capi.Context(top, itp, local)
        ctx.open_memory((host.data_ptr(), n_text))
        n_atoms, n_chromo, n_groups, G = ctx.dims()
        S = ctx.n_sites
        ctx.set_option("batch_frames", F)
        ctx.set_option("profile", 1)
        if world > 1:
            ctx.set_option("stage_threads", max(2, min(16, cores_available() // world)))
        ndq = int(ctx.get_option("site_cap"))
        n_c = int((ctx.groups() == -1).sum())
        pairs_per_frame = S * ndq * (n_atoms - n_c)                 # dQ != 0 convention
        pairs_per_frame_140 = S * n_c * (n_atoms - n_c)             # the reference's own count (all 140 site atoms)

        # Correlate partial sums live in a torch tensor so torch.distributed can all-reduce it in place
        cov = None
        if cov_on:
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
        if cov_on:
            ctx.cov_set_shift_device(row0.data_ptr(), stream)

        def step_device():
            ctx.run_device(d_text.data_ptr(), n_text, offs, F, d_dE.data_ptr(), stream)
            if cov_on:
                ctx.cov_accumulate_device(d_dE.data_ptr(), F, stream)

        def combine(buf):
            if world > 1 and buf is not None:
                cdist.allreduce_partials(buf, S, G)                  # ONE collective: n, sum(x-c), sum (x-c)(x-c)^T

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
        combine(cov)
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
                print(json.dumps({"metric": WORKLOADS[name][0], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                                  "ms_per_step": ms_total / args.steps, "device_leg_only": True,
                                  "kernel_ms": {k: v[0] for k, v in ktimes.items()}, "gpu_launches": int(launches)}))
            if world > 1:
                dist.destroy_process_group()
            return 0

        # ---------------- end to end through cgc_run: pinned host text in, dE rows out to host memory
        # The leg is bound by the host->device copies, and the GPUs of a box do not get equal shares of the host's
        # bandwidth when all of them copy (8-GPU boxes of this pool: 23 GB/s each for GPUs 0-3, 35 GB/s each for 4-7).  Frames
        # are therefore dealt out in proportion to the rate each rank's link sustains under full load, measured here: rank r
        # streams F_r = F * rate_r / max(rate) frames per step (the fastest links keep F).
        h2d_alone, h2d_rates = calibrate_h2d(torch, dist, world, host, d_text)
        share = [max(1, int(round(F * r / max(h2d_rates)))) for r in h2d_rates]
        F_me = share[rank]
        out_host = np.empty((F, S, G), dtype=np.float32)
        e2e_batch = args.e2e_batch or int(min(F, max(1, (144 << 20) // frame_bytes)))
        ctx.set_option("batch_frames", e2e_batch)                    # several batches per call: H2D overlaps the kernels
        ctx.run(0, F, out=out_host)                                   # all F rows once, for the comparisons below
        for _ in range(max(3, args.warmup)):
            ctx.run(0, F_me, out=out_host)
        barrier()
        ctx.kernel_times(reset=True)
        ctx.copy_times(reset=True)
        t0 = time.perf_counter()
        e0.record()
        step_s = []
        for _ in range(args.steps):
            ts = time.perf_counter()
            ctx.run(0, F_me, out=out_host)
            step_s.append(time.perf_counter() - ts)
        if os.environ.get("CGC_BENCH_STEP_TIMES") and rank == 0:
            log("bench: e2e per-step ms: first %s ... last %s | min %.2f median %.2f max %.2f" % (
                ["%.1f" % (x * 1e3) for x in step_s[:6]], ["%.1f" % (x * 1e3) for x in step_s[-6:]],
                min(step_s) * 1e3, float(np.median(step_s)) * 1e3, max(step_s) * 1e3))
        combine(cov)
        e1.record()
        barrier()
        e2e_s = max(time.perf_counter() - t0, e0.elapsed_time(e1) / 1e3)
        if world > 1:
            t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        e2e_value = sum(share) * args.steps / e2e_s
        e2e_ktimes = ctx.kernel_times(reset=True)
        copy_ms, n_copies = ctx.copy_times(reset=True)
        link = [copy_ms / (e2e_s * 1e3), F_me * args.steps * frame_bytes / max(copy_ms, 1e-9) / 1e6]   # busy fraction, GB/s while copying
        if world > 1:
            t = torch.zeros(2 * world, dtype=torch.float64, device=dev)
            t[2 * rank], t[2 * rank + 1] = link[0], link[1]
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            link = [float(v) for v in t.tolist()]
        same = bool(np.array_equal(out_host.view(np.uint32), d_dE.cpu().numpy().view(np.uint32)))
        h2d_ceiling_fps = sum(h2d_rates) * 1e9 / frame_bytes           # every link at its concurrent rate

        # ---------------- the one collective, checked: merged sums == one rank over all the frames
        allreduce_check = None
        if world > 1 and cov_on:
            n_chk = 8
            d_chunk = d_text[: n_chk * frame_bytes].contiguous()
            gathered = [torch.empty_like(d_chunk) for _ in range(world)] if rank == 0 else None
            dist.gather(d_chunk, gathered, dst=0)
            cov_chk = torch.zeros(ctx.cov_size(), dtype=torch.float64, device=dev)
            ctx.cov_bind(cov_chk.data_ptr())
            ctx.cov_set_shift_device(row0.data_ptr(), stream)
            ctx.set_option("batch_frames", n_chk)
            ctx.run_device(d_text.data_ptr(), n_text, offs[:n_chk], n_chk, d_dE.data_ptr(), stream)
            ctx.cov_accumulate_device(d_dE.data_ptr(), n_chk, stream)
            torch.cuda.synchronize()
            cdist.allreduce_partials(cov_chk, S, G)
            torch.cuda.synchronize()
            if rank == 0:
                _, cov_merged = ctx.cov_finalize(1)
                n_merged = float(cov_chk[0].item())
                all_text = torch.cat(gathered + [torch.zeros(64, dtype=torch.uint8, device=dev)])
                cov_one = torch.zeros(ctx.cov_size(), dtype=torch.float64, device=dev)
                ctx.cov_bind(cov_one.data_ptr())
                ctx.cov_set_shift_device(row0.data_ptr(), stream)
                rel = offs[:n_chk] - (offs[0] - ctx.frame_atoms_offset(0))
                for r in range(world):
                    ctx.run_device(all_text.data_ptr(), all_text.numel() - 64, rel + r * n_chk * frame_bytes, n_chk, d_dE.data_ptr(), stream)
                    ctx.cov_accumulate_device(d_dE.data_ptr(), n_chk, stream)
                torch.cuda.synchronize()
                _, cov_single = ctx.cov_finalize(1)
                sd = np.sqrt(np.einsum("sii->si", cov_single))
                scale = np.maximum(np.abs(cov_single), sd[:, :, None] * sd[:, None, :])
                ok = scale > 0
                err = float((np.abs(cov_merged - cov_single)[ok] / scale[ok]).max())
                allreduce_check = {"ok": bool(err <= 1e-9 and n_merged == n_chk * world), "max_rel_diff": err, "tolerance": 1e-9,
                                   "frames": n_chk * world, "frames_counted_after_allreduce": n_merged,
                                   "what": "covariance (ddof 1) from the NCCL all-reduced partial sums of %d ranks x %d frames vs ONE "
                                           "context accumulating the same %d frames" % (world, n_chk, n_chk * world)}
                del all_text, cov_one
            del cov_chk
        ctx.close()

        # ---------------- the same frames as a GROMACS .trr (single precision, coordinates + box): 12 B/atom over PCIe
        # instead of 45.  Not the headline (the reference reads .gro text); SURVEY.md section 8f-3.
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
        trr_batch = args.e2e_batch or int(min(F, max(1, (40 << 20) // (n_trr // F))))
        ctx2.set_option("batch_frames", trr_batch)
        ctx2.set_option("profile", 1)
        if not cov_on:
            ctx2.set_option("covariance", 0)
        cov2 = None
        if cov_on:
            cov2 = torch.zeros(ctx2.cov_size(), dtype=torch.float64, device=dev)
            ctx2.cov_bind(cov2.data_ptr())
            ctx2.cov_set_shift_device(row0.data_ptr(), stream)            # the same shift on every rank, as above
        out_trr = np.empty((F, S, G), dtype=np.float32)
        d_trr = torch.empty(n_trr + 64, dtype=torch.uint8, device=dev)
        trr_alone, trr_rates = calibrate_h2d(torch, dist, world, host_trr, d_trr)
        del d_trr
        trr_share = [max(1, int(round(F * r / max(trr_rates)))) for r in trr_rates]
        ctx2.run(0, F, out=out_trr)
        for _ in range(max(3, args.warmup)):
            ctx2.run(0, trr_share[rank], out=out_trr)
        barrier()
        ctx2.kernel_times(reset=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            ctx2.run(0, trr_share[rank], out=out_trr)
        combine(cov2)
        e1.record()
        barrier()
        trr_s = max(time.perf_counter() - t0, e0.elapsed_time(e1) / 1e3)
        if world > 1:
            t = torch.tensor([trr_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            trr_s = float(t.item())
        trr_ktimes = ctx2.kernel_times(reset=True)
        trr_value = sum(trr_share) * args.steps / trr_s
        trr_same = bool(np.array_equal(out_trr.view(np.uint32), out_host.view(np.uint32)))
        ctx2.close()
        del cov2

        # ---------------- the same frames as a GROMACS .xtc (compressed integers, precision 1000): ~5 B/atom over PCIe, decoded
        # on the device (K1x, csrc/xtc_kernel.cu).  The leg is kernel-bound, so every rank streams F frames.  Reported, never
        # required: a failure leaves "e2e_xtc": {"failed": ...} and the line stands.
        xtc = None
        xtc_s, xtc_err = float("inf"), None
        xtc_ktimes, xtc_same, n_xtc, xtc_batch = {}, False, 0, 0
        ctx3 = cov3 = host_xtc = None
        try:
            t_enc = time.time()
            xtc_blob = b"".join(synth.xtc_frame_bytes(s, rank * F + f) for f in range(F))
            n_xtc = len(xtc_blob)
            host_xtc = torch.empty(n_xtc + 64, dtype=torch.uint8).pin_memory()
            host_xtc[:n_xtc] = torch.from_numpy(np.frombuffer(xtc_blob, dtype=np.uint8).copy())
            host_xtc[n_xtc:] = 0
            del xtc_blob
            if rank == 0:
                log("bench: %d .xtc frames encoded in %.1f s (%.2f B/atom)" % (F, time.time() - t_enc, n_xtc / F / n_atoms))
            ctx3 = capi.Context(top, itp, local)
            ctx3.open_memory_xtc((host_xtc.data_ptr(), n_xtc), struct_gro)
            # a call of F frames in four batches: the copy and the kernels of consecutive batches overlap inside the call
            # (every launch of K1x pays the latency of its index walk once, so long runs use batches of up to 512 frames)
            xtc_batch = args.e2e_batch or int(min(F, 64))
            ctx3.set_option("batch_frames", xtc_batch)
            ctx3.set_option("profile", 1)
            if not cov_on:
                ctx3.set_option("covariance", 0)
            if cov_on:
                cov3 = torch.zeros(ctx3.cov_size(), dtype=torch.float64, device=dev)
                ctx3.cov_bind(cov3.data_ptr())
                ctx3.cov_set_shift_device(row0.data_ptr(), stream)
            out_xtc = np.empty((F, S, G), dtype=np.float32)
            ctx3.run(0, F, out=out_xtc)
            xtc_same = bool(np.array_equal(out_xtc.view(np.uint32), out_host.view(np.uint32)))
            for _ in range(max(3, args.warmup)):
                ctx3.run(0, F, out=out_xtc)
            torch.cuda.synchronize()
            ctx3.kernel_times(reset=True)
            t0 = time.perf_counter()
            e0.record()
            for _ in range(args.steps):
                ctx3.run(0, F, out=out_xtc)
            e1.record()
            torch.cuda.synchronize()
            xtc_s = max(time.perf_counter() - t0, e0.elapsed_time(e1) / 1e3)
            xtc_ktimes = ctx3.kernel_times(reset=True)
        except Exception as e:  # reported, never required
            xtc_err = "%s: %s" % (type(e).__name__, e)
            log("bench: e2e_xtc leg failed:", xtc_err)
        if world > 1:
            # the ranks first agree that the leg ran everywhere; only then the one collective (timed, added to the leg)
            t = torch.tensor([xtc_s if xtc_err is None else 1e30], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if float(t.item()) >= 1e29:
                xtc_err = xtc_err or "failed on another rank"
            elif cov3 is not None:
                e0.record()
                combine(cov3)
                e1.record()
                torch.cuda.synchronize()
                t = torch.tensor([xtc_s + e0.elapsed_time(e1) / 1e3], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            xtc_s = float(t.item())
        try:
            if ctx3 is not None:
                ctx3.close()
        except Exception:
            pass
        del cov3, host_xtc
        if xtc_err is None:
            kd_ms, kd_n = xtc_ktimes.get("decode", (0.0, 0))
            xtc = {"value": world * F * args.steps / xtc_s, "unit": UNIT, "h2d_bytes_per_step": n_xtc + 8 * F,
                   "d2h_bytes_per_step": int(out_host.nbytes),
                   "api": "cgc_open_memory_xtc + cgc_run on pinned .xtc frames (compressed integers, precision 1000, %.2f B/atom, "
                          "%d B/frame on average), %d-frame batches; decoded on the device (K1x: index pass + unpack pass)" % (
                              n_xtc / F / n_atoms, n_xtc // F, xtc_batch),
                   "note": "same frames as the headline; not the reference's input format (it reads .gro text); at N > 1 the one all-reduce of the Correlate sums is timed after the ranks agreed that the leg ran and added to the leg's time",
                   "rows_bit_identical_to_gro_path": xtc_same,
                   "kernel_ms": {k: v[0] for k, v in xtc_ktimes.items()},
                   "k1x_unpack_ms_per_launch": (kd_ms / kd_n) if kd_n else None, "k1x_launches": int(kd_n),
                   "k1x_frames_per_launch": xtc_batch, "wall_ms": xtc_s * 1e3,
                   "k1x_note": "kernel_ms.decode is the unpack pass; the index pass runs on a stream of its own beside the previous "
                               "batch's kernels (0.5-0.6 ms per launch on these frames, profiles/r02_xtc_launches.csv)"}
        else:
            xtc = {"failed": xtc_err}
        del cov, d_text, d_dE

        # ---------------- the drop-in program on a real file (tmpfs = a file in the page cache)
        cli = None
        if not args.no_cli and name in ("C1", "C2"):
            shm = "/dev/shm" if os.path.isdir("/dev/shm") else d
            per_rank = args.cli_frames or (6144 if world == 1 else 1024)   # long enough that the fill and drain of the pipeline do not dominate
            try:
                free = shutil.disk_usage(shm).free
                per_rank = int(max(F, min(per_rank, (free // 3) // (frame_bytes * world))))
                devices = [device_for_rank(r, world, torch.cuda.device_count()) for r in range(world)]
                cli = cli_leg(shm, top, itp, bytes(blob), frame_bytes, per_rank, rank, world, barrier, out_host, S, G, devices)
            except Exception as e:  # reported, never required
                log("bench: e2e_cli leg failed:", e)
                barrier()
        del blob

        fp64_peak = calibrate_fp64_tensor(torch, dev) if rank == 0 else None
        k7 = None
        if rank == 0 and name == "C2":
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
    traffic_pf, traffic_src = k2_traffic_per_frame()
    roofline = {
        "kernel": "cdc_kernel (K2, CDC pair sums)", "bound": "fp32+mufu issue (16 pairs/clk/SM)",
        "achieved": achieved, "peak": peak, "unit": "Gpairs/s", "frac": achieved / peak if peak else None,
        "traffic": (traffic_pf * F) if (traffic_pf and name == "C2") else None,
        "traffic_source": (traffic_src + " (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per frame) x frames per launch") if traffic_src else None,
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
    ]
    if cov_on:
        others.append(
            {"kernel": "cov_centre_kernel+cov_syrk_kernel (K3)", "bound": "fp64 tensor (DMMA)", "achieved": k3_flop / k3_avg_s / 1e12 if k3_avg_s else 0.0,
             "peak": fp64_peak, "unit": "TFLOP/s", "frac": (k3_flop / k3_avg_s / 1e12 / fp64_peak) if (k3_avg_s and fp64_peak) else None,
             "peak_source": "cuBLAS DGEMM 4096^3 measured in this run (no FP64 figure in MEASURED_PEAKS.json)",
             "algorithmic_flop_per_launch": k3_flop, "avg_launch_ms": k3_avg_s * 1e3,
             "share_of_kernel_time": k3_ms / kern_total if kern_total else None})
    if k7:
        others.append(k7)
    cpu = None
    if world == 1 and not args.no_cpu_baseline and name in ("C1", "C2"):
        with tempfile.TemporaryDirectory(prefix="cgc_cpu_") as d:
            try:
                cpu = time_reference_cpu(name, d, 3, 1, cores_available())
            except Exception as e:  # the baseline is reported, never required
                log("bench: cpu baseline failed:", e)
    cfg = shared_config(name, F, world)
    line = {
        "metric": WORKLOADS[name][0], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "step_detail": {"frames_per_step_all_ranks": F * world,
                        "step": "K1 decode + K2 CDC (all sites) + float32 rows" + (" + K3 Correlate" if cov_on else "") + " over one batch of frames",
                        "input_bytes_per_step": int(n_text), "dims": {"atoms": n_atoms, "sites": S, "dq_atoms_per_site": ndq, "numTot": G},
                        "parallelism": "frame-sharded x%d, one all-reduce of the Correlate sums" % world,
                        "numa": "rank 0 on CUDA device %d%s; %s" % (local, " (CGC_BENCH_DEVICES=%s)" % os.environ["CGC_BENCH_DEVICES"] if os.environ.get("CGC_BENCH_DEVICES") else "",
                                                                  ("bound to %d cores local to its GPU" % len(numa_cores)) if numa_cores else "cores unbound (the box reports one NUMA node)"),
                        "pairs_per_frame": pairs_per_frame},
        "pairs_per_s": value * pairs_per_frame, "pairs_per_s_reference_count_140": value * pairs_per_frame_140,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n_text + 8 * F), "d2h_bytes_per_step": int(F * S * G * 4),
                "api": "cgc_run (C-ABI) on pinned host text, %d-frame batches in a 6-slot ring; dE rows copied back to host" % e2e_batch,
                "rows_bit_identical_to_device_path": same, "kernel_ms": {k: v[0] for k, v in e2e_ktimes.items()}, "wall_ms": e2e_s * 1e3,
                "bound": "PCIe host->device copy of the text", "h2d_gbs": e2e_value * frame_bytes / 1e9,
                "h2d_peak_gbs": sum(h2d_rates), "frac_of_h2d_peak": e2e_value / h2d_ceiling_fps,
                "h2d_peak_source": "sum over ranks of the rate each rank's pinned cudaMemcpyAsync of the same text buffer sustains for 0.5 s "
                                   "while ALL ranks copy (barrier-synchronised start); measured in this run, just before the leg",
                "h2d_concurrent_gbs_per_rank": h2d_rates, "h2d_alone_gbs": h2d_alone,
                "frames_per_step_per_rank": share,
                "h2d_link_busy_fraction_per_rank": link[0::2], "h2d_gbs_while_copying_per_rank": link[1::2],
                "sharding": "frames dealt out in proportion to each rank's concurrent H2D rate"},
        "e2e_trr": {"value": trr_value, "unit": UNIT, "h2d_bytes_per_step": int(n_trr + 8 * F), "d2h_bytes_per_step": int(F * S * G * 4),
                    "api": "cgc_open_memory_trr + cgc_run on pinned .trr frames (float32 coordinates + box, %d B/frame), %d-frame batches"
                           % (n_trr // F, trr_batch),
                    "note": "same frames as the headline; not the reference's input format (it reads .gro text)",
                    "rows_bit_identical_to_gro_path": trr_same, "h2d_gbs": trr_value * (n_trr / F) / 1e9,
                    "h2d_peak_gbs": sum(trr_rates), "frac_of_h2d_peak": trr_value * (n_trr / F) / 1e9 / sum(trr_rates),
                    "h2d_concurrent_gbs_per_rank": trr_rates, "h2d_alone_gbs": trr_alone, "frames_per_step_per_rank": trr_share,
                    "kernel_ms": {k: v[0] for k, v in trr_ktimes.items()}, "wall_ms": trr_s * 1e3},
        "e2e_xtc": xtc,
        "e2e_cli": cli,
        "allreduce_check": allreduce_check,
        "gpu_launches": int(launches),
        "kernel_ms": {k: v[0] for k, v in ktimes.items()},
        "roofline": roofline, "rooflines_other": others, "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_c5(args, torch, dist, capi, world, rank, local, dev, F, peaks):
    """The Correlate step alone (BASELINE.json configs[4]): 24 sites x 1099 columns, F frames of rows per step."""
    name = "C5"

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with tempfile.TemporaryDirectory(prefix="cgc_bench_") as d:
        s, top, itp, blob, frame_bytes = make_inputs("C2", d, 1)
        ctx = capi.Context(top, itp, local)
        ctx.open_memory(blob)
        S, G = ctx.n_sites, ctx.num_tot
        ctx.set_option("profile", 1)
        rng = np.random.default_rng(1005 + rank)
        X = (rng.normal(size=(F, S, G)).astype(np.float32) * 0.05 + rng.normal(size=(1, S, G)).astype(np.float32) * 30.0)
        host = torch.from_numpy(X).pin_memory()
        d_x = host.to(dev)
        cov = torch.zeros(ctx.cov_size(), dtype=torch.float64, device=dev)
        ctx.cov_bind(cov.data_ptr())
        row0 = d_x[0].clone()
        if world > 1:
            dist.broadcast(row0, 0)
        stream = torch.cuda.current_stream().cuda_stream
        ctx.cov_set_shift_device(row0.data_ptr(), stream)

        def step():
            ctx.cov_accumulate_device(d_x.data_ptr(), F, stream)

        for _ in range(max(3, args.warmup)):
            step()
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
            step()
        if world > 1:
            from cgromcorrfmo_b200 import dist as cdist
            cdist.allreduce_partials(cov, S, G)
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
        # end to end: rows from pinned host memory, the frame count read back
        n_host = torch.empty(1, dtype=torch.float64).pin_memory()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            d_x.copy_(host, non_blocking=True)
            step()
            n_host.copy_(cov[:1], non_blocking=True)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        fp64_peak = calibrate_fp64_tensor(torch, dev) if rank == 0 else None
        # correctness of what was timed: the covariance of the accumulated copies of X equals np.cov of one copy
        _, c_dev = ctx.cov_finalize(0)
        ref = np.cov(X[:, 0, :].astype(np.float64), rowvar=False, ddof=0)
        sd = np.sqrt(np.diag(ref))
        err = float((np.abs(c_dev[0] - ref) / np.maximum(np.abs(ref), np.outer(sd, sd))).max()) if world == 1 else None
        ctx.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    k3_ms, k3_n = ktimes["cov"]
    k3_avg_s = (k3_ms / max(1, k3_n)) / 1e3
    k3_flop = F * S * G * (G + 1)
    ach = k3_flop / k3_avg_s / 1e12 if k3_avg_s else 0.0
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = time_c5_cpu(F, 2)
    line = {
        "metric": WORKLOADS[name][0], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": shared_config(name, F, world), "clocks": clocks,
        "e2e": {"value": world * args.steps * F / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(F * S * G * 4), "d2h_bytes_per_step": 8,
                "api": "float32 rows from pinned host memory -> cgc_cov_accumulate_device; the frame count read back"},
        "gpu_launches": int(launches), "kernel_ms": {k: v[0] for k, v in ktimes.items()},
        "covariance_max_scaled_error_vs_numpy": err,
        "roofline": {"kernel": "cov_centre_kernel+cov_syrk_kernel (K3)", "bound": "tensor", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": ach / fp64_peak if fp64_peak else None,
                     "traffic": 902321408.0 if F == 1024 else None,
                     "traffic_source": "profiles/r02_k3_cov_syrk_kernel_ncu.txt (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum of one cov_syrk launch of 1024 frames)",
                     "algorithmic_bytes_per_launch": F * S * ((G + 15) // 16 * 16) * 8 + 2 * S * G * G * 8,
                     "peak_source": "cuBLAS DGEMM 4096^3 measured in this run (no FP64 figure in MEASURED_PEAKS.json)",
                     "algorithmic_flop_per_launch": k3_flop, "avg_launch_ms": k3_avg_s * 1e3},
        "cpu_baseline": cpu,
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
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--frames-per-step", type=int, default=0, help="0 = the workload's default")
    ap.add_argument("--e2e-batch", type=int, default=0, help="frames per device batch inside cgc_run (e2e legs); 0 = sized from the frame")
    ap.add_argument("--cli-frames", type=int, default=0, help="frames per GPU of the e2e_cli leg's trajectory file (0 = default)")
    ap.add_argument("--no-cli", action="store_true")
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

"""Frame sharding across ranks and the one collective of the hot path.

Frames are independent (reference src/FMOparse.cpp:68-179 carries nothing from one frame to the next except the
Correlate sums, :173-175), so rank r of R streams the contiguous frame range [r*T/R, (r+1)*T/R) and writes its own slice
of the time series; the only exchange is ONE all-reduce(sum) over the flat buffer of partial sums
    [ n | shift c (S*G) | sum(x-c) (S*G) | sum (x-c)(x-c)^T (S*G*G) ]
whose shift block is identical on every rank (the dE row of the trajectory's frame 0) and is therefore restored after the
sum.  The reference itself has no communication layer: its authors ran one process per trajectory chunk file
(reference scripts/2014-06-gro2dEcsv.sh:3) and concatenated later (reference scripts/2014-06-csv2hdf5.py:69-74).
"""
from __future__ import annotations

import os
import shutil


def bind_to_gpu_numa(device_index: int):
    """Pin this process to the CPU cores NVML reports as local to GPU `device_index` (the host side of its PCIe link).
    Called BEFORE any pinned host memory is allocated, so that first-touch places the staging buffers on that NUMA node:
    with one rank per GPU the end-to-end rate is bound by host memory -> PCIe traffic, and a rank whose pinned text
    sits on the other socket pays the inter-socket link twice.  Returns the list of cores, or None when NVML or the
    affinity call is unavailable (the run then proceeds unbound)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            n_words = (os.cpu_count() + 63) // 64
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        finally:
            pynvml.nvmlShutdown()
        cores = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cores = sorted(set(cores) & allowed)
        if not cores:
            return None
        os.sched_setaffinity(0, cores)
        return cores
    except Exception:
        return None


def frame_range(rank: int, world: int, total_frames: int):
    """Contiguous, balanced, ordered: concatenating the ranges in rank order gives range(total_frames)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    lo = (rank * total_frames) // world
    hi = ((rank + 1) * total_frames) // world
    return lo, hi - lo


def partials_layout(n_sites: int, num_tot: int):
    """Offsets (in doubles) of the blocks of the flat partial-sum buffer."""
    sg = n_sites * num_tot
    return dict(n=0, shift=1, sum=1 + sg, sumsq=1 + 2 * sg, size=1 + 2 * sg + sg * num_tot)


def allreduce_partials(buf, n_sites: int, num_tot: int, group=None):
    """In-place all-reduce(sum) of a partial-sum buffer (a 1-D float64 torch tensor on any backend's device).
    Every rank must have used the same shift; it is put back after the sum.  Returns buf."""
    import torch.distributed as dist
    lay = partials_layout(n_sites, num_tot)
    if buf.numel() != lay["size"]:
        raise ValueError("partial-sum buffer has %d doubles, expected %d" % (buf.numel(), lay["size"]))
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return buf
    shift = buf[lay["shift"]:lay["sum"]].clone()
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    buf[lay["shift"]:lay["sum"]] = shift
    return buf


def concat_time_parts(out_prefix: str, world: int, remove: bool = True) -> str:
    """<out>.time.part0 .. part<world-1>  ->  <out>.time, in rank order (frame-major, site-minor, like the reference's
    single file, src/FMOparse.cpp:158-165)."""
    dst = out_prefix + ".time"
    with open(dst, "wb") as o:
        for r in range(world):
            p = "%s.time.part%d" % (out_prefix, r)
            with open(p, "rb") as f:
                shutil.copyfileobj(f, o, 16 << 20)
            if remove:
                os.remove(p)
    return dst

"""traj2dEcsv as a Python call, one process per GPU (launch with torchrun for more than one):

    python -m cgromcorrfmo_b200.run traj_in topology excited_itp csv_out num_frames [--sites N] [--mtx-compat]
                                    [--structure frame0.gro   (traj_in is then a GROMACS .trr or .xtc)]

Same five positionals and the same two output files as the reference program (reference README.md:25-31); with R ranks
every rank streams its frame range, writes csv_out.time.part<r>, and the Correlate sums are combined by one NCCL
all-reduce before rank 0 writes csv_out.mtx and joins the parts."""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

from . import capi, dist as cdist


def open_trajectory(c, traj, structure=None):
    """.gro text on its own; a binary .trr / compressed .xtc (by extension) with the names taken from `structure`."""
    if structure:
        return c.open_xtc(traj, structure) if traj.endswith(".xtc") else c.open_trr(traj, structure)
    if traj.endswith((".trr", ".xtc")):
        raise capi.CgcError(capi.CGC_ERR_ARG, "a %s trajectory holds no atom names: pass structure=<file.gro>" % traj[-4:])
    return c.open(traj)


def traj2dEcsv(traj, top, itp, out, num_frames, sites=0, mtx_compat=False, chunk_frames=0, device=None, structure=None):
    import torch
    import torch.distributed as tdist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0")) if device is None else device
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: the hot path is CUDA-only (no CPU fallback)")
    if world > 1 and mtx_compat:
        # the literal .mtx is built from float32 sums added frame by frame in order (reference src/FMOparse.cpp:173-175);
        # partial sums of several ranks cannot reproduce that order, and a .mtx from one rank's sums would be wrong
        raise capi.CgcError(capi.CGC_ERR_UNSUPPORTED, "--mtx-compat cannot be combined across ranks: run it on one GPU")
    torch.cuda.set_device(local)
    if world > 1:
        cdist.bind_to_gpu_numa(local)           # pinned staging buffers land on the GPU's own NUMA node
        c_threads = max(2, min(16, (os.cpu_count() or 16) // world))
    if world > 1 and not tdist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        backend = os.environ.get("CGC_DIST_BACKEND", "nccl")    # "gloo" lets two test ranks share one GPU
        if backend == "nccl":
            tdist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            tdist.init_process_group(backend)
    with capi.Context(top, itp, local) as c:
        open_trajectory(c, traj, structure)
        if sites:
            c.set_option("sites", sites)
        if mtx_compat:
            c.set_option("mtx_compat", 1)
        if world > 1:
            c.set_option("stage_threads", c_threads)            # the ranks share the host cores
        total = min(int(num_frames), c.frame_count())          # stops cleanly where the reference would overrun
        first, count = cdist.frame_range(rank, world, total)
        S, G = c.n_sites, c.num_tot
        cov = None
        if c.cov_size():
            cov = torch.zeros(c.cov_size(), dtype=torch.float64, device="cuda")
            c.cov_bind(cov.data_ptr())
        part = out + (".time.part%d" % rank if world > 1 else ".time")
        capi.write_time(part, np.empty((0, G), np.float32), truncate=True)
        step = chunk_frames or max(1, count)                    # one call: one pipeline fill and drain
        done = 0
        while done < count:
            n = min(step, count - done)
            _, got, rc = c.run_write_time(first + done, n, part, truncate=False)   # text formatted on the device (K5)
            done += got
            if rc == capi.CGC_EOF or got == 0:
                break
        if world > 1:
            if cov is not None:
                cdist.allreduce_partials(cov, S, G)
            tdist.barrier()
        if rank == 0:
            if world > 1:
                cdist.concat_time_parts(out, world)
            if cov is not None and total > 0:
                c.write_mtx(out + ".mtx", compat=mtx_compat)
        return total


def main(argv=None):
    ap = argparse.ArgumentParser(prog="cgromcorrfmo_b200.run")
    for name in ("traj_in", "topology", "excited_itp", "csv_out"):
        ap.add_argument(name)
    ap.add_argument("num_frames", type=int)
    ap.add_argument("--sites", type=int, default=0)
    ap.add_argument("--mtx-compat", action="store_true")
    ap.add_argument("--device", type=int, default=None, help="CUDA device for this rank (default: LOCAL_RANK)")
    ap.add_argument("--structure", default=None, help="a .gro frame of the same system; traj_in is then a .trr or .xtc")
    a = ap.parse_args(argv)
    traj2dEcsv(a.traj_in, a.topology, a.excited_itp, a.csv_out, a.num_frames, a.sites, a.mtx_compat, device=a.device,
               structure=a.structure)
    return 0


if __name__ == "__main__":
    sys.exit(main())

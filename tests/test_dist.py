"""CPU tests of the multi-rank host logic with the gloo backend, world_size 2 (no GPU needed): frame sharding, the
merge of Correlate partial sums by one all-reduce, and the joining of per-rank .time slices."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cgromcorrfmo_b200 import dist as cdist


def test_frame_range_partitions():
    for total in (0, 1, 7, 10, 10000, 99999):
        for world in (1, 2, 3, 4, 8):
            got = []
            for r in range(world):
                lo, n = cdist.frame_range(r, world, total)
                assert n >= 0 and (n - total / world) ** 2 <= 1.0       # balanced to within one frame
                got.extend(range(lo, lo + n))
            assert got == list(range(total))                             # ordered, no gap, no overlap
    with pytest.raises(ValueError):
        cdist.frame_range(2, 2, 10)


def _partials(X, shift):
    """numpy restatement of what K3 accumulates for rows X [T,S,G] with shift c [S,G] (test helper only)."""
    T, S, G = X.shape
    d = X.astype(np.float64) - shift[None]
    buf = np.zeros(cdist.partials_layout(S, G)["size"])
    lay = cdist.partials_layout(S, G)
    buf[0] = T
    buf[lay["shift"]:lay["sum"]] = shift.reshape(-1)
    buf[lay["sum"]:lay["sumsq"]] = d.sum(0).reshape(-1)
    buf[lay["sumsq"]:] = np.einsum("tsi,tsj->sij", d, d).reshape(-1)
    return buf


def _finalize(buf, S, G, ddof):
    lay = cdist.partials_layout(S, G)
    n = buf[0]
    c = buf[lay["shift"]:lay["sum"]].reshape(S, G)
    s = buf[lay["sum"]:lay["sumsq"]].reshape(S, G)
    ss = buf[lay["sumsq"]:].reshape(S, G, G)
    return c + s / n, (ss - np.einsum("si,sj->sij", s, s) / n) / (n - ddof)


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(123)
        T, S, G = 41, 3, 9
        X = (rng.normal(size=(T, S, G)) * 0.3 + 20.0).astype(np.float32)      # every rank builds the same rows
        shift = X[0].astype(np.float64)                                        # the row of the trajectory's frame 0
        lo, n = cdist.frame_range(rank, world, T)
        mine = torch.from_numpy(_partials(X[lo:lo + n], shift))
        cdist.allreduce_partials(mine, S, G)
        whole = _partials(X, shift)
        assert np.allclose(mine.numpy(), whole, rtol=1e-13, atol=1e-12)
        assert np.array_equal(mine.numpy()[1:1 + S * G], shift.reshape(-1))     # shift block restored, not doubled
        mean, cov = _finalize(mine.numpy(), S, G, 1)
        for k in range(S):
            assert np.allclose(cov[k], np.cov(X[:, k].astype(np.float64), rowvar=False), rtol=1e-9)
            assert np.allclose(mean[k], X[:, k].astype(np.float64).mean(0), rtol=1e-13)
        # per-rank .time slices join into frame order
        with open(os.path.join(tmp, "o.time.part%d" % rank), "w") as f:
            for t in range(lo, lo + n):
                for s in range(S):
                    f.write("%d,%d\n" % (t, s))
        dist.barrier()
        if rank == 0:
            cdist.concat_time_parts(os.path.join(tmp, "o"), world)
            rows = open(os.path.join(tmp, "o.time")).read().split()
            assert rows == ["%d,%d" % (t, s) for t in range(T) for s in range(S)]
        with pytest.raises(ValueError):
            cdist.allreduce_partials(torch.zeros(5, dtype=torch.float64), S, G)
    finally:
        dist.destroy_process_group()


def test_two_rank_merge_gloo(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)


def test_numa_binding_is_a_no_op_without_nvml():
    """bind_to_gpu_numa never raises and leaves the affinity alone when NVML / the GPU is absent."""
    import os
    from cgromcorrfmo_b200 import dist as cdist
    before = os.sched_getaffinity(0)
    cores = cdist.bind_to_gpu_numa(0)
    after = os.sched_getaffinity(0)
    if cores is None:
        assert before == after
    else:
        assert set(cores) == after and after <= before
        os.sched_setaffinity(0, before)

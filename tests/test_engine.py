"""GPU tests of the streaming engine behind cgc_run / cgc_run_write_time (csrc/capi.cu run_impl) and of the multi-GPU
paths: reproducibility of the rows (K2 adds grid-rounded partial sums: exact, order-independent), recovery after an error in the middle of a trajectory,
the stop request, single-process multi-GPU (traj2dEcsv --gpus N, cgc_cov_allreduce over NCCL).

The frame loop these replace: reference src/FMOparse.cpp:68-184; its SIGINT flag :24-29,62; the authors' one process per
trajectory chunk scripts/2014-06-gro2dEcsv.sh:3-27."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

COV_TOL = 1e-9


@pytest.fixture(scope="module")
def capi_mod():
    from cgromcorrfmo_b200 import capi
    capi.load()
    return capi


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _repeat(traj, reps, path):
    blob = open(traj, "rb").read()
    with open(path, "wb") as f:
        for _ in range(reps):
            f.write(blob)
    return path


@pytest.mark.parametrize("tag", ["tiny", "permol"])
def test_rows_bit_identical_across_runs_batch_sizes_and_ranges(tag, small_inputs, capi_mod, tmp_path):
    """Exact bin adds (partials rounded to a 2^-36 grid): the float32 rows do not depend on the order in which the device's atomics land, nor on how the
    frames were cut into batches, launches or frame ranges (a reproducible scheme)."""
    capi = capi_mod
    traj, top, itp, frames = small_inputs(tag)
    long_traj = _repeat(traj, 6, str(tmp_path / "long.gro"))
    total = frames * 6
    runs = []
    for batch in (0, 1, 2, 5, 64):
        for rep in range(2 if batch == 0 else 1):
            with capi.Context(top, itp, 0) as c:
                c.open(long_traj)
                if batch:
                    c.set_option("batch_frames", batch)
                dE, done, rc = c.run(0, total)
                assert rc == capi.CGC_OK and done == total
                runs.append(dE.copy())
    for r in runs[1:]:
        assert np.array_equal(r.view(np.uint32), runs[0].view(np.uint32))
    # the same frames repeat in the file: so do their rows, bit for bit
    assert np.array_equal(runs[0][:frames].view(np.uint32), runs[0][frames:2 * frames].view(np.uint32))
    # a frame range that starts in the middle (what a second rank does)
    with capi.Context(top, itp, 0) as c:
        c.open(long_traj)
        part, done, _ = c.run(total // 2, total - total // 2)
    assert np.array_equal(part.view(np.uint32), runs[0][total // 2:].view(np.uint32))


def test_error_in_a_later_frame_leaves_a_consistent_state(small_inputs, capi_mod, tmp_path):
    """A malformed frame in the middle of the range: the call fails with the reference's exception class, and the frames
    delivered before it are complete and consistent everywhere — rows, frames_done, Correlate count.  The next call on the
    same context, with a smaller output buffer, neither replays nor overruns anything (round-1 advisor finding)."""
    capi = capi_mod
    traj, top, itp, frames = small_inputs("tiny")
    blob = open(traj, "rb").read()
    one = len(blob) // frames
    reps = 4
    total = frames * reps
    bad = 8
    data = bytearray(blob * reps)
    # second line of frame `bad` = the atom count: make it disagree with frame 0
    p0 = bad * one
    e_title = data.index(b"\n", p0)
    e_count = data.index(b"\n", e_title + 1)
    count = data[e_title + 1:e_count]
    data[e_title + 1:e_count] = (b"%*d" % (len(count), int(count) - 1))
    path = str(tmp_path / "broken.gro")
    open(path, "wb").write(bytes(data))
    with capi.Context(top, itp, 0) as c:
        c.open(_repeat(traj, reps, str(tmp_path / "sound.gro")))
        want, _, _ = c.run(0, total)
    with capi.Context(top, itp, 0) as c:
        c.open(path)
        c.set_option("batch_frames", 2)
        out = np.full((total,) + want.shape[1:], np.float32(7e7))
        with pytest.raises(capi.CgcError) as ei:
            c.run(0, total, out=out)
        assert ei.value.code == capi.CGC_ERR_INPUT_MALFORMED
        done = c.last_frames_done
        assert 1 <= done <= bad
        assert np.array_equal(out[:done].view(np.uint32), want[:done].view(np.uint32))
        assert np.all(out[done:] == np.float32(7e7))              # nothing written past what was delivered
        assert c.cov_partials()[0] == done                         # the Correlate sums cover the same frames
        # a second call with a buffer for ONE frame: no stale batch may be drained into it
        guard = np.full((3,) + want.shape[1:], np.float32(-5e5))
        one_row, d2, rc = c.run(0, 1, out=guard[1:2])
        assert rc == capi.CGC_OK and d2 == 1
        assert np.array_equal(guard[1].view(np.uint32), want[0].view(np.uint32))
        assert np.all(guard[0] == np.float32(-5e5)) and np.all(guard[2] == np.float32(-5e5))
        assert c.cov_partials()[0] == done + 1


def test_stop_request(small_inputs, capi_mod, tmp_path):
    """cgc_request_stop: the run ends with CGC_STOPPED after the batches under way; the flag is consumed."""
    capi = capi_mod
    traj, top, itp, frames = small_inputs("tiny")
    with capi.Context(top, itp, 0) as c:
        c.open(traj)
        c.request_stop()
        dE, done, rc = c.run(0, frames)
        assert rc == capi.CGC_STOPPED and done == 0
        dE, done, rc = c.run(0, frames)
        assert rc == capi.CGC_OK and done == frames
        tp = str(tmp_path / "t.time")
        c.request_stop()
        _, done, rc = c.run_write_time(0, frames, tp)
        assert rc == capi.CGC_STOPPED and done == 0 and os.path.getsize(tp) == 0


def test_write_time_one_call_equals_chunked_calls(small_inputs, capi_mod, tmp_path):
    """The device-formatted `.time` text does not depend on how the range is cut into calls or batches."""
    capi = capi_mod
    traj, top, itp, frames = small_inputs("permol")
    long_traj = _repeat(traj, 5, str(tmp_path / "long.gro"))
    total = frames * 5
    a, b = str(tmp_path / "a.time"), str(tmp_path / "b.time")
    with capi.Context(top, itp, 0) as c:
        c.open(long_traj)
        _, done, _ = c.run_write_time(0, total, a)
        assert done == total
        mean_a, cov_a = c.cov_finalize(0)
    with capi.Context(top, itp, 0) as c:
        c.open(long_traj)
        c.set_option("batch_frames", 3)
        done = 0
        while done < total:
            _, got, _ = c.run_write_time(done, min(4, total - done), b, truncate=(done == 0))
            done += got
        mean_b, cov_b = c.cov_finalize(0)
    assert open(a, "rb").read() == open(b, "rb").read()
    scale = np.maximum(np.abs(cov_a), np.sqrt(np.einsum("sii,sjj->sij", cov_a, cov_a)))
    ok = scale > 0
    assert np.all(np.abs(cov_a - cov_b)[ok] <= COV_TOL * scale[ok])
    assert np.allclose(mean_a, mean_b, rtol=1e-13, atol=0)


# ----------------------------------------------------------------------------------------------------------------------
# several GPUs (skipped on a one-GPU box; run with gpurun --gpus 2)
# ----------------------------------------------------------------------------------------------------------------------
def test_allreduce_of_two_contexts_over_nccl(small_inputs, capi_mod, tmp_path):
    """cgc_cov_allreduce: two contexts on two GPUs of this process, each over its own frame range; after the ONE NCCL
    all-reduce both hold the sums of all frames: the covariance equals the single-context one to 1e-9."""
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs")
    capi = capi_mod
    traj, top, itp, frames = small_inputs("permol")
    long_traj = _repeat(traj, 8, str(tmp_path / "long.gro"))
    total = frames * 8
    with capi.Context(top, itp, 0) as c:
        c.open(long_traj)
        c.run(0, total, want_dE=False)
        mean1, cov1 = c.cov_finalize(1)
    c0, c1 = capi.Context(top, itp, 0), capi.Context(top, itp, 1)
    try:
        c0.open(long_traj)
        c1.open(long_traj)
        half = total // 2 + 1
        c0.run(0, half, want_dE=False)
        c1.run(half, total - half, want_dE=False)
        capi.cov_allreduce([c0, c1])
        for c in (c0, c1):
            assert c.cov_partials()[0] == total
            mean2, cov2 = c.cov_finalize(1)
            scale = np.maximum(np.abs(cov1), np.sqrt(np.einsum("sii,sjj->sij", cov1, cov1)))
            ok = scale > 0
            assert np.all(np.abs(cov2 - cov1)[ok] <= COV_TOL * scale[ok])
            assert np.allclose(mean2, mean1, rtol=1e-12, atol=0)
    finally:
        c0.close()
        c1.close()


def test_cli_two_gpus_in_one_process(small_inputs, tmp_path):
    """traj2dEcsv --gpus 2: same `.time` bytes as the one-GPU run (the rows are bit-reproducible), `.mtx` equal to the
    printed digits' accuracy; replaces the authors' bash loop over trajectory chunks."""
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs")
    from cgromcorrfmo_b200 import build
    traj, top, itp, frames = small_inputs("tiny")
    long_traj = _repeat(traj, 7, str(tmp_path / "long.gro"))
    total = frames * 7
    one, two = str(tmp_path / "one"), str(tmp_path / "two")
    r = subprocess.run([build.CLI, long_traj, top, itp, one, str(total)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([build.CLI, long_traj, top, itp, two, str(total + 5), "--gpus", "2", "--timing"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "ended after %d of %d" % (total, total + 5) in r.stderr
    assert open(one + ".time", "rb").read() == open(two + ".time", "rb").read()
    assert not os.path.exists(two + ".time.part0") and not os.path.exists(two + ".time.part1")
    load = lambda p: np.array([[float(v) for v in ln.split(",")] for ln in open(p).read().strip().split("\n")])
    ma, mb = load(one + ".mtx"), load(two + ".mtx")
    assert ma.shape == mb.shape and np.all(np.abs(ma - mb) <= 2e-6 * np.abs(ma).max())
    r = subprocess.run([build.CLI, long_traj, top, itp, two, str(total), "--gpus", "2", "--mtx-compat"], capture_output=True, text=True)
    assert r.returncode == 1 and "one GPU" in r.stderr


# ----------------------------------------------------------------------------------------------------------------------
# What stands in for compute-sanitizer (closed on the B200 pool, profiles/r02_compute_sanitizer_refused.log): guard bands
# around every device buffer, and a stress run whose results must be bit-identical — a race in K2's lock-free ring or its
# per-warp reduction rows would show up as a changed bit.
# ----------------------------------------------------------------------------------------------------------------------
def test_guard_bands_stay_intact(small_inputs, tmp_path):
    """Every hot-path kernel, all its buffers with 4 KiB guard bands on both sides (CGC_GUARD=1, own process): no band changes.
    Covers K1 (text and .trr modes), K2, rows, K3 (several launch sizes), K5 text formatting, the device entry points."""
    import sys
    traj, top, itp, frames = small_inputs("permol")
    traj2, top2, itp2, frames2 = small_inputs("tiny")
    script = r'''
import sys, os, numpy as np
sys.path.insert(0, %r)
from cgromcorrfmo_b200 import capi, synth
import torch
checked = 0
for traj, top, itp, frames, reps in ((%r, %r, %r, %d, 7), (%r, %r, %r, %d, 30)):
    blob = open(traj, "rb").read() * reps
    path = os.path.join(%r, "g_%%d.gro" %% reps)
    open(path, "wb").write(blob)
    total = frames * reps
    for batch in (1, 3, 0):
        with capi.Context(top, itp, 0) as c:
            c.open(path)
            if batch:
                c.set_option("batch_frames", batch)
            c.set_option("mtx_compat", 1)
            dE, done, rc = c.run_write_time(0, total, path + ".time", want_dE=True)
            assert done == total
            c.cov_finalize(1)
            c.write_mtx(path + ".mtx")
            text = np.frombuffer(blob, dtype=np.uint8)
            d_text = torch.from_numpy(np.concatenate([text, np.zeros(((-text.size) %% 16) + 16, np.uint8)])).cuda()
            offs = [c.frame_atoms_offset(f) for f in range(total)]
            d_dE = torch.empty((total, c.n_sites, c.num_tot), dtype=torch.float32, device="cuda")
            c.run_device(d_text.data_ptr(), text.size, offs, total, d_dE.data_ptr())
            c.cov_accumulate_device(d_dE.data_ptr(), total)
            xyz = torch.empty((total, c.dims()[0], 3), dtype=torch.float32, device="cuda")
            c.decode_device(d_text.data_ptr(), text.size, offs, total, xyz.data_ptr())
            torch.cuda.synchronize()
            assert np.array_equal(d_dE.cpu().numpy().view(np.uint32), dE.view(np.uint32))
            n, bad = capi.debug_guard_check()
            assert n >= 10 and bad == 0, (n, bad)
            checked += n
print("GUARDS_OK", checked)
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), traj, top, itp, frames, traj2, top2, itp2, frames2, str(tmp_path))
    r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, env=dict(os.environ, CGC_GUARD="1"))
    assert r.returncode == 0 and "GUARDS_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_stress_many_launch_shapes_bit_identical(small_inputs, capi_mod, tmp_path):
    """K2's warps run free on an mbarrier ring and reuse per-warp shared-memory rows behind __syncwarp; K3 is a persistent
    kernel on a TMA ring.  A hazard in either shows up as a result that depends on timing: 40 runs over different batch sizes
    (hence unit ranges, chunk sizes, CTA counts, gather boundaries) must give bit-identical rows and covariance sums."""
    capi = capi_mod
    traj, top, itp, frames = small_inputs("nosolv")
    long_traj = _repeat(traj, 24, str(tmp_path / "long.gro"))
    total = frames * 24
    ref_rows = ref_cov = None
    for rep in range(40):
        batch = (1, 2, 3, 5, 7, 11, 16, 48)[rep % 8]
        with capi.Context(top, itp, 0) as c:
            c.open(long_traj)
            c.set_option("batch_frames", batch)
            dE, done, _ = c.run(0, total)
            assert done == total
            part = c.cov_partials()
        if ref_rows is None:
            ref_rows = dE.copy()
        assert np.array_equal(dE.view(np.uint32), ref_rows.view(np.uint32)), "rows differ at repetition %d (batch %d)" % (rep, batch)
        # the Correlate sums depend on how the frames were cut into K3 launches only through FP64 rounding of exact sums of
        # products: equal to 1e-12; for the SAME batch size they must be bit-identical
        if rep < 8:
            ref_cov = ref_cov or {}
            ref_cov[batch] = part.copy()
        else:
            assert np.array_equal(part.view(np.uint64), ref_cov[batch].view(np.uint64)), "sums differ at repetition %d" % rep

"""Developer check of the .xtc path on a B200: C2-shaped frames through cgc_run from pinned memory, kernel times of K1x (index
pass + unpack pass, one CUDA-event pair around both) for several batch sizes and both index kernels, rows compared with the
.gro path.   python scripts/dev_xtc_check.py [frames [modes [batches [water]]]]

With a fourth argument "water" the solvent atoms of every frame are first moved into water-like triples (O where the atom was,
two H 0.09-0.1 nm away): the streams then hold runs and set flag bits the way a real solvated trajectory does, which is what
the index walk's cost depends on (the bench's synthetic frames are isolated atoms: few flags, long wide steps).  Rows are
not compared in that mode (tests/test_xtc.py covers such streams)."""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cgromcorrfmo_b200 import capi, synth  # noqa: E402


def main():
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["0", "1"]      # 0: warp-speculative index pass, 1: scalar
    batches = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [32, 64, 128, 256]
    water = len(sys.argv) > 4 and sys.argv[4] == "water"
    s = synth.named_system("C2")
    with tempfile.TemporaryDirectory() as d:
        top, itp, gro = os.path.join(d, "t.top"), os.path.join(d, "b.itp"), os.path.join(d, "f0.gro")
        synth.write_top(top, s)
        synth.write_itp(itp, s)
        prefix = synth._static_prefix(s)
        text = b"".join(synth.frame_bytes(s, prefix, f) for f in range(8))
        open(gro, "wb").write(text[: len(text) // 8])
        t0 = time.time()
        if water:
            rng = np.random.default_rng(99)
            sol = np.nonzero(np.array([n == "SOL" for n in s.resname]))[0]
            sol = sol[: sol.size // 3 * 3].reshape(-1, 3)
            frames = []
            for f in range(F):
                m = synth.frame_milli(s, f)
                v = rng.normal(size=(sol.shape[0], 2, 3))
                v *= (rng.uniform(90, 100, (sol.shape[0], 2, 1)) / np.linalg.norm(v, axis=2, keepdims=True))
                m[sol[:, 1]] = m[sol[:, 0]] + np.round(v[:, 0]).astype(np.int64)
                m[sol[:, 2]] = m[sol[:, 0]] + np.round(v[:, 1]).astype(np.int64)
                for _ in range(20):   # no two successive atoms closer than a bond: the writer's first table index follows from the closest pair
                    bad = np.nonzero(np.abs(np.diff(m, axis=0)).sum(1) < 90)[0]
                    if bad.size == 0:
                        break
                    m[bad + 1] += 150
                frames.append(capi.xtc_encode(m, step=f, time_ps=0.2 * f, box=np.eye(3) * s.box))
            blob = b"".join(frames)
        else:
            blob = b"".join(synth.xtc_frame_bytes(s, f) for f in range(F))
        print("encoded %d frames in %.1f s: %.2f B/atom" % (F, time.time() - t0, len(blob) / F / s.n_atoms), flush=True)
        host = torch.empty(len(blob) + 64, dtype=torch.uint8).pin_memory()
        host[: len(blob)] = torch.from_numpy(np.frombuffer(blob, np.uint8).copy())
        with capi.Context(top, itp, 0) as c:
            c.open_memory(text)
            ref, done, rc = c.run(0, 8)
        for scalar in modes:
            os.environ["CGC_XTC_SCALAR_INDEX"] = scalar
            for batch in batches:
                if batch > F:
                    continue
                with capi.Context(top, itp, 0) as c:
                    c.open_memory_xtc((host.data_ptr(), len(blob)), gro)
                    c.set_option("batch_frames", batch)
                    c.set_option("profile", 1)
                    out = np.empty((F, c.n_sites, c.num_tot), np.float32)
                    c.run(0, F, out=out)
                    same = "not compared (water-like frames)" if water else np.array_equal(out[:8].view(np.uint32), ref.view(np.uint32))
                    c.run(0, F, out=out)
                    c.kernel_times(reset=True)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    reps = 5
                    for _ in range(reps):
                        c.run(0, F, out=out)
                    dt = (time.perf_counter() - t0) / reps
                    kt = c.kernel_times(reset=True)
                    print("overlap %s index %s batch %3d: %7.0f frames/s e2e | per launch: K1x %.3f ms, K2 %.3f ms (%d launches) | rows == .gro path: %s | errors %d" % (
                        os.environ.get("CGC_XTC_OVERLAP", "1"), "scalar" if scalar == "1" else "warp  ", batch, F / dt, kt["decode"][0] / max(1, kt["decode"][1]),
                        kt["cdc"][0] / max(1, kt["cdc"][1]), kt["decode"][1], same, c.decode_errors()), flush=True)
        os.environ.pop("CGC_XTC_SCALAR_INDEX", None)


if __name__ == "__main__":
    main()

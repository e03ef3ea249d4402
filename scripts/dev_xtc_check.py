"""Developer check of the .xtc path on a B200: C2-shaped frames through cgc_run from pinned memory, kernel times of K1x (index
pass + unpack pass, one CUDA-event pair around both) for several batch sizes and both index kernels, rows compared with the
.gro path.   python scripts/dev_xtc_check.py [frames [modes [batches]]]"""
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
    s = synth.named_system("C2")
    with tempfile.TemporaryDirectory() as d:
        top, itp, gro = os.path.join(d, "t.top"), os.path.join(d, "b.itp"), os.path.join(d, "f0.gro")
        synth.write_top(top, s)
        synth.write_itp(itp, s)
        prefix = synth._static_prefix(s)
        text = b"".join(synth.frame_bytes(s, prefix, f) for f in range(8))
        open(gro, "wb").write(text[: len(text) // 8])
        t0 = time.time()
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
                    same = np.array_equal(out[:8].view(np.uint32), ref.view(np.uint32))
                    c.run(0, F, out=out)
                    c.kernel_times(reset=True)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    reps = 5
                    for _ in range(reps):
                        c.run(0, F, out=out)
                    dt = (time.perf_counter() - t0) / reps
                    kt = c.kernel_times(reset=True)
                    print("index %s batch %3d: %7.0f frames/s e2e | per launch: K1x %.3f ms, K2 %.3f ms (%d launches) | rows == .gro path: %s | errors %d" % (
                        "scalar" if scalar == "1" else "warp  ", batch, F / dt, kt["decode"][0] / max(1, kt["decode"][1]),
                        kt["cdc"][0] / max(1, kt["cdc"][1]), kt["decode"][1], same, c.decode_errors()), flush=True)
        os.environ.pop("CGC_XTC_SCALAR_INDEX", None)


if __name__ == "__main__":
    main()

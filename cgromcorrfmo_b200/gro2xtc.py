"""Convert a fixed-column GROMACS .gro trajectory into an .xtc with the library's writer (host code; no GPU needed):

    python -m cgromcorrfmo_b200.gro2xtc traj.gro traj.xtc [--precision 1000]

what `gmx trjconv -f traj.gro -o traj.xtc` does, for machines without GROMACS.  The .xtc stores round(x * precision) per
coordinate; with the default precision 1000 and "%8.3f" text that is exactly the number the text prints, so
`traj2dEcsv traj.xtc ... --structure traj.gro` gives the same dE rows, bit for bit, as `traj2dEcsv traj.gro ...`
(tests/test_xtc.py) while moving a tenth of the bytes.  Time stamps are taken from "t=" in the title lines, the box from the
last line of every frame.  A converter for files, not part of the hot path: the coordinates are parsed here with numpy.
"""
from __future__ import annotations

import argparse
import re
import sys

import numpy as np

from . import capi


def frames_of(path: str):
    """Yields (title, coordinate text uint8 [n_atoms, 3, width], box floats) for every complete frame of a .gro file."""
    data = np.fromfile(path, dtype=np.uint8)
    raw = data.tobytes()
    pos, n_bytes = 0, len(raw)
    while pos < n_bytes:
        e1 = raw.find(b"\n", pos)
        e2 = raw.find(b"\n", e1 + 1) if e1 >= 0 else -1
        if e1 < 0 or e2 < 0:
            return
        title = raw[pos:e1].decode(errors="replace")
        try:
            n = int(raw[e1 + 1:e2].split()[0])
        except (ValueError, IndexError):
            raise ValueError("%s: atom count expected at byte %d" % (path, e1 + 1))
        first = e2 + 1
        le = raw.find(b"\n", first)
        if le < 0:
            return
        line = le - first + 1
        end = first + n * line
        be = raw.find(b"\n", end)
        if end > n_bytes or be < 0:
            return                                                     # a cut frame: the frames before it stand
        # "%5d%-5s%5s%5d" then three fields of equal width whose decimal points are equally spaced (GROMACS' own rule)
        l0 = raw[first:le]
        dots = [m.start() for m in re.finditer(rb"\.", l0[20:])]
        if len(dots) < 3:
            raise ValueError("%s: no three coordinate fields in the first atom line of the frame at byte %d" % (path, pos))
        width = dots[1] - dots[0]
        if 20 + 3 * width > line - 1 or dots[2] - dots[1] != width:
            raise ValueError("%s: atom lines are not fixed-column" % path)
        block = data[first:end].reshape(n, line)
        if not np.all(block[:, line - 1] == 10):
            raise ValueError("%s: atom lines of different lengths in the frame at byte %d" % (path, pos))
        box = [float(v) for v in raw[end:be].split()]
        yield title, block[:, 20:20 + 3 * width].reshape(n, 3, width), box
        pos = be + 1


def field_integers(fields: np.ndarray, precision: float) -> np.ndarray:
    """round(value * precision) of fixed-width decimal fields, exactly: digits -> integer -> scaled in integer arithmetic
    when precision is the power of ten the text has decimals for, else through float64."""
    n, three, width = fields.shape
    txt = fields.reshape(n * three, width)
    is_digit = (txt >= 48) & (txt <= 57)
    neg = np.any(txt == 45, axis=1)
    dot = np.argmax(txt == 46, axis=1)
    if not np.all(txt[np.arange(txt.shape[0]), dot] == 46) or not np.all(dot == dot[0]):
        raise ValueError("decimal points are not aligned")
    ndec = width - 1 - int(dot[0])
    val = np.zeros(txt.shape[0], dtype=np.int64)
    for c in range(width):
        val = np.where(is_digit[:, c], val * 10 + (txt[:, c].astype(np.int64) - 48), val)
    val = np.where(neg, -val, val)
    scale = precision / (10.0 ** ndec)
    if scale >= 1 and float(scale).is_integer():
        out = val * int(scale)
    else:
        out = np.round(val.astype(np.float64) * scale).astype(np.int64)
    return out.reshape(n, three)


def convert(gro_path: str, xtc_path: str, precision: float = 1000.0) -> int:
    frames = 0
    with open(xtc_path, "wb") as out:
        for title, fields, box in frames_of(gro_path):
            m = re.search(r"t=\s*([-+0-9.eE]+)", title)
            t = float(m.group(1)) if m else 0.0
            b = np.zeros((3, 3), np.float32)
            for d in range(min(3, len(box))):
                b[d, d] = box[d]
            if len(box) == 9:                                          # v1(y) v1(z) v2(x) v2(z) v3(x) v3(y)
                b[0, 1], b[0, 2], b[1, 0], b[1, 2], b[2, 0], b[2, 1] = box[3:9]
            out.write(capi.xtc_encode(field_integers(fields, precision), step=frames, time_ps=t, box=b, precision=precision))
            frames += 1
    return frames


def main(argv=None):
    ap = argparse.ArgumentParser(prog="cgromcorrfmo_b200.gro2xtc", description=__doc__.split("\n\n")[0])
    ap.add_argument("gro")
    ap.add_argument("xtc")
    ap.add_argument("--precision", type=float, default=1000.0)
    a = ap.parse_args(argv)
    n = convert(a.gro, a.xtc, a.precision)
    print("%d frames written to %s" % (n, a.xtc))
    return 0


if __name__ == "__main__":
    sys.exit(main())

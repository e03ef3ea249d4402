"""Writes tests/golden/xtc_water.xtc, xtc_wide.xtc + xtc_fixture.npz: small GROMACS .xtc files produced by an encoder written
in plain Python for this purpose, NOT by the repo's own writer (cgc_xtc_encode in cgromcorrfmo_b200/csrc) or reader.

Provenance, stated plainly: no GROMACS / MDAnalysis / mdtraj / xdrfile exists in the build container (no network), so these
files are not gmx output.  The encoder follows the WRITE branch of GROMACS' xdr3dfcoord (src/gromacs/fileio/libxdrf.cpp; the
same routine is distributed as xdrfile.c), restated from its published algorithm:
    frame:  int 1995, int natoms, int step, float time, float box[9],
            int natoms, float precision, int minint[3], int maxint[3], int smallidx, int nbytes, bit stream (padded to 4)
    stream: per atom or run: the "large" triple ((x-minx) * sy + (y-miny)) * sz + (z-minz) as a little-endian number in
            bitlength(sx*sy*sz) bits (three separate fields when a size exceeds 2^24), one flag bit; flag set: 5 bits =
            run + is_smaller + 1 with run = 3 * (atoms that follow as small differences, at most 8) and is_smaller in
            {-1, 0, 1} = the step of the small-number index applied after this run; the small triples take `smallidx`
            bits each, relative to the previous atom, offset by magicints[smallidx] / 2; when a run follows, the large atom
            and the first small one change places (the O of a water is then a small difference from its first H).
    choices: the first smallidx is the first magic number not below the smallest |dx|+|dy|+|dz| between successive atoms;
            an atom starts a run while it lies within magicints[smallidx]/2 of its predecessor in every coordinate; the
            index goes up when a large atom is within magicints[min(72, smallidx+8)]/2 of its predecessor, down when every
            squared distance inside the run is below (magicints[smallidx-1]/2)^2.
  xtc_water.xtc  3 frames, precision 1000: a 400-atom system whose solvent atoms are moved into water-like triples
                 (so runs, index steps in both directions and flag bits all occur), steps 0/500/1000
  xtc_wide.xtc   2 frames, precision 100: 60 atoms scattered over +-90 000 nm — a coordinate range above 2^24, which
                 switches the large triples to three separate bit fields — with two close pairs
xtc_fixture.npz holds the integers that went in, steps, times, precisions, and the text of a one-frame .gro of the 400-atom
system (names and residue numbers for --structure); the topology / itp are regenerated from the seed.

    python tests/golden/make_xtc_fixture.py
"""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cgromcorrfmo_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SYSTEM = dict(n_atoms=400, n_bcl=2, seed=77, solvent="shared")

MAGICINTS = [
    0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512, 645, 812, 1024,
    1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384, 20642, 26007, 32768, 41285, 52015, 65536, 82570,
    104031, 131072, 165140, 208063, 262144, 330280, 416127, 524287, 660561, 832255, 1048576, 1321122, 1664510, 2097152, 2642245,
    3329021, 4194304, 5284491, 6658042, 8388607, 10568983, 13316085, 16777216]
FIRSTIDX = 9
LASTIDX = len(MAGICINTS)


class BitWriter:
    def __init__(self):
        self.acc = 0       # all bits so far as one integer
        self.n = 0

    def bits(self, nbits, value):
        assert 0 <= value < (1 << nbits) or nbits == 0
        self.acc = (self.acc << nbits) | value
        self.n += nbits

    def ints3(self, nbits, sizes, nums):
        """sendints: (n0 * s1 + n1) * s2 + n2, least significant byte first, the last byte with the bits that are left"""
        assert all(0 <= nums[i] < sizes[i] for i in range(3)), (nums, sizes)
        v = (nums[0] * sizes[1] + nums[1]) * sizes[2] + nums[2]
        assert v < (1 << nbits)
        while nbits > 8:
            self.bits(8, v & 0xFF)
            v >>= 8
            nbits -= 8
        self.bits(nbits, v)

    def tobytes(self):
        pad = (-self.n) % 8
        return ((self.acc << pad).to_bytes((self.n + pad) // 8, "big")) if self.n else b""


def encode_coords(ints):
    """ints: int [natoms, 3], natoms > 9.  Returns (minint, maxint, smallidx, payload bytes, statistics)."""
    c = [[int(v) for v in row] for row in ints]
    n = len(c)
    minint = [min(r[d] for r in c) for d in range(3)]
    maxint = [max(r[d] for r in c) for d in range(3)]
    mindiff = min(sum(abs(c[i][d] - c[i - 1][d]) for d in range(3)) for i in range(1, n))
    sizeint = [maxint[d] - minint[d] + 1 for d in range(3)]
    if (sizeint[0] | sizeint[1] | sizeint[2]) > 0xFFFFFF:
        bitsizeint = [s.bit_length() for s in sizeint]
        bitsize = 0
    else:
        bitsize = (sizeint[0] * sizeint[1] * sizeint[2]).bit_length()
    smallidx = FIRSTIDX
    while smallidx < LASTIDX - 1 and MAGICINTS[smallidx] < mindiff:
        smallidx += 1
    first_smallidx = smallidx
    maxidx = min(LASTIDX - 1, smallidx + 8)
    minidx = maxidx - 8
    smaller = MAGICINTS[max(FIRSTIDX, smallidx - 1)] // 2
    smallnum = MAGICINTS[smallidx] // 2
    sizesmall = MAGICINTS[smallidx]
    larger = MAGICINTS[maxidx] // 2
    w = BitWriter()
    stats = dict(groups=0, flags=0, up=0, down=0, runs={})
    prev = [0, 0, 0]
    prevrun = -1
    i = 0
    while i < n:
        this = c[i]
        is_small = False
        if smallidx < maxidx and i >= 1 and all(abs(this[d] - prev[d]) < larger for d in range(3)):
            is_smaller = 1
        elif smallidx > minidx:
            is_smaller = -1
        else:
            is_smaller = 0
        if i + 1 < n and all(abs(this[d] - c[i + 1][d]) < smallnum for d in range(3)):
            c[i], c[i + 1] = c[i + 1], c[i]       # the second atom of a close pair goes first
            this = c[i]
            is_small = True
        if bitsize == 0:
            for d in range(3):
                w.bits(bitsizeint[d], this[d] - minint[d])
        else:
            w.ints3(bitsize, sizeint, [this[d] - minint[d] for d in range(3)])
        prev = this
        i += 1
        run = []
        if not is_small and is_smaller == -1:
            is_smaller = 0
        while is_small and len(run) < 8:
            cur = c[i]
            if is_smaller == -1 and sum((cur[d] - prev[d]) ** 2 for d in range(3)) >= smaller * smaller:
                is_smaller = 0
            run.append([cur[d] - prev[d] + smallnum for d in range(3)])
            prev = cur
            i += 1
            is_small = i < n and all(abs(c[i][d] - prev[d]) < smallnum for d in range(3))
        runlen = 3 * len(run)
        stats["groups"] += 1
        stats["runs"][len(run)] = stats["runs"].get(len(run), 0) + 1
        if runlen != prevrun or is_smaller != 0:
            prevrun = runlen
            w.bits(1, 1)
            w.bits(5, runlen + is_smaller + 1)
            stats["flags"] += 1
        else:
            w.bits(1, 0)
        for t in run:
            w.ints3(smallidx, [sizesmall] * 3, t)
        if is_smaller != 0:
            stats["up" if is_smaller > 0 else "down"] += 1
            smallidx += is_smaller
            if is_smaller < 0:
                smallnum = smaller
                smaller = MAGICINTS[smallidx - 1] // 2
            else:
                smaller = smallnum
                smallnum = MAGICINTS[smallidx] // 2
            sizesmall = MAGICINTS[smallidx]
    return minint, maxint, first_smallidx, w.tobytes(), stats


def encode_frame(ints, step, time, box, precision):
    n = len(ints)
    head = struct.pack(">iii", 1995, n, step) + struct.pack(">f", time) + struct.pack(">9f", *np.asarray(box, float).reshape(-1))
    head += struct.pack(">i", n)
    if n <= 9:
        return head + struct.pack(">%df" % (3 * n), *(np.asarray(ints, float).reshape(-1) / precision)), None
    minint, maxint, smallidx, payload, stats = encode_coords(ints)
    body = struct.pack(">f", precision) + struct.pack(">3i", *minint) + struct.pack(">3i", *maxint) + struct.pack(">i", smallidx)
    body += struct.pack(">i", len(payload)) + payload + b"\0" * ((-len(payload)) % 4)
    return head + body, stats


def water_like(milli, is_bcl, rng):
    """Move the non-chromophore atoms after the first 60 into water-like triples: O where the atom was, two H ~0.1 nm away."""
    m = milli.copy()
    idx = np.nonzero(~is_bcl)[0][60:]
    for k in range(0, idx.size - 2, 3):
        o = m[idx[k]]
        for j in (1, 2):
            v = rng.normal(size=3)
            v *= rng.uniform(88, 108) / np.linalg.norm(v)
            m[idx[k + j]] = o + np.round(v).astype(np.int64)
    return m


def main():
    s = synth.make_system(**SYSTEM)
    rng = np.random.default_rng(20240611)
    steps, times = [0, 500, 1000], [0.0, 1.0, 2.0]
    water_ints, stats_all = [], []
    blob = b""
    for k in range(3):
        m = water_like(synth.frame_milli(s, k), s.is_bcl, rng)
        fb, stats = encode_frame(m, steps[k], times[k], np.eye(3) * s.box, 1000.0)
        blob += fb
        water_ints.append(m)
        stats_all.append(stats)
    open(os.path.join(HERE, "xtc_water.xtc"), "wb").write(blob)
    wide_ints = []
    blob = b""
    for k in range(2):
        m = rng.integers(-9_000_000, 9_000_000, (60, 3))
        m[11] = m[10] + [3, -2, 5]          # close pairs: runs of one small atom among the separate-field large ones
        m[41] = m[40] + [-40, 17, 1]
        m[42] = m[41] + [2, 2, -60]
        fb, stats = encode_frame(m, 7 + k, 0.5 * k, np.eye(3) * 180000.0, 100.0)
        blob += fb
        wide_ints.append(m)
        stats_all.append(stats)
    open(os.path.join(HERE, "xtc_wide.xtc"), "wb").write(blob)
    prefix = synth._static_prefix(s)
    gro = synth.frame_bytes(s, prefix, 0)
    np.savez_compressed(os.path.join(HERE, "xtc_fixture.npz"), system=np.array([SYSTEM["n_atoms"], SYSTEM["n_bcl"], SYSTEM["seed"]]),
                        water_ints=np.stack(water_ints), water_steps=np.array(steps), water_times=np.array(times),
                        wide_ints=np.stack(wide_ints), wide_steps=np.array([7, 8]), wide_times=np.array([0.0, 0.5]),
                        structure_gro=np.frombuffer(gro, dtype=np.uint8))
    for st in stats_all:
        print(st)


if __name__ == "__main__":
    main()

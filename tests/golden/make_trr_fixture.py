"""Writes tests/golden/trr_{single,double}.trr + trr_fixture.npz: small GROMACS .trr files assembled field by field with the
Python standard library's XDR encoder (xdrlib.Packer), NOT with the repo's own writer (cgromcorrfmo_b200/synth.py) or reader.

Provenance, stated plainly: no GROMACS / MDAnalysis / mdtraj exists in the build container (no network), so these files are
not gmx output.  They follow the on-disk layout of GROMACS' trajectory writer (src/gromacs/fileio/trrio.cpp,
do_trr_frame_header / do_trr_frame_data, unchanged since trnio.c of GROMACS 3):
    int magic = 1993; int slen = 13; string "GMX_trn_file" (XDR: int length 12 + bytes padded to 4);
    int ir_size, e_size, box_size, vir_size, pres_size, top_size, sym_size, x_size, v_size, f_size, natoms, step, nre;
    real t, lambda;   then box[9], vir[9], pres[9] (each only if its size is non-zero), x[3 natoms], v[3 natoms], f[3 natoms]
with `real` = float or double according to the build's precision, everything big-endian XDR.
  trr_single.trr   3 frames, float:  box + x + v + f, steps 5000/10000/15000, lambda 0.25      (what gmx mdrun writes today)
  trr_double.trr   3 frames, double: box + vir + pres + x + v (the layout of GROMACS 3.x/4.0 double-precision builds)
trr_fixture.npz holds the numbers that went in (coordinates as float64, steps, times) and the text of a one-frame .gro of
the same 400-atom system (names and residue numbers for --structure); the topology / itp are regenerated from the seed.

    python tests/golden/make_trr_fixture.py        (needs Python <= 3.12 for xdrlib)
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cgromcorrfmo_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SYSTEM = dict(n_atoms=400, n_bcl=2, seed=77, solvent="shared")


def frame(p, double, natoms, step, t, lam, box, x, v=None, f=None, vir=None, pres=None):
    rb = 8 if double else 4
    real = p.pack_double if double else p.pack_float
    p.pack_int(1993)
    p.pack_int(13)
    p.pack_string(b"GMX_trn_file")
    for size in (0, 0, 9 * rb if box is not None else 0, 9 * rb if vir is not None else 0, 9 * rb if pres is not None else 0, 0, 0,
                 3 * natoms * rb, 3 * natoms * rb if v is not None else 0, 3 * natoms * rb if f is not None else 0):
        p.pack_int(size)
    p.pack_int(natoms)
    p.pack_int(step)
    p.pack_int(0)           # nre
    real(t)
    real(lam)
    for block in (box, vir, pres, x, v, f):
        if block is not None:
            for val in np.asarray(block, dtype=np.float64).reshape(-1):
                real(float(val))


def main():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", DeprecationWarning)
        import xdrlib
    s = synth.make_system(**SYSTEM)
    n = s.n_atoms
    rng = np.random.default_rng(20240607)
    steps = [5000, 10000, 15000]
    times = [st * 0.002 for st in steps]
    # coordinates that are NOT multiples of 0.001 nm: a .trr carries full-precision numbers
    xs = [s.base_xyz + rng.normal(0, 0.013, (n, 3)) for _ in steps]
    box = np.diag([s.box, s.box * 1.01, s.box * 0.99])
    out = {}
    for name, double in (("single", False), ("double", True)):
        p = xdrlib.Packer()
        for k, st in enumerate(steps):
            x = xs[k].astype(np.float32).astype(np.float64) if not double else xs[k]
            v = rng.normal(0, 0.5, (n, 3))
            if double:
                frame(p, True, n, st, times[k], 0.25, box, x, v=v, vir=rng.normal(0, 1e3, (3, 3)), pres=rng.normal(0, 1.0, (3, 3)))
            else:
                frame(p, False, n, st, times[k], 0.25, box, x, v=v, f=rng.normal(0, 100.0, (n, 3)))
        data = p.get_buffer()
        open(os.path.join(HERE, "trr_%s.trr" % name), "wb").write(data)
        out[name + "_bytes"] = len(data)
    gro = synth.frame_bytes(s, synth._static_prefix(s), 0)
    np.savez_compressed(os.path.join(HERE, "trr_fixture.npz"), x=np.stack(xs), steps=np.array(steps), times=np.array(times),
                        structure_gro=np.frombuffer(gro, dtype=np.uint8), system=np.array([SYSTEM["n_atoms"], SYSTEM["n_bcl"], SYSTEM["seed"]]),
                        sizes=np.array([out["single_bytes"], out["double_bytes"]]))
    print("wrote", out)


if __name__ == "__main__":
    main()

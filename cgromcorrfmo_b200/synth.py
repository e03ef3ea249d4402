"""Synthetic FMO-like inputs (.gro trajectory, topology, excited-state itp) for parity tests and benchmarks.

The files obey BOTH readers:

* the reference's whitespace tokeniser (reference src/grom2atomFMO.cpp:396-455): atom serials are written modulo
  10 000 so they never glue onto the atom name, residue/atom names are <= 4 characters, residue names never start with
  a digit, coordinates stay in (-999, 9999);
* a strict fixed-column GROMACS reader ("%5d%-5s%5s%5d%8.3f%8.3f%8.3f" + optional 3 x "%8.4f"), i.e. what the
  device decode kernel consumes.

Topology convention (reference src/grom2atomFMO.cpp:183,440-441,460): every non-chromophore atom is looked up under the
key "<residue>,<atom><ordinal>" where <ordinal> is the running count of non-BCL atoms in .gro order, so the topology is
one "[ atoms ]" block that enumerates those atoms with nr == ordinal.  Chromophore atoms are looked up as
"BCX,<atom><k>", k = 1..140 restarting per BCL residue, and come from the excited-state itp
(reference src/grom2atomFMO.cpp:139-177).

Named shapes (BASELINE.json configs / SURVEY.md section 8d):
  C1 monomer  ~30k atoms,  7 BCL          seed 1001
  C2 trimer  ~100k atoms, 24 BCL, solvent in ONE residue group (numTot ~ 1.1k)   seed 1002
  C4 membrane ~350k atoms, 24 BCL, every solvent molecule its own group          seed 1004
"""
from __future__ import annotations

import dataclasses
import os
from typing import Optional

import numpy as np

N_BCL_ATOMS = 140
N_BCL_DQ = 44  # atoms with a non-zero excited-minus-ground charge difference (as in the reference's bcx.itp)

_RES_NAMES = ["ALA", "ARG", "ASN", "ASP", "CYS", "GLN", "GLU", "GLY", "HIS", "ILE",
              "LEU", "LYS", "MET", "PHE", "PRO", "SER", "THR", "TRP", "TYR", "VAL"]
_RES_ATOMS = ["N", "H", "CA", "HA", "CB", "HB1", "HB2", "CG", "HG1", "HG2", "CD", "HD1", "HD2", "CE", "HE1",
              "HE2", "NZ", "HZ1", "HZ2", "HZ3", "OG", "HG", "C", "O"]


def bcl_atom_names() -> list[str]:
    """140 unique chromophore atom names (<= 4 chars)."""
    names = ["MG"] + [f"N{i}" for i in range(1, 5)] + [f"O{i}" for i in range(1, 7)]
    names += [f"C{i}" for i in range(1, 56)] + [f"H{i}" for i in range(1, 75)]
    assert len(names) == N_BCL_ATOMS and len(set(names)) == N_BCL_ATOMS
    return names


@dataclasses.dataclass
class System:
    """Static description of a synthetic system (everything except per-frame coordinates)."""
    resnr: np.ndarray        # int   [N]  residue number as printed (already modulo 100000)
    resname: list            # str   [N]
    atomname: list           # str   [N]
    is_bcl: np.ndarray       # bool  [N]
    charge: np.ndarray       # float [N]  topology charge text value (non-BCL atoms; 0 for BCL)
    base_xyz: np.ndarray     # float [N,3] nm
    box: float
    bcl_qA: np.ndarray       # [140] ground charges written to the itp
    bcl_qB: np.ndarray       # [140] excited charges (== qA where the line has no B state)
    bcl_has_b: np.ndarray    # bool [140]
    seed: int

    @property
    def n_atoms(self) -> int:
        return int(self.resnr.shape[0])


def _bcl_charges(rng: np.random.Generator):
    qA = np.round(rng.normal(0.0, 0.3, N_BCL_ATOMS), 3)
    qA[0] = 0.7
    has_b = np.zeros(N_BCL_ATOMS, dtype=bool)
    # atoms carrying a B state: 47 lines have one, 44 of them with a different charge (like the reference fixture)
    idx_b = np.sort(rng.choice(N_BCL_ATOMS, size=N_BCL_DQ + 3, replace=False))
    has_b[idx_b] = True
    dq = np.zeros(N_BCL_ATOMS)
    idx_dq = np.sort(rng.choice(idx_b, size=N_BCL_DQ, replace=False))
    # neutral transition: integer milli-charges, all non-zero, summing to zero
    while True:
        milli = np.round(rng.normal(0.0, 30.0, N_BCL_DQ)).astype(int)
        milli[-1] = -int(milli[:-1].sum())
        if np.all(milli != 0) and abs(milli[-1]) < 90:
            break
    dq[idx_dq] = milli / 1000.0
    qB = np.round(qA + dq, 3)
    return qA, qB, has_b


def make_system(n_atoms: int, n_bcl: int, seed: int, solvent: str = "shared",
                protein_fraction: float = 0.165, box: Optional[float] = None,
                first_resnr: int = 8, min_sep: float = 0.12) -> System:
    """Build a static system.

    solvent = "shared"        all solvent atoms carry one residue number -> one dE column (per-residue configs)
            = "per_molecule"  every water has its own residue number     -> one dE column each (per-atom-group configs)
            = "none"          protein + BCL only (the reference fixture's shape)
    """
    rng = np.random.default_rng(seed)
    n_bcl_atoms = n_bcl * N_BCL_ATOMS
    if solvent == "none":
        n_prot = n_atoms - n_bcl_atoms
    else:
        n_prot = int(round(n_atoms * protein_fraction))
    resnr, resname, atomname, is_bcl, charge = [], [], [], [], []
    r = first_resnr
    # --- protein residues
    placed = 0
    while placed < n_prot:
        k = int(rng.integers(7, 25))
        k = min(k, n_prot - placed)
        name = _RES_NAMES[int(rng.integers(len(_RES_NAMES)))]
        q = np.round(rng.normal(0.0, 0.4, k), 4)
        q[-1] = np.round(q[-1] - np.round(q.sum(), 4), 4)  # neutral residue
        for a in range(k):
            resnr.append(r); resname.append(name); atomname.append(_RES_ATOMS[a]); is_bcl.append(False)
            charge.append(float(q[a]))
        placed += k
        r += 1
    # --- chromophores
    names = bcl_atom_names()
    for _ in range(n_bcl):
        for a in range(N_BCL_ATOMS):
            resnr.append(r); resname.append("BCL"); atomname.append(names[a]); is_bcl.append(True); charge.append(0.0)
        r += 1
    # --- solvent
    n_left = n_atoms - len(resnr)
    if solvent != "none":
        n_wat = n_left // 3
        for w in range(n_wat):
            rr = r if solvent == "shared" else r + w
            for nm, q in (("OW", -0.834), ("HW1", 0.417), ("HW2", 0.417)):
                resnr.append(rr); resname.append("SOL"); atomname.append(nm); is_bcl.append(False); charge.append(q)
        r = r + (1 if solvent == "shared" else n_wat)
        # left-over atoms become single-atom ions, each its own residue
        for i in range(n_atoms - len(resnr)):
            resnr.append(r); resname.append("NA"); atomname.append("NA"); is_bcl.append(False); charge.append(1.0)
            r += 1
    n = len(resnr)
    assert n == n_atoms, (n, n_atoms)
    is_bcl = np.asarray(is_bcl)
    if box is None:
        # ~100 atoms / nm^3, but never so small that the chromophore blobs cannot be placed apart
        box = float(max(3.0, (n / 100.0) ** (1.0 / 3.0), 1.8 + 2.4 * max(1, n_bcl) ** (1.0 / 3.0)))
    # --- coordinates: BCL blobs first, then everything else uniform with a minimum separation from BCL atoms
    xyz = np.empty((n, 3))
    centres = []
    tries = 0
    sep = 1.75  # > blob diameter: atoms of different chromophores never come close
    while len(centres) < n_bcl:
        c = rng.uniform(0.9, box - 0.9, 3)
        if all(np.linalg.norm(c - o) >= sep for o in centres):
            centres.append(c)
        tries += 1
        if tries > 200000:
            raise RuntimeError("could not place chromophore centres")
    bidx = np.nonzero(is_bcl)[0]
    for k, c in enumerate(centres):
        # a flat-ish disk of 140 atoms on a jittered grid so that intra-BCL atoms do not coincide
        g = np.stack(np.meshgrid(np.arange(12), np.arange(12), indexing="ij"), -1).reshape(-1, 2)[:N_BCL_ATOMS]
        p = np.zeros((N_BCL_ATOMS, 3))
        p[:, :2] = (g - 5.5) * 0.1
        p[:, 2] = rng.normal(0, 0.05, N_BCL_ATOMS)
        p[:, :2] += rng.normal(0, 0.01, (N_BCL_ATOMS, 2))
        # random rotation
        qv = rng.normal(size=4); qv /= np.linalg.norm(qv)
        a, b, cc, d = qv
        R = np.array([[a*a+b*b-cc*cc-d*d, 2*(b*cc-a*d), 2*(b*d+a*cc)],
                      [2*(b*cc+a*d), a*a-b*b+cc*cc-d*d, 2*(cc*d-a*b)],
                      [2*(b*d-a*cc), 2*(cc*d+a*b), a*a-b*b-cc*cc+d*d]])
        xyz[bidx[k*N_BCL_ATOMS:(k+1)*N_BCL_ATOMS]] = p @ R.T + c
    eidx = np.nonzero(~is_bcl)[0]
    pts = rng.uniform(0.0, box, (eidx.size, 3))
    if n_bcl:
        from scipy.spatial import cKDTree
        tree = cKDTree(xyz[bidx])
        for _ in range(50):
            dmin, _i = tree.query(pts, k=1)
            bad = np.nonzero(dmin < min_sep)[0]
            if bad.size == 0:
                break
            pts[bad] = rng.uniform(0.0, box, (bad.size, 3))
        else:
            raise RuntimeError("could not satisfy the minimum separation")
    xyz[eidx] = pts
    qA, qB, has_b = _bcl_charges(rng)
    return System(resnr=np.asarray(resnr, dtype=np.int64) % 100000, resname=resname, atomname=atomname,
                  is_bcl=is_bcl, charge=np.asarray(charge), base_xyz=xyz, box=box,
                  bcl_qA=qA, bcl_qB=qB, bcl_has_b=has_b, seed=seed)


# ----------------------------------------------------------------------------------------------------------------------
# writers
# ----------------------------------------------------------------------------------------------------------------------
def write_itp(path: str, s: System) -> None:
    """Excited-state itp in the fixture's style: 8 columns, or 11 with a B state, each line ending in ' ;'
    (so that the reference's token count is 9 or 12, reference src/grom2atomFMO.cpp:149)."""
    names = bcl_atom_names()
    with open(path, "w") as f:
        f.write("[ moleculetype ]\n; Name            nrexcl\nBCX                 3\n\n[ atoms ]\n")
        f.write(";   nr       type  resnr residue  atom   cgnr     charge       mass  typeB    chargeB      massB\n")
        for i, nm in enumerate(names):
            mass = {"M": 24.305, "N": 14.01, "O": 16.0, "C": 12.01, "H": 1.008}[nm[0]]
            line = f"{i+1:6d} {('T'+nm):>10s} {1:6d}    BCX {nm:>6s} {i+1:6d} {s.bcl_qA[i]:10.3f} \t{mass:g} "
            if s.bcl_has_b[i]:
                line += f"\t{('T'+nm)} \t{s.bcl_qB[i]:.3f}\t{mass:g} "
            f.write(line + ";\n")
        f.write("\n[ bonds ]\n;  ai    aj funct\n    1     2     1\n")


def write_top(path: str, s: System, with_include: Optional[str] = None) -> None:
    """Topology with one [ atoms ] block enumerating the non-BCL atoms in .gro order (nr == ordinal)."""
    with open(path, "w") as f:
        f.write("; synthetic topology\n[ defaults ]\n; nbfunc comb-rule gen-pairs\n1 2 yes 0.5 0.8333\n\n")
        if with_include:
            f.write(f'#include "{with_include}"\n\n')
        f.write("[ moleculetype ]\n; Name nrexcl\nSystem 3\n\n[ atoms ]\n")
        f.write(";   nr       type  resnr residue  atom   cgnr     charge       mass\n")
        o = 0
        lines = []
        for i in range(s.n_atoms):
            if s.is_bcl[i]:
                continue
            o += 1
            lines.append(f"{o:6d} {'X':>6s} {int(s.resnr[i]):6d} {s.resname[i]:>6s} {s.atomname[i]:>6s} {o:6d} "
                         f"{s.charge[i]:10.4f} {12.01:10.3f}   ; qtot\n")
        f.writelines(lines)
        f.write("\n[ bonds ]\n;  ai    aj funct\n    1     2     1\n\n[ system ]\nsynthetic\n\n[ molecules ]\nSystem 1\n")


def _static_prefix(s: System) -> np.ndarray:
    """uint8 [N,20]: '%5d%-5s%5s%5d' of every atom line (serial written modulo 10000)."""
    n = s.n_atoms
    out = np.empty((n, 20), dtype=np.uint8)
    for i in range(n):
        out[i] = np.frombuffer(
            f"{int(s.resnr[i]):5d}{s.resname[i]:<5s}{s.atomname[i]:>5s}{(i + 1) % 10000:5d}".encode(), dtype=np.uint8)
    return out


def _fmt_fixed(milli: np.ndarray, width: int, ndec: int) -> np.ndarray:
    """Vectorised '%{width}.{ndec}f' of integer values scaled by 10**ndec -> uint8 [..., width]."""
    neg = milli < 0
    a = np.abs(milli).astype(np.int64)
    out = np.full(milli.shape + (width,), ord(" "), dtype=np.uint8)
    pos = width - 1
    for _ in range(ndec):
        out[..., pos] = ord("0") + (a % 10)
        a //= 10
        pos -= 1
    out[..., pos] = ord(".")
    pos -= 1
    out[..., pos] = ord("0") + (a % 10)
    a //= 10
    pos -= 1
    nz = a > 0
    sign_done = np.zeros(milli.shape, dtype=bool)
    while pos >= 0:
        put_sign = neg & ~nz & ~sign_done
        out[..., pos] = np.where(nz, ord("0") + (a % 10), np.where(put_sign, ord("-"), ord(" ")))
        sign_done |= put_sign
        a = a // 10
        nz = a > 0
        pos -= 1
    assert not np.any(nz) and np.all(sign_done | ~neg), "value does not fit the field"
    return out


def frame_milli(s: System, frame: int, jitter: float = 0.01) -> np.ndarray:
    """Integer milli-nm coordinates [N,3] of one frame (values are rounded to the 3 printed decimals)."""
    rng = np.random.default_rng([s.seed, 7919, frame])
    x = s.base_xyz + rng.normal(0.0, jitter, s.base_xyz.shape)
    return np.round(x * 1000.0).astype(np.int64)


def frame_bytes(s: System, prefix: np.ndarray, frame: int, velocities: bool = False, jitter: float = 0.01,
                dt_ps: float = 0.2) -> bytes:
    n = s.n_atoms
    milli = frame_milli(s, frame, jitter)
    width = 44 + (24 if velocities else 0)
    lines = np.empty((n, width + 1), dtype=np.uint8)
    lines[:, :20] = prefix
    lines[:, 20:44] = _fmt_fixed(milli, 8, 3).reshape(n, 24)
    if velocities:
        rng = np.random.default_rng([s.seed, 104729, frame])
        v = np.round(rng.normal(0, 0.5, (n, 3)) * 10000).astype(np.int64)
        lines[:, 44:68] = _fmt_fixed(v, 8, 4).reshape(n, 24)
    lines[:, width] = ord("\n")
    head = f"Generated by synth : FMO-like system t={frame * dt_ps:10.5f}\n{n:5d}\n".encode()
    b = s.box
    tail = f"{b:10.5f}{b:10.5f}{b:10.5f}\n".encode()
    return head + lines.tobytes() + tail


def write_gro(path: str, s: System, n_frames: int, velocities: bool = False, jitter: float = 0.01) -> int:
    """Write the trajectory; returns the (constant) frame size in bytes."""
    prefix = _static_prefix(s)
    size = 0
    with open(path, "wb") as f:
        for t in range(n_frames):
            fb = frame_bytes(s, prefix, t, velocities, jitter)
            size = len(fb)
            f.write(fb)
    return size


def trr_frame_bytes(s: System, frame: int, double: bool = False, velocities: bool = False, forces: bool = False,
                    box: bool = True, jitter: float = 0.01, dt_ps: float = 0.2) -> bytes:
    """One GROMACS .trr frame (XDR, big-endian) holding exactly the numbers frame_bytes() prints as "%8.3f" text:
    single precision stores float32(milli / 1000) — the value (float)atof of the text yields — double stores milli/1000."""
    n = s.n_atoms
    real = ">f8" if double else ">f4"
    rb = 8 if double else 4
    milli = frame_milli(s, frame, jitter)
    x = (milli.astype(np.float64) / 1000.0)
    if not double:
        x = x.astype(np.float32)
    sizes = [0, 0, 9 * rb if box else 0, 0, 0, 0, 0, 3 * n * rb, 3 * n * rb if velocities else 0,
             3 * n * rb if forces else 0, n, frame, 0]
    out = [np.array([1993, 13, 12], ">i4").tobytes(), b"GMX_trn_file", np.array(sizes, ">i4").tobytes(),
           np.array([frame * dt_ps, 0.0], real).tobytes()]
    if box:
        out.append((np.eye(3) * s.box).astype(real).tobytes())
    out.append(x.astype(real).tobytes())
    rng = np.random.default_rng([s.seed, 104729, frame])
    if velocities:
        out.append(rng.normal(0, 0.5, (n, 3)).astype(real).tobytes())
    if forces:
        out.append(rng.normal(0, 100.0, (n, 3)).astype(real).tobytes())
    return b"".join(out)


def write_trr(path: str, s: System, n_frames: int, double: bool = False, velocities: bool = False,
              forces: bool = False, box: bool = True, jitter: float = 0.01) -> int:
    """The .trr twin of write_gro (same frames).  Returns the byte size of frame 0."""
    fb = 0
    with open(path, "wb") as f:
        for t in range(n_frames):
            b = trr_frame_bytes(s, t, double, velocities, forces, box, jitter)
            fb = fb or len(b)
            f.write(b)
    return fb


def xtc_frame_bytes(s: System, frame: int, jitter: float = 0.01, dt_ps: float = 0.2, precision: float = 1000.0) -> bytes:
    """One GROMACS .xtc frame holding exactly the integers frame_bytes() prints as "%8.3f" text (precision 1000: the
    stored integer is the milli-nm value).  Compressed by the library's writer (cgc_xtc_encode, host code)."""
    from . import capi
    milli = frame_milli(s, frame, jitter)
    if precision != 1000.0:
        milli = np.round(milli * (precision / 1000.0)).astype(np.int64)
    return capi.xtc_encode(milli, step=frame, time_ps=frame * dt_ps, box=np.eye(3) * s.box, precision=precision)


def write_xtc(path: str, s: System, n_frames: int, jitter: float = 0.01, precision: float = 1000.0) -> int:
    """The .xtc twin of write_gro (same frames).  Returns the byte size of the largest frame."""
    fb = 0
    with open(path, "wb") as f:
        for t in range(n_frames):
            b = xtc_frame_bytes(s, t, jitter, precision=precision)
            fb = max(fb, len(b))
            f.write(b)
    return fb


def write_inputs(outdir: str, s: System, n_frames: int, velocities: bool = False, jitter: float = 0.01):
    """Write traj.gro / topol.top / bcx.itp into outdir and return their absolute paths."""
    os.makedirs(outdir, exist_ok=True)
    traj = os.path.abspath(os.path.join(outdir, "traj.gro"))
    top = os.path.abspath(os.path.join(outdir, "topol.top"))
    itp = os.path.abspath(os.path.join(outdir, "bcx.itp"))
    write_gro(traj, s, n_frames, velocities, jitter)
    write_top(top, s)
    write_itp(itp, s)
    return traj, top, itp


NAMED = {
    # name: (n_atoms, n_bcl, seed, solvent)
    "tiny": (2000, 7, 1000, "shared"),
    "C1": (30000, 7, 1001, "shared"),
    "C2": (100000, 24, 1002, "shared"),
    "C4": (350000, 24, 1004, "per_molecule"),
}


def named_system(name: str) -> System:
    n_atoms, n_bcl, seed, solvent = NAMED[name]
    return make_system(n_atoms, n_bcl, seed, solvent)

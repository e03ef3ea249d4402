import hashlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def scaled_err(dE, ref64, scale):
    """The parity metric of SURVEY.md section 8(c): |d| / max(|dE_exact|, sum over pairs of |term|) per entry.
    Returns (max scaled error over entries with a non-zero scale, max |value| over entries whose scale is zero)."""
    d = np.abs(np.asarray(dE, dtype=np.float64) - ref64)
    s = np.maximum(np.abs(ref64), scale)
    m = s > 0
    e = float((d[m] / s[m]).max()) if m.any() else 0.0
    z = float(np.abs(np.asarray(dE)[~m]).max()) if (~m).any() else 0.0
    return e, z


def sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


# (tag) -> make_system arguments; small enough for the oracle to finish in seconds
SMALL_SYSTEMS = {
    "tiny": dict(n_atoms=2000, n_bcl=8, seed=42, solvent="shared", frames=3, velocities=False),
    "permol": dict(n_atoms=5000, n_bcl=7, seed=43, solvent="per_molecule", frames=2, velocities=True),
    "nosolv": dict(n_atoms=3000, n_bcl=9, seed=44, solvent="none", frames=2, velocities=False),
    "golden7": dict(n_atoms=1500, n_bcl=7, seed=7, solvent="shared", frames=3, velocities=False),
}


@pytest.fixture(scope="session")
def small_inputs(tmp_path_factory):
    """Synthetic inputs written once per session: tag -> (traj, top, itp, frames)."""
    from cgromcorrfmo_b200 import synth
    cache = {}

    def get(tag):
        if tag not in cache:
            cfg = SMALL_SYSTEMS[tag]
            d = tmp_path_factory.mktemp("syn_" + tag)
            s = synth.make_system(cfg["n_atoms"], cfg["n_bcl"], cfg["seed"], cfg["solvent"])
            traj, top, itp = synth.write_inputs(str(d), s, cfg["frames"], velocities=cfg["velocities"])
            cache[tag] = (traj, top, itp, cfg["frames"])
        return cache[tag]

    return get


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def ref_fixture(oracle_mod):
    """The reference's own test inputs (copied to oracle/_ref/fixtures by `make -C oracle ref`); skip when absent."""
    d = os.path.join(oracle_mod.REF_DIR, "fixtures")
    paths = [os.path.join(d, n) for n in ("test.gro", "4BCL_pp.top", "bcx.itp")]
    if not all(os.path.exists(p) for p in paths):
        pytest.skip("oracle/_ref/fixtures not built (make -C oracle ref, needs /root/reference)")
    return tuple(paths)

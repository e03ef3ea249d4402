"""Generates tests/golden/timecorr.npz by EXECUTING the reference's own time-correlation functions.

Run in the build container, where /root/reference exists:   python tests/golden/make_timecorr_golden.py

scripts/depca/depca/sidechain_corr.py is Python 2 and imports h5py and matplotlib, neither of which exists here, so the
module cannot be imported.  The four functions that matter are taken from its source text instead (TimeCorr :133-139,
HDFStoreCt_v1 :170-182, _v2 :190-218, _v3 :221-242), with the two py2 idioms they contain rewritten mechanically
(`print x` -> `print(x)`, `xrange` -> `range`, `len(mycorr)/10` -> `len(mycorr)//10`, the py2 integer division it was) and run
against a stand-in for the one h5py object they touch (`f[timetag]` -> an ndarray with `.attrs['dt']`).  Everything
numerical in them runs unmodified.
"""
import os
import re
import textwrap

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/scripts/depca/depca/sidechain_corr.py"


def _function_source(text: str, name: str) -> str:
    lines = text.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith("def %s(" % name))
    end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith("def ")), len(lines))
    body = "\n".join(lines[start:end]).expandtabs(4)
    body = re.sub(r"^(\s*)print (.*)$", r"\1print(\2)", body, flags=re.M)
    body = body.replace("xrange(", "range(").replace("len(mycorr)/10", "len(mycorr)//10")
    return textwrap.dedent(body)


class _Dataset(np.ndarray):
    """An ndarray with the `.attrs` mapping the functions read `dt` from."""
    def __new__(cls, a, dt):
        obj = np.asarray(a).view(cls)
        obj.attrs = {"dt": dt}
        return obj


class _File:
    store = {}

    def __init__(self, name, mode="r"):
        self.name = name

    def __enter__(self):
        return _File.store[self.name]

    def __exit__(self, *a):
        return False


class _H5py:
    File = _File


def reference_functions():
    text = open(SRC).read()
    ns = {"np": np, "h5py": _H5py, "print": lambda *a, **k: None}
    for name in ("TimeCorr", "HDFStoreCt_v1", "HDFStoreCt_v2", "HDFStoreCt_v3"):
        exec(compile(_function_source(text, name), SRC + ":" + name, "exec"), ns)
    return ns


def main():
    ns = reference_functions()
    rng = np.random.default_rng(424242)
    T, S, G = 700, 2, 5
    # correlated noise with different means per column, float32 like the .time rows
    white = rng.normal(size=(T + 50, S, G))
    kern = np.exp(-np.arange(50) / 9.0)
    E = np.stack([np.stack([np.convolve(white[:, s, g], kern, "valid")[:T] for g in range(G)], -1) for s in range(S)], 1)
    E = (E * [3.0, 1.0, 0.5, 2.0, 0.1] + [50.0, -20.0, 3.0, 0.0, 7.0]).astype(np.float32)
    _File.store["mem"] = {"tag": _Dataset(E.astype(np.float64), 0.005)}
    cases = []   # (version, chromo (1-based), site_a, site_b, offset, nframes, corr_len)
    out = {}
    for k, (ver, chromo, a, b, offset, nframes, corr_len) in enumerate([
            (1, 1, 0, 3, 0, 700, 0), (1, 2, 4, 4, 37, 600, 0), (2, 1, 1, 2, 0, 700, 128), (2, 2, 0, 0, 100, 500, 500),
            (3, 1, 1, 2, 0, 700, 128), (3, 2, 3, 0, 13, 640, 640), (3, 1, 2, 2, 0, 512, 100)]):
        fn = ns["HDFStoreCt_v%d" % ver]
        if ver == 1:
            c = fn("mem", "tag", None, a, b, chromo=chromo, offset=offset, nframes=nframes)
        else:
            c = fn("mem", "tag", None, a, b, chromo=chromo, offset=offset, nframes=nframes, corr_len=corr_len)
        cases.append((ver, chromo, a, b, offset, nframes, corr_len))
        out["c%d" % k] = np.asarray(c, dtype=np.float64)
    f1 = rng.normal(size=300)
    f2 = rng.normal(size=300)
    out["tc_f1"], out["tc_f2"], out["tc"] = f1, f2, ns["TimeCorr"](f1.copy(), f2.copy())
    np.savez_compressed(os.path.join(HERE, "timecorr.npz"), E=E, cases=np.array(cases, dtype=np.int64), **out)
    print("wrote timecorr.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

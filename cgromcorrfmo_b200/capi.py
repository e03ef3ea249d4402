"""ctypes binding of include/cgromcorr.h (libcgromcorr.so).  No torch types cross this boundary: pointers and sizes only.

The library is CUDA-only.  Loading works without a GPU (so the symbol table can be checked on a CPU box), but every
compute call fails with CgcError(CGC_ERR_CUDA) when no sm_100-class device is present — there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p

import numpy as np

from . import build as _build

CGC_OK = 0
CGC_EOF = 1
CGC_STOPPED = 2
CGC_ERR_INPUT_MALFORMED = -1
CGC_ERR_INPUT_MINOR = -2
CGC_ERR_READ = -3
CGC_ERR_ARG = -4
CGC_ERR_CUDA = -5
CGC_ERR_STATE = -6
CGC_ERR_UNSUPPORTED = -7

_CODE_NAMES = {v: k for k, v in list(globals().items()) if k.startswith("CGC_")}

# every symbol include/cgromcorr.h declares: (restype, argtypes)
SIGNATURES = {
    "cgc_version": (c_char_p, []),
    "cgc_last_error": (c_char_p, [c_void_p]),
    "cgc_create": (c_int, [c_char_p, c_char_p, c_int, POINTER(c_void_p)]),
    "cgc_destroy": (None, [c_void_p]),
    "cgc_table_size": (c_int, [c_void_p]),
    "cgc_table_entry": (c_int, [c_void_p, c_int, c_char_p, c_int, POINTER(c_float), POINTER(c_float)]),
    "cgc_topology_warnings": (c_int, [c_void_p]),
    "cgc_open": (c_int, [c_void_p, c_char_p]),
    "cgc_open_memory": (c_int, [c_void_p, c_void_p, c_int64]),
    "cgc_open_trr": (c_int, [c_void_p, c_char_p, c_char_p]),
    "cgc_open_memory_trr": (c_int, [c_void_p, c_void_p, c_int64, c_char_p]),
    "cgc_open_xtc": (c_int, [c_void_p, c_char_p, c_char_p]),
    "cgc_open_memory_xtc": (c_int, [c_void_p, c_void_p, c_int64, c_char_p]),
    "cgc_xtc_encode": (c_int64, [c_void_p, c_int, c_int, c_float, c_void_p, c_float, c_void_p, c_int64]),
    "cgc_dims": (c_int, [c_void_p, POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "cgc_get_groups": (c_int, [c_void_p, c_void_p]),
    "cgc_get_columns": (c_int, [c_void_p, c_void_p]),
    "cgc_get_site_charges": (c_int, [c_void_p, c_int, c_void_p]),
    "cgc_get_key": (c_int, [c_void_p, c_int, c_char_p, c_int]),
    "cgc_get_layout": (c_int, [c_void_p, POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "cgc_frame_count": (c_int, [c_void_p, POINTER(c_int64)]),
    "cgc_frame_atoms_offset": (c_int64, [c_void_p, c_int64]),
    "cgc_set_option": (c_int, [c_void_p, c_char_p, c_double]),
    "cgc_get_option": (c_double, [c_void_p, c_char_p]),
    "cgc_run": (c_int, [c_void_p, c_int64, c_int64, c_void_p, POINTER(c_int64)]),
    "cgc_request_stop": (None, [c_void_p]),
    "cgc_prepare": (c_int, [c_void_p, c_int]),
    "cgc_comm_create": (c_int, [POINTER(c_int), c_int, POINTER(c_void_p)]),
    "cgc_comm_destroy": (None, [c_void_p]),
    "cgc_cov_allreduce": (c_int, [POINTER(c_void_p), c_int, c_void_p]),
    "cgc_run_write_time": (c_int, [c_void_p, c_int64, c_int64, c_char_p, c_int, c_void_p, POINTER(c_int64)]),
    "cgc_format_text_device": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_int64, POINTER(c_int64), POINTER(c_int)]),
    "cgc_run_device": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_void_p, c_void_p]),
    "cgc_decode_device": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_void_p, c_void_p]),
    "cgc_kernel_times": (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
    "cgc_copy_times": (c_int, [c_void_p, POINTER(c_double), POINTER(c_int64), c_int]),
    "cgc_launch_count": (c_int64, [c_void_p]),
    "cgc_debug_guard_check": (c_int, [POINTER(c_int), POINTER(c_int)]),
    "cgc_decode_errors": (c_int, [c_void_p, POINTER(c_int)]),
    "cgc_cov_size": (c_int64, [c_void_p]),
    "cgc_cov_bind": (c_int, [c_void_p, c_void_p]),
    "cgc_cov_reset": (c_int, [c_void_p]),
    "cgc_cov_partials": (c_int, [c_void_p, c_void_p]),
    "cgc_cov_state_bytes": (c_int64, [c_void_p]),
    "cgc_cov_save_state": (c_int, [c_void_p, c_void_p, c_int64]),
    "cgc_cov_load_state": (c_int, [c_void_p, c_void_p, c_int64]),
    "cgc_cov_set_shift_device": (c_int, [c_void_p, c_void_p, c_void_p]),
    "cgc_cov_accumulate_device": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "cgc_eigh": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, POINTER(c_int)]),
    "cgc_project": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_void_p]),
    "cgc_time_corr": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p,
                      POINTER(c_double)]),
    "cgc_parse_text": (c_int, [c_void_p, c_void_p, c_int64, POINTER(c_int), c_void_p, c_int64, POINTER(c_int64),
                       POINTER(c_double)]),
    "cgc_cov_finalize": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "cgc_write_time": (c_int, [c_char_p, c_int, c_void_p, c_int64, c_int]),
    "cgc_write_mtx": (c_int, [c_void_p, c_char_p, c_int]),
}

_LIB = None


class CgcError(RuntimeError):
    """A non-OK status from the C-ABI.  `.code` is the CGC_* value; the three reference exception classes map to
    CGC_ERR_INPUT_MALFORMED / CGC_ERR_INPUT_MINOR / CGC_ERR_READ (reference src/grom2atomFMO.h:12-27)."""

    def __init__(self, code: int, message: str):
        super().__init__("%s (%d): %s" % (_CODE_NAMES.get(code, "CGC_?"), code, message))
        self.code = code
        self.message = message


def library_path() -> str:
    return _build.LIB


def load(build_if_needed: bool = True):
    """Load libcgromcorr.so (building it first when the sources are newer).  Fails loudly when it cannot be had."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB
    if build_if_needed:
        try:
            _build.build()
        except Exception as e:  # a prebuilt library that is merely "stale" is still the CUDA path
            if not os.path.exists(path):
                raise RuntimeError("libcgromcorr.so is missing and cannot be built (%s); there is no CPU fallback" % e)
    if not os.path.exists(path):
        raise RuntimeError("libcgromcorr.so is missing (run python -m cgromcorrfmo_b200.build); there is no CPU fallback")
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(c_void_p)
    return c_void_p(int(a))  # a raw address (e.g. torch.Tensor.data_ptr())


def xtc_encode(coords, step: int = 0, time_ps: float = 0.0, box=None, precision: float = 1000.0) -> bytes:
    """One GROMACS .xtc frame from integer coordinates [n_atoms, 3] (units of 1 / precision nm); host only, no GPU."""
    lib = load()
    c = np.ascontiguousarray(coords, dtype=np.int32).reshape(-1, 3)
    b = None if box is None else np.ascontiguousarray(box, dtype=np.float32).reshape(9)
    cap = 64 + 16 * c.shape[0]
    while True:
        out = np.empty(cap, np.uint8)
        n = lib.cgc_xtc_encode(_ptr(c), c.shape[0], int(step), float(time_ps), _ptr(b), float(precision), _ptr(out), cap)
        if n < 0:
            raise CgcError(CGC_ERR_UNSUPPORTED, lib.cgc_last_error(None).decode(errors="replace"))
        if n <= cap:
            return out[:n].tobytes()
        cap = int(n)


class Context:
    """Thin object wrapper over cgc_context.  Mirrors the call order of the reference program:
    Context(top, itp) == the two ParseTopology calls; open() == the frame-invariant half of ReadOneTimeGrom2Atom +
    AtomDataLookup_v2; run() == the frame loop of main_Cr."""

    def __init__(self, topology: str, excited_itp: str, device: int = 0):
        self._lib = load()
        h = c_void_p()
        rc = self._lib.cgc_create(os.fsencode(topology), os.fsencode(excited_itp), int(device), ctypes.byref(h))
        if rc != CGC_OK:
            raise CgcError(rc, self._lib.cgc_last_error(None).decode(errors="replace"))
        self._h = h
        self._keep = None

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None):
            self._lib.cgc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc, ok=(CGC_OK,)):
        if rc not in ok:
            raise CgcError(rc, self._lib.cgc_last_error(self._h).decode(errors="replace"))
        return rc

    # -- tables
    def table(self):
        n = self._lib.cgc_table_size(self._h)
        names, masses, charges = [], np.empty(n, np.float32), np.empty(n, np.float32)
        buf = ctypes.create_string_buffer(256)
        m, q = c_float(), c_float()
        for i in range(n):
            self._check(self._lib.cgc_table_entry(self._h, i, buf, 256, ctypes.byref(m), ctypes.byref(q)))
            names.append(buf.value.decode())
            masses[i] = m.value
            charges[i] = q.value
        return names, masses, charges

    @property
    def topology_warnings(self) -> int:
        return self._lib.cgc_topology_warnings(self._h)

    # -- trajectory
    def open(self, traj_path: str):
        self._keep = None
        self._check(self._lib.cgc_open(self._h, os.fsencode(traj_path)))
        return self

    def open_memory(self, text):
        """text: bytes / numpy uint8 array / (address, nbytes).  The buffer is NOT copied and is kept alive here."""
        if isinstance(text, tuple):
            addr, n = text
            self._keep = None
        elif isinstance(text, (bytes, bytearray)):
            arr = np.frombuffer(text, dtype=np.uint8)
            self._keep = (text, arr)
            addr, n = arr.ctypes.data, arr.size
        else:
            arr = np.ascontiguousarray(text).view(np.uint8).reshape(-1)
            self._keep = arr
            addr, n = arr.ctypes.data, arr.size
        self._check(self._lib.cgc_open_memory(self._h, c_void_p(addr), n))
        return self

    def open_trr(self, trr_path: str, structure_gro: str):
        """GROMACS .trr coordinates; names, residues and hence groups/charges from frame 0 of `structure_gro`."""
        self._keep = None
        self._check(self._lib.cgc_open_trr(self._h, os.fsencode(trr_path), os.fsencode(structure_gro)))
        return self

    def open_memory_trr(self, data, structure_gro: str):
        """data: bytes / numpy uint8 array / (address, nbytes) holding .trr frames.  NOT copied; kept alive here."""
        if isinstance(data, tuple):
            addr, n = data
            self._keep = None
        elif isinstance(data, (bytes, bytearray)):
            arr = np.frombuffer(data, dtype=np.uint8)
            self._keep = (data, arr)
            addr, n = arr.ctypes.data, arr.size
        else:
            arr = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
            self._keep = arr
            addr, n = arr.ctypes.data, arr.size
        self._check(self._lib.cgc_open_memory_trr(self._h, c_void_p(addr), n, os.fsencode(structure_gro)))
        return self

    def open_xtc(self, xtc_path: str, structure_gro: str):
        """GROMACS .xtc (compressed) coordinates, decoded on the device; names / residues from frame 0 of `structure_gro`."""
        self._keep = None
        self._check(self._lib.cgc_open_xtc(self._h, os.fsencode(xtc_path), os.fsencode(structure_gro)))
        return self

    def open_memory_xtc(self, data, structure_gro: str):
        """data: bytes / numpy uint8 array / (address, nbytes) holding .xtc frames.  NOT copied; kept alive here."""
        if isinstance(data, tuple):
            addr, n = data
            self._keep = None
        elif isinstance(data, (bytes, bytearray)):
            arr = np.frombuffer(data, dtype=np.uint8)
            self._keep = (data, arr)
            addr, n = arr.ctypes.data, arr.size
        else:
            arr = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
            self._keep = arr
            addr, n = arr.ctypes.data, arr.size
        self._check(self._lib.cgc_open_memory_xtc(self._h, c_void_p(addr), n, os.fsencode(structure_gro)))
        return self

    def dims(self):
        a, c, g, t = c_int(), c_int(), c_int(), c_int()
        self._check(self._lib.cgc_dims(self._h, ctypes.byref(a), ctypes.byref(c), ctypes.byref(g), ctypes.byref(t)))
        return a.value, c.value, g.value, t.value

    @property
    def n_atoms(self):
        return self.dims()[0]

    @property
    def n_chromo(self):
        return self.dims()[1]

    @property
    def n_groups(self):
        return self.dims()[2]

    @property
    def num_tot(self):
        return self.dims()[3]

    @property
    def n_sites(self) -> int:
        return int(self.get_option("sites"))

    def groups(self):
        out = np.empty(self.n_atoms, np.int32)
        self._check(self._lib.cgc_get_groups(self._h, _ptr(out)))
        return out

    def columns(self):
        out = np.empty(self.n_atoms, np.int32)
        self._check(self._lib.cgc_get_columns(self._h, _ptr(out)))
        return out

    def site_charges(self, site: int):
        out = np.empty(self.n_atoms, np.float32)
        self._check(self._lib.cgc_get_site_charges(self._h, int(site), _ptr(out)))
        return out

    def key(self, atom: int) -> str:
        buf = ctypes.create_string_buffer(256)
        self._check(self._lib.cgc_get_key(self._h, int(atom), buf, 256))
        return buf.value.decode()

    def keys(self):
        return [self.key(a) for a in range(self.n_atoms)]

    def layout(self):
        a, b, c, d = c_int(), c_int(), c_int(), c_int()
        self._check(self._lib.cgc_get_layout(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(d)))
        return dict(line_bytes=a.value, x_col=b.value, field_width=c.value, decimals=d.value)

    def frame_count(self) -> int:
        n = c_int64()
        self._check(self._lib.cgc_frame_count(self._h, ctypes.byref(n)))
        return n.value

    def frame_atoms_offset(self, frame: int) -> int:
        return int(self._lib.cgc_frame_atoms_offset(self._h, int(frame)))

    def set_option(self, key: str, value: float):
        self._check(self._lib.cgc_set_option(self._h, key.encode(), float(value)))
        return self

    def get_option(self, key: str) -> float:
        return float(self._lib.cgc_get_option(self._h, key.encode()))

    def prepare(self, with_time_text: bool = False):
        """Allocate the streaming buffers now rather than inside the first run()."""
        self._check(self._lib.cgc_prepare(self._h, 1 if with_time_text else 0))
        return self

    def request_stop(self):
        """Ask the run in progress (or the next one) to stop after the batches under way (the reference's SIGINT flag)."""
        self._lib.cgc_request_stop(self._h)

    # -- hot path
    def run(self, first_frame: int, n_frames: int, want_dE: bool = True, out=None):
        """Returns (dE float32 [frames_done, sites, numTot] or None, frames_done, status)."""
        S, G = self.n_sites, self.num_tot
        if want_dE and out is None:
            out = np.empty((int(n_frames), S, G), np.float32)
        done = c_int64()
        rc = self._lib.cgc_run(self._h, int(first_frame), int(n_frames), _ptr(out) if want_dE else None, ctypes.byref(done))
        self.last_frames_done = done.value          # also meaningful when the call raises: the frames delivered before the error
        self._check(rc, ok=(CGC_OK, CGC_EOF, CGC_STOPPED))
        if want_dE:
            out = out[: done.value]
        return (out if want_dE else None), done.value, rc

    def run_write_time(self, first_frame: int, n_frames: int, time_path: str, truncate: bool = True, want_dE: bool = False):
        """run() that also appends the .time text of the processed frames to `time_path`, formatted on the device.
        Returns (dE or None, frames_done, status)."""
        S, G = self.n_sites, self.num_tot
        dE = np.empty((n_frames, S, G), dtype=np.float32) if want_dE else None
        done = c_int64(0)
        rc = self._lib.cgc_run_write_time(self._h, first_frame, n_frames, os.fsencode(time_path), 1 if truncate else 0,
                                          _ptr(dE), ctypes.byref(done))
        self.last_frames_done = done.value
        self._check(rc, ok=(CGC_OK, CGC_EOF, CGC_STOPPED))
        return (dE[:done.value] if dE is not None else None), done.value, rc

    def format_text_device(self, values, to_cm: bool = False):
        """The device "%g" formatter on a 2-D float32 array: returns (text bytes, needs_host flag)."""
        v = np.ascontiguousarray(values, dtype=np.float32)
        if v.ndim != 2:
            raise ValueError("format_text_device wants a 2-D array")
        cap = v.size * 13 + 16
        out = np.empty(cap, dtype=np.uint8)
        n, flag = c_int64(0), c_int(0)
        self._check(self._lib.cgc_format_text_device(self._h, _ptr(v), v.shape[0], v.shape[1], 1 if to_cm else 0, _ptr(out),
                                                     cap, ctypes.byref(n), ctypes.byref(flag)))
        return out[:n.value].tobytes(), flag.value

    def run_device(self, d_text, text_bytes: int, atoms_offset, n_frames: int, d_dE, stream=0):
        off = np.ascontiguousarray(atoms_offset, dtype=np.int64)
        self._check(self._lib.cgc_run_device(self._h, _ptr(d_text), int(text_bytes), _ptr(off), int(n_frames), _ptr(d_dE),
                                             c_void_p(int(stream))))

    def decode_device(self, d_text, text_bytes: int, atoms_offset, n_frames: int, d_xyz, stream=0):
        off = np.ascontiguousarray(atoms_offset, dtype=np.int64)
        self._check(self._lib.cgc_decode_device(self._h, _ptr(d_text), int(text_bytes), _ptr(off), int(n_frames),
                                                _ptr(d_xyz), c_void_p(int(stream))))

    def copy_times(self, reset: bool = False):
        """(ms, copies): device time of the H2D copies of the trajectory text since "profile" was switched on."""
        ms, n = c_double(0), c_int64(0)
        self._check(self._lib.cgc_copy_times(self._h, ctypes.byref(ms), ctypes.byref(n), 1 if reset else 0))
        return ms.value, n.value

    def kernel_times(self, reset: bool = False):
        """{kernel: (ms, launches)} accumulated since option "profile" was switched on (CUDA events on the launch stream)."""
        ms = np.zeros(4, np.float64)
        n = np.zeros(4, np.int64)
        self._check(self._lib.cgc_kernel_times(self._h, _ptr(ms), _ptr(n), 1 if reset else 0))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(("decode", "cdc", "finish", "cov"))}

    @property
    def launch_count(self) -> int:
        return int(self._lib.cgc_launch_count(self._h))

    def decode_errors(self) -> int:
        f = c_int()
        self._check(self._lib.cgc_decode_errors(self._h, ctypes.byref(f)))
        return f.value

    # -- Correlate
    def cov_size(self) -> int:
        return int(self._lib.cgc_cov_size(self._h))

    def cov_bind(self, d_buffer):
        self._check(self._lib.cgc_cov_bind(self._h, _ptr(d_buffer)))

    def cov_reset(self):
        self._check(self._lib.cgc_cov_reset(self._h))

    def cov_partials(self):
        out = np.empty(self.cov_size(), np.float64)
        self._check(self._lib.cgc_cov_partials(self._h, _ptr(out)))
        return out

    def cov_save_state(self) -> bytes:
        """Checkpoint of the Correlate sums (cgc_cov_save_state)."""
        n = int(self._lib.cgc_cov_state_bytes(self._h))
        buf = np.empty(n, dtype=np.uint8)
        self._check(self._lib.cgc_cov_save_state(self._h, _ptr(buf), n))
        return buf.tobytes()

    def cov_load_state(self, blob: bytes):
        buf = np.frombuffer(blob, dtype=np.uint8)
        self._check(self._lib.cgc_cov_load_state(self._h, _ptr(buf), buf.size))

    def cov_set_shift_device(self, d_row, stream=0):
        self._check(self._lib.cgc_cov_set_shift_device(self._h, _ptr(d_row), c_void_p(int(stream))))

    def cov_accumulate_device(self, d_dE, n_frames: int, stream=0):
        self._check(self._lib.cgc_cov_accumulate_device(self._h, _ptr(d_dE), int(n_frames), c_void_p(int(stream))))

    def cov_finalize(self, ddof: int = 0):
        S, G = self.n_sites, self.num_tot
        mean = np.empty((S, G), np.float64)
        cov = np.empty((S, G, G), np.float64)
        self._check(self._lib.cgc_cov_finalize(self._h, int(ddof), _ptr(mean), _ptr(cov)))
        return mean, cov

    def eigh(self, matrices):
        """Batched symmetric eigen-decomposition on the device.  matrices (M, n, n) float64 -> (evals (M, n) ascending,
        evecs (M, n, n) with eigenvectors as ROWS, sweeps)."""
        a = np.ascontiguousarray(matrices, dtype=np.float64)
        if a.ndim != 3 or a.shape[1] != a.shape[2]:
            raise ValueError("eigh wants (M, n, n)")
        ev = np.empty(a.shape[:2], dtype=np.float64)
        vt = np.empty(a.shape, dtype=np.float64)
        sw = c_int(0)
        self._check(self._lib.cgc_eigh(self._h, _ptr(a), a.shape[0], a.shape[1], _ptr(ev), _ptr(vt), ctypes.byref(sw)))
        return ev, vt, sw.value

    def project(self, dE_rows, mean, modes):
        """out[t, s, m] = sum_g (dE_rows[t, s, g] - mean[s, g]) * modes[s, m, g] on the device (FP64 tensor cores).
        dE_rows float32 (T, S, G); mean (S, G); modes (S, M, G), eigenvectors as rows.  Returns float64 (T, S, M)."""
        x = np.ascontiguousarray(dE_rows, dtype=np.float32)
        mu = np.ascontiguousarray(mean, dtype=np.float64)
        w = np.ascontiguousarray(modes, dtype=np.float64)
        S, G = self.n_sites, self.num_tot
        if x.ndim != 3 or x.shape[1:] != (S, G) or mu.shape != (S, G) or w.ndim != 3 or w.shape[0] != S or w.shape[2] != G:
            raise ValueError("project: expected rows (T, %d, %d), mean (%d, %d), modes (%d, M, %d)" % (S, G, S, G, S, G))
        out = np.empty((x.shape[0], S, w.shape[1]), dtype=np.float64)
        self._check(self._lib.cgc_project(self._h, _ptr(x), x.shape[0], _ptr(mu), _ptr(w), w.shape[1], _ptr(out)))
        return out

    def time_corr(self, series, pair_first, pair_second, corr_len: int, circular: bool = False, want_ms: bool = False):
        """out[p, tau] = sum_t series[first[p], t] * series[second[p], t + tau] / (n - tau) for tau < corr_len, on the device
        in FP64 (K7).  series float64 (n_series, n), mean-free; circular=True takes t + tau modulo n.  Returns float64
        (n_pairs, corr_len) (and the device time of the kernels in ms when want_ms)."""
        s = np.ascontiguousarray(series, dtype=np.float64)
        if s.ndim != 2:
            raise ValueError("time_corr wants series (n_series, n)")
        a = np.ascontiguousarray(pair_first, dtype=np.int32).ravel()
        b = np.ascontiguousarray(pair_second, dtype=np.int32).ravel()
        if a.shape != b.shape:
            raise ValueError("time_corr: pair_first and pair_second differ in length")
        out = np.empty((a.size, int(corr_len)), dtype=np.float64)
        ms = c_double(0.0)
        self._check(self._lib.cgc_time_corr(self._h, _ptr(s), s.shape[0], s.shape[1], _ptr(a), _ptr(b), a.size, int(corr_len),
                                            1 if circular else 0, _ptr(out), ctypes.byref(ms)))
        return (out, ms.value) if want_ms else out

    def parse_text(self, text, num_col: int = 0, want_ms: bool = False):
        """`.time` / `.mtx` text -> float64 (rows, num_col), every value bit-identical to float(token) (K8).  text: bytes, or a
        uint8 numpy array / memmap of the file; num_col 0 = the width of the first row."""
        if isinstance(text, (bytes, bytearray, memoryview)):
            buf = np.frombuffer(text, dtype=np.uint8)
        else:
            buf = np.ascontiguousarray(text).view(np.uint8).ravel()
        ncol = c_int(int(num_col))
        rows = c_int64(0)
        ms = c_double(0.0)
        if buf.size == 0:
            return (np.empty((0, 0)), 0.0) if want_ms else np.empty((0, 0))
        self._check(self._lib.cgc_parse_text(self._h, _ptr(buf), buf.size, ctypes.byref(ncol), None, 0, ctypes.byref(rows), None))
        out = np.empty((rows.value, ncol.value), dtype=np.float64)
        if rows.value:
            self._check(self._lib.cgc_parse_text(self._h, _ptr(buf), buf.size, ctypes.byref(ncol), _ptr(out), out.size,
                                                 ctypes.byref(rows), ctypes.byref(ms)))
        return (out, ms.value) if want_ms else out

    def write_mtx(self, path: str, compat: bool = False):
        self._check(self._lib.cgc_write_mtx(self._h, os.fsencode(path), 1 if compat else 0))


def debug_guard_check():
    """(buffers guarded, buffers with a damaged guard band) — needs CGC_GUARD=1 in the environment before the library
    allocates anything; raises otherwise."""
    lib = load()
    n, bad = c_int(0), c_int(0)
    rc = lib.cgc_debug_guard_check(ctypes.byref(n), ctypes.byref(bad))
    if rc != CGC_OK:
        raise CgcError(rc, "guard bands are off (set CGC_GUARD=1 before the first context is created)")
    return n.value, bad.value


def cov_allreduce(contexts):
    """The one collective for several contexts of THIS process (one per GPU): ncclCommInitAll + one grouped
    ncclAllReduce(sum, double) over n, sum(x-c), sum (x-c)(x-c)^T (cgc_cov_allreduce).  Afterwards every context holds the
    sums of all frame ranges."""
    lib = load()
    arr = (c_void_p * len(contexts))(*[c._h for c in contexts])
    rc = lib.cgc_cov_allreduce(arr, len(contexts), None)
    if rc != CGC_OK:
        contexts[0]._check(rc)


def write_time(path: str, dE_kcal, truncate: bool = True):
    """Append dE rows to a .time file exactly as the reference prints them (reference src/FMOparse.cpp:151-165)."""
    lib = load()
    a = np.ascontiguousarray(dE_kcal, dtype=np.float32)
    rows = int(np.prod(a.shape[:-1])) if a.ndim > 1 else 1
    rc = lib.cgc_write_time(os.fsencode(path), 1 if truncate else 0, _ptr(a), rows, int(a.shape[-1]))
    if rc != CGC_OK:
        raise CgcError(rc, "cannot write " + path)

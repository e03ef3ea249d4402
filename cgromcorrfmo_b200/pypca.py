"""ctypes shim for the pyPCA / depca post-processing scripts (reference scripts/depca, scripts/2014-06-csv2hdf5.py).

Those scripts only ever consume `csv_out.time`: the text is parsed back into a
(frames, 7, 357) array (scripts/2014-06-csv2hdf5.py:56-74), split per site (scripts/depca/depca/dedata.py:96-120), and
reduced with np.cov + mean per site (scripts/depca/depca/sidechain_corr.py:50-72) before eigh and projection.  This module
hands them the same arrays straight from the GPU path, with depca's names, shapes, units and return order, so the text
round trip disappears.  Everything numerical that is on the hot path (dE rows, mean, covariance) comes from the C-ABI;
what follows it (eigen-decomposition, projection) is plain numpy, as in depca.

    E_t_ia               = load_E_t_ia(traj, top, itp, num_frames)            # float32 (T, Nsites, Ncoarse), cm^-1
    Corr_i_ab, AvgEia    = AvgAndCorrelateSidechains_gpu(traj, top, itp, T)   # depca's return order
    vi, wi, impact_i     = ComputeModes(Corr_i_ab)                            # or ComputeModes_gpu(ctx, Corr_i_ab)
    proj_t_im            = ApplyPCA_gpu(ctx, E_t_ia, AvgEia, wi)              # (E - mean) . modes on FP64 tensor cores
    C_t                  = HDFStoreCt_v3(ctx, E_t_ia, site_a, site_b, chromo=1, nframes=T, corr_len=L)   # time correlation (K7)
"""
from __future__ import annotations

import numpy as np

from . import capi

CM_PER_KCAL = 349.7   # reference src/FMOparse.cpp:151


def _run(traj, top, itp, num_frames, sites, device, want_rows):
    with capi.Context(top, itp, device) as c:
        c.open(traj)
        if sites:
            c.set_option("sites", sites)
        dE, done, _ = c.run(0, num_frames, want_dE=want_rows)
        mean = cov = None
        if c.cov_size() and done > 1:
            mean, cov = c.cov_finalize(ddof=1)     # np.cov's default, as depca uses it (sidechain_corr.py:67)
        return dE, done, mean, cov


def load_E_t_ia(traj, top, itp, num_frames, sites=7, device=0, units="cm-1"):
    """The array depca calls E_t_ia: (frames, Nsites, Ncoarse).  `.time` holds cm^-1 (x349.7 of the kcal/mol rows), and so
    does this by default; the conversion is the reference's own: float(double(kcal) * 349.7) (src/FMOparse.cpp:151-154).
    sites=7 is the stock program's site count (src/FMOparse.cpp:133); 0 = every chromophore."""
    dE, done, _, _ = _run(traj, top, itp, num_frames, sites, device, True)
    if units == "kcal":
        return dE
    return (dE.astype(np.float64) * CM_PER_KCAL).astype(np.float32)


def load_E_t_ia_npy(time_npy_path, units="cm-1"):
    """The array the CLI wrote with --npy (csv_out.time.npy: float32 (frames, Nsites, Ncoarse) in kcal/mol) as depca's
    E_t_ia — the text round trip of scripts/2014-06-csv2hdf5.py:56-74 without the text.  Memory-mapped when units="kcal"."""
    E = np.load(time_npy_path, mmap_mode="r")
    if units == "kcal":
        return E
    return (np.asarray(E, dtype=np.float64) * CM_PER_KCAL).astype(np.float32)


def load_E_t_ia_csv(ctx, csv_file, frames_per_file=None, Nsites=7, dtype=np.float32):
    """A `.time` text file as depca's E_t_ia, for files that exist only as text (earlier runs of the reference): what
    scripts/2014-06-csv2hdf5.py:56-74 does — every token through float(), reshape to (frames_per_file, 7, 357), stored in an
    HDF5 dataset of h5py's default dtype float32 — with the text parsed on the device (K8, C-ABI cgc_parse_text; values
    bit-identical to float(token)).  frames_per_file=None takes what the file holds; a number must be right, as in the
    script ("must be correct for script to function").  dtype=np.float64 keeps what float() returned."""
    text = np.memmap(csv_file, dtype=np.uint8, mode="r") if not isinstance(csv_file, (bytes, bytearray)) else csv_file
    rows = ctx.parse_text(text)
    if rows.shape[0] % Nsites:
        raise ValueError("%d rows are not a whole number of frames of %d sites" % (rows.shape[0], Nsites))
    frames = rows.shape[0] // Nsites
    if frames_per_file is not None and frames_per_file != frames:
        raise ValueError("cannot reshape array of size %d into shape (%d,%d,%d)" % (rows.size, frames_per_file, Nsites, rows.shape[1]))
    return rows.reshape(frames, Nsites, rows.shape[1]).astype(dtype, copy=False)


def AvgAndCorrelateSidechains_gpu(traj, top, itp, num_frames, sites=7, device=0, units="cm-1"):
    """depca.sidechain_corr.AvgAndCorrelateSidechains (sidechain_corr.py:36-72) without materialising the time series on
    the host: returns (Corr_i_ab [Nsites, Ncoarse, Ncoarse], AvgEia [Nsites, Ncoarse]) — np.cov(rowvar=0) (ddof 1) and the
    mean per site, computed by the Correlate kernels on the float32 rows."""
    _, done, mean, cov = _run(traj, top, itp, num_frames, sites, device, False)
    if cov is None:
        raise ValueError("covariance is unavailable (fewer than two frames, or numTot too large for the partial sums)")
    if units == "kcal":
        return cov, mean
    # the rows are scaled by 349.7 before np.cov in depca's pipeline (they come from the .time text)
    return cov * (CM_PER_KCAL * CM_PER_KCAL), mean * CM_PER_KCAL


def AvgAndCorrelateSidechains(E_t_ia):
    """Same reduction for an array that is already on the host (numpy), per site: np.cov(X, rowvar=0) and the mean
    (sidechain_corr.py:50-72).  Post-processing, not the hot path."""
    E = np.asarray(E_t_ia, dtype=np.float64)
    corr = np.stack([np.cov(E[:, i, :], rowvar=False) for i in range(E.shape[1])])
    return corr, E.mean(axis=0)


def ComputeModes(corr, sort=True):
    """depca.sidechain_corr.ComputeModes (sidechain_corr.py:74-129): eigh per site, eigenvectors as rows, sign chosen so that
    the component sum has the sign of the eigenvalue, modes sorted by eigenvalue x impact^2 (ascending, like the script)."""
    corr = np.asarray(corr)
    vi, wi, imp = [], [], []
    for c in corr:
        v, w = np.linalg.eigh(c)
        w = w.T.copy()
        n = w.sum(axis=1)
        flip = ((v < 0) & (n > 0)) | ((v > 0) & (n < 0))
        w[flip] *= -1
        n[flip] *= -1
        if sort:
            order = np.argsort(v * n * n, kind="stable")
            v, w, n = v[order], w[order], n[order]
        vi.append(v); wi.append(w); imp.append(n)
    return np.array(vi), np.array(wi), np.array(imp)


def ComputeModes_gpu(ctx, corr, sort=True):
    """ComputeModes with the eigen-decomposition on the device (K6, batched block Jacobi; C-ABI cgc_eigh) instead of
    numpy.linalg.eigh; the sign convention and the sort are depca's, as above.  24 sites of 1099 groups: 0.55 s against
    1.9 s for numpy on 16 cores."""
    v_all, w_all, _ = ctx.eigh(np.asarray(corr, dtype=np.float64))
    vi, wi, imp = [], [], []
    for v, w in zip(v_all, w_all):
        w = w.copy()
        n = w.sum(axis=1)
        flip = ((v < 0) & (n > 0)) | ((v > 0) & (n < 0))
        w[flip] *= -1
        n[flip] *= -1
        if sort:
            order = np.argsort(v * n * n, kind="stable")
            v, w, n = v[order], w[order], n[order]
        vi.append(v); wi.append(w); imp.append(n)
    return np.array(vi), np.array(wi), np.array(imp)


def ApplyPCA(E_t_ia, AvgEia, wi):
    """depca.convert.ApplyPCA_hdf5 (convert.py:84-125) in memory: (E - mean) projected on the modes, per site."""
    E = np.asarray(E_t_ia, dtype=np.float64)
    return np.stack([np.inner(E[:, i, :] - AvgEia[i], wi[i]) for i in range(E.shape[1])], axis=1)


def ApplyPCA_gpu(ctx, E_t_ia, AvgEia, wi):
    """ApplyPCA on the device: `ctx` is an open capi.Context of the same system (it fixes Nsites and Ncoarse); the rows,
    the means and the modes are host arrays as above.  Returns float64 (T, Nsites, Nmodes) == ApplyPCA(E_t_ia, AvgEia, wi)
    up to FP64 summation order (K4, DMMA; C-ABI cgc_project)."""
    return ctx.project(E_t_ia, AvgEia, wi)


# ---------------------------------------------------------------------------------------------------------------------
# time-correlation functions (sidechain_corr.py:133-242) on the device: K7, C-ABI cgc_time_corr.  The reference reads the
# (T, Nsites, Ncoarse) dataset from HDF5; here it is the array itself (load_E_t_ia / load_E_t_ia_npy), everything else —
# argument names, 1-based `chromo`, the mean over the WHOLE series, the window [offset, offset + nframes), the error
# texts — is the scripts'.
# ---------------------------------------------------------------------------------------------------------------------
def TimeCorr_gpu(ctx, f1, f2):
    """depca TimeCorr (sidechain_corr.py:133-139): np.correlate(f1, f2, 'full')[n-1:] / (n - tau) for every lag, i.e.
    sum_t f2[t] f1[t + tau] / (n - tau); None when both series are identically zero, as there."""
    f1 = np.asarray(f1, dtype=np.float64).ravel()
    f2 = np.asarray(f2, dtype=np.float64).ravel()
    assert f1.size == f2.size
    if not (np.any(f1 != 0.0) or np.any(f2 != 0.0)):
        return None
    return ctx.time_corr(np.stack([f1, f2]), [1], [0], f1.size)[0]


def _ct_series(E_tij, columns, chromo, offset, nframes, check):
    start, stop = offset, offset + nframes
    if check:
        if start >= E_tij.shape[0]:
            raise RuntimeError("offset > len(E_tij) [{} > {}]; Use a smaller offset.".format(start, E_tij.shape[0]))
        if stop > E_tij.shape[0]:
            raise RuntimeError("offset + nframes > len(E_tij) [{} > {}]; Use a smaller offset or fewer frames.".format(
                stop, E_tij.shape[0]))
    cols = np.asarray(columns, dtype=np.int64)
    X = np.asarray(E_tij[:, chromo - 1, :][:, cols], dtype=np.float64)          # (T, n_cols)
    return np.ascontiguousarray((X[start:stop] - X.mean(axis=0)).T)               # mean over the whole series (:179-180)


def HDFStoreCt_v1(ctx, E_tij, site_a, site_b, chromo=1, offset=0, nframes=200000):
    """sidechain_corr.py:170-182: TimeCorr of the two columns, first tenth of the lags (no bounds checks there)."""
    s = _ct_series(E_tij, [site_a, site_b], chromo, offset, nframes, check=False)
    n = s.shape[1]
    if not np.any(s != 0.0):
        raise TypeError("'NoneType' object is not subscriptable")   # TimeCorr returned None in the script
    return ctx.time_corr(s, [1], [0], n)[0][0:n // 10]


def HDFStoreCt_v2(ctx, E_tij, site_a, site_b, chromo=1, offset=0, nframes=200000, corr_len=20000):
    """sidechain_corr.py:190-218: mycorr[tc] = inner(f1[0:n-tc], f2[tc:n]) / (n - tc)."""
    s = _ct_series(E_tij, [site_a, site_b], chromo, offset, nframes, check=True)
    return ctx.time_corr(s, [0], [1], corr_len)[0]


def HDFStoreCt_v3(ctx, E_tij, site_a, site_b, chromo=1, offset=0, nframes=200000, corr_len=20000):
    """sidechain_corr.py:221-242: the FFT form, i.e. the circular sum over the window, / (n - tau)."""
    s = _ct_series(E_tij, [site_a, site_b], chromo, offset, nframes, check=True)
    return ctx.time_corr(s, [0], [1], corr_len, circular=True)[0]


def StoreCt_pairs_gpu(ctx, E_tij, pairs, chromo=1, offset=0, nframes=200000, corr_len=20000, circular=True):
    """Many (site_a, site_b) pairs of one chromophore in one call (the scripts loop over pairs in separate processes):
    returns float64 (len(pairs), corr_len); circular=True is HDFStoreCt_v3, False HDFStoreCt_v2."""
    pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    cols, inv = np.unique(pairs, return_inverse=True)
    inv = inv.reshape(-1, 2)
    s = _ct_series(E_tij, cols, chromo, offset, nframes, check=True)
    return ctx.time_corr(s, inv[:, 0], inv[:, 1], corr_len, circular=circular)

/* cgromcorr.h — C-ABI of the B200-native traj2dEcsv hot path.
 *
 * One shared library (cgromcorrfmo_b200/libcgromcorr.so), plain C types only, never throws across the boundary.
 * Two clients bind it: the C++ CLI driver `traj2dEcsv` (drop-in for the reference program) and the ctypes layer
 * cgromcorrfmo_b200/capi.py used by the pyPCA shim, the tests and bench.py.
 *
 * The reference (jrhaberstroh/cGromCorrFMO) has no FFI layer; the interface this replaces is the C++ library
 * src/grom2atomFMO.h plus the frame loop of src/FMOparse.cpp.  Every entry point cites what it stands in for.
 * There is NO CPU fallback: every compute entry point needs a CUDA device (sm_100a) and fails with CGC_ERR_CUDA
 * otherwise.
 *
 * Units: coordinates nm, charges e, dE in kcal/mol (the reference's cdc_kcal).  `.time` text is in cm^-1
 * (x 349.7, reference src/FMOparse.cpp:151).
 * Layout of every dE block: float32 [frames][sites][numTot], C order, column = chromophore index 0..nChromo-1 first,
 * then environment groups (reference src/grom2atomFMO.cpp:688-693).
 */
#ifndef CGROMCORR_H
#define CGROMCORR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cgc_context cgc_context;
typedef struct cgc_comm cgc_comm;   /* an NCCL communicator set over the GPUs one process drives (cgc_comm_create) */

/* ---- status codes.  The three reference exception classes (src/grom2atomFMO.h:12-27) map to the first three. ---- */
#define CGC_OK                    0
#define CGC_EOF                   1   /* fewer frames in the trajectory than requested; *frames_done tells how many.
                                         (The reference segfaults here, README.md:31.) */
#define CGC_STOPPED               2   /* cgc_request_stop ended the call early; *frames_done tells how many frames were
                                         delivered (the reference's SIGINT flag, src/FMOparse.cpp:24-29,62,68) */
#define CGC_ERR_INPUT_MALFORMED  (-1) /* InputFileMalformed */
#define CGC_ERR_INPUT_MINOR      (-2) /* InputFileMinorMalformed — only ever reported, never fatal (main() swallows it,
                                         src/FMOparse.cpp:358-366) */
#define CGC_ERR_READ             (-3) /* ReadError */
#define CGC_ERR_ARG              (-4)
#define CGC_ERR_CUDA             (-5)
#define CGC_ERR_STATE            (-6) /* call order violated (e.g. run before open) */
#define CGC_ERR_UNSUPPORTED      (-7) /* input is legal for the reference's tokeniser but not fixed-column */

const char *cgc_version(void);

/* Message of the last failure on this context; ctx == NULL returns the message of the last failed cgc_create on the
 * calling thread. */
const char *cgc_last_error(const cgc_context *ctx);

/* ---- tables -----------------------------------------------------------------------------------------------------
 * cgc_create replaces the two ParseTopology calls of main() (src/FMOparse.cpp:355-366; ParseTopology itself
 * src/grom2atomFMO.cpp:78-234): excited_itp is parsed in Excited::YES mode first, then topology in Excited::NO mode
 * (following #include).  Keys are "<residue>,<atom><nr>" and "<residue>,<atom><nr>,EXC"; first occurrence wins.
 * device < 0 builds the tables only (no CUDA call is made; useful for inspecting the parse). */
int cgc_create(const char *topology_path, const char *excited_itp_path, int device, cgc_context **out);
void cgc_destroy(cgc_context *ctx);

int cgc_table_size(const cgc_context *ctx);
/* i-th table entry in insertion order == the reference's (mergeNameTable, massTable, chargeTable)[i]. */
int cgc_table_entry(const cgc_context *ctx, int i, char *name, int name_cap, float *mass, float *charge);
/* number of "conflicts with prior entry" notes (what the reference packs into InputFileMinorMalformed). */
int cgc_topology_warnings(const cgc_context *ctx);

/* ---- trajectory -------------------------------------------------------------------------------------------------
 * cgc_open replaces everything ReadOneTimeGrom2Atom (src/grom2atomFMO.cpp:338-498) and AtomDataLookup_v2 (:549-610)
 * derive that does NOT change from frame to frame: group ids, lookup keys, per-atom charges, per-site dQ lists.  It
 * is done once on frame 0 with the reference's whitespace-token rules; the coordinates of every frame are then decoded
 * on the device from fixed columns (layout checked against the token values on frame 0, CGC_ERR_UNSUPPORTED if they
 * disagree).  cgc_open_memory is the same for trajectory text that already sits in host memory (not copied: it must
 * stay valid until the context is destroyed or re-opened). */
int cgc_open(cgc_context *ctx, const char *traj_path);
int cgc_open_memory(cgc_context *ctx, const char *text, int64_t nbytes);
/* GROMACS .trr trajectories (binary, big-endian, single or double precision) read directly.  Not a reference format:
 * the authors run trjconv to turn .trr into .gro text before traj2dEcsv (scripts/2014-09-trr2dEcsv.sh:22); this entry
 * point removes that step and the 45-69 bytes of text per atom and frame with it.  A .trr holds
 * no names, so residue numbers, residue and atom names — hence groups, keys and charges — come from the first frame of
 * `structure_gro_path` (any .gro of the same system, e.g. the frame trjconv would have written first); its coordinates
 * are not used.  Frames may carry box, velocities and forces: only the coordinate array is decoded.  Coordinates are
 * used as stored (double precision is rounded once to float32): a .trr written from the same numbers as a "%8.3f" .gro
 * gives bit-identical dE.  CGC_ERR_INPUT_MALFORMED for a bad magic number / version string / size field or an atom
 * count different from the structure file's, CGC_ERR_UNSUPPORTED for frames without coordinates. */
int cgc_open_trr(cgc_context *ctx, const char *trr_path, const char *structure_gro_path);
int cgc_open_memory_trr(cgc_context *ctx, const char *data, int64_t nbytes, const char *structure_gro_path);
/* GROMACS .xtc trajectories (compressed integer coordinates, the format production runs are usually kept in) read
 * directly and DECODED ON THE DEVICE: 3-6 bytes per atom and frame cross the PCIe link instead of the 45-69 bytes of .gro
 * text or the 12 of a .trr ("a binary trr/xtc reader").  Not a reference format either (same trjconv
 * step, scripts/2014-09-trr2dEcsv.sh:22).  Names and residues come from `structure_gro_path` as for cgc_open_trr.  A stored
 * integer k becomes the float32 nearest to k / precision — the value (float)atof (src/grom2atomFMO.cpp:448-450) yields on
 * the "%8.3f" text trjconv would have written — so an .xtc of precision 1000 and the .gro of the same frames give
 * bit-identical dE.  The host only walks the frame headers; CGC_ERR_INPUT_MALFORMED for a bad magic number, impossible
 * header fields, an atom count different from the structure file's, or (from cgc_run) a bit stream whose runs do not add
 * up; CGC_ERR_UNSUPPORTED for frames of at most 9 atoms (stored uncompressed).  For cgc_run_device / cgc_decode_device the
 * offsets are those of cgc_frame_atoms_offset: the frame's coordinate block (a multiple of 4 inside a 16-byte aligned
 * buffer), and text_bytes must cover the last frame's stream. */
int cgc_open_xtc(cgc_context *ctx, const char *xtc_path, const char *structure_gro_path);
int cgc_open_memory_xtc(cgc_context *ctx, const char *data, int64_t nbytes, const char *structure_gro_path);
/* Write one .xtc frame (host; no device needed): coords = int32[3 * n_atoms] in units of 1 / precision nm, box9 = the nine
 * box floats (NULL: zeros), compressed the way GROMACS' writer chooses runs and table indices (n_atoms > 9).  Returns the
 * frame's size in bytes — the frame is stored in `out` when it fits `cap` (call with cap 0 to size the buffer) — or -1
 * (cgc_last_error(NULL)).  What trjconv -o traj.xtc does for files; used here to produce test and benchmark inputs. */
int64_t cgc_xtc_encode(const int32_t *coords, int n_atoms, int step, float time_ps, const float *box9, float precision,
                       char *out, int64_t cap);

int cgc_dims(const cgc_context *ctx, int *n_atoms, int *n_chromo, int *n_groups, int *num_tot);
/* atomGroupi of ReadOneTimeGrom2Atom: -k for the k-th chromophore, +g for the g-th other residue. int32[n_atoms] */
int cgc_get_groups(const cgc_context *ctx, int32_t *out);
/* dE column of every atom. int32[n_atoms] */
int cgc_get_columns(const cgc_context *ctx, int32_t *out);
/* atomChargei_e of AtomDataLookup_v2 for excitationSite = site (1-based). float32[n_atoms] */
int cgc_get_site_charges(const cgc_context *ctx, int site, float *out);
/* lookup key "<mol>,<atom><ordinal>" of one atom (src/grom2atomFMO.cpp:460) */
int cgc_get_key(const cgc_context *ctx, int atom, char *buf, int cap);
/* frame geometry found by cgc_open: bytes per atom line (incl. newline), column of x, field width, decimals.  .trr: line_bytes
 * is the 12 or 24 bytes of one atom's coordinates; .xtc (frames of varying size): the largest frame's bytes per atom, rounded up. */
int cgc_get_layout(const cgc_context *ctx, int *line_bytes, int *x_col, int *field_width, int *decimals);
/* Index the whole trajectory and return the number of complete frames. */
int cgc_frame_count(cgc_context *ctx, int64_t *n_frames);
/* Byte offset (from the start of the text) of the first atom line of frame f; -1 past the end. */
int64_t cgc_frame_atoms_offset(cgc_context *ctx, int64_t frame);

/* ---- options ----------------------------------------------------------------------------------------------------
 * "sites"        number of excitation sites to compute, 1..nChromo (default nChromo; the reference hard-codes 7,
 *                src/FMOparse.cpp:133)
 * "covariance"   0/1: accumulate n, sum(x-c), sum (x-c)(x-c)^T per site while running (default 1 when it fits)
 * "batch_frames" frames per device batch (default: sized from the frame text size)
 * "ring_slots"   batches in the streaming ring, 2..6 (default 6: three staged ahead by the copier threads, three on the device)
 * "stage_threads" host threads that copy trajectory text into the pinned staging ring (default min(16, cores); a process
 *                that runs one context per GPU divides its cores between them)
 * "profile"      0/1: time every hot-path kernel with CUDA events (see cgc_kernel_times)
 * "mtx_compat"   0/1: also keep, frame by frame in order, the two float32 sums the reference's LITERAL `.mtx` arithmetic
 *                uses (cgc_write_mtx with compat != 0 needs them); default 0
 * Must be set after cgc_open and before the first cgc_run. */
int cgc_set_option(cgc_context *ctx, const char *key, double value);
double cgc_get_option(const cgc_context *ctx, const char *key);

/* ---- the hot path -----------------------------------------------------------------------------------------------
 * cgc_run replaces the per-frame body of main_Cr (src/FMOparse.cpp:68-179): for frames [first, first+n) it streams the
 * text through a ring of 4 pinned staging buffers (filled by background copier threads, two batches ahead) and H2D
 * copies, decodes it (K1), evaluates ComputeCDC_v1 for every site (K2), and accumulates the Correlate sums (K3).
 * dE_kcal (nullable) receives float32 [n][sites][numTot], caller-owned HOST memory.  The rows are bit-identical from run
 * to run and for any batch size or frame range (K2 rounds its partial sums to a fixed grid, so the bin adds are exact).
 * Returns CGC_OK, CGC_EOF with *frames_done < n when the trajectory ends early, or CGC_STOPPED after cgc_request_stop.
 * On an error in the middle of the range (a malformed later frame) *frames_done, the rows delivered so far, the `.time`
 * text and the Correlate sums all cover the same frames: the sound batches before the error. */
int cgc_run(cgc_context *ctx, int64_t first_frame, int64_t n_frames, float *dE_kcal, int64_t *frames_done);
/* Allocate everything the streaming loop needs (device buffers, the pinned staging ring, the Correlate sums; with_time_text
 * != 0: the buffers of the device text formatter and its writer) now instead of inside the first cgc_run /
 * cgc_run_write_time — so that a caller can keep buffer set-up (tens of milliseconds) out of its frame loop.  Optional. */
int cgc_prepare(cgc_context *ctx, int with_time_text);
/* Ask the cgc_run / cgc_run_write_time in progress on this context (or the next one) to stop after the batches already
 * under way — the reference's SIGINT handling, "finish the current frame, then write .mtx" (src/FMOparse.cpp:24-29,62).
 * One atomic store: callable from a signal handler or another thread. */
void cgc_request_stop(cgc_context *ctx);
/* cgc_run that also appends the `.time` rows of the frames it processed to `time_path` (truncate != 0 empties the file
 * first, src/FMOparse.cpp:48-50) — the text of src/FMOparse.cpp:151-165, byte for byte what cgc_write_time prints for the
 * same rows, but formatted ON THE DEVICE (K5, csrc/format_kernel.cu) and written by a background thread, so the host cores
 * are left to stage the trajectory.  dE_kcal may be NULL when the caller wants the file only. */
int cgc_run_write_time(cgc_context *ctx, int64_t first_frame, int64_t n_frames, const char *time_path, int truncate,
                       float *dE_kcal, int64_t *frames_done);
/* The device formatter on its own (HOST in, HOST out): rows x num_col float32 -> "%g" text, ',' between values, '\n' after
 * each row; to_cm != 0 applies float(double(v) * 349.7) first.  *needs_host != 0 when a value was non-finite or outside
 * [1e-16, 1e16) in magnitude (printed as '?': cgc_write_time handles those).  cap >= rows * num_col * 13 always suffices. */
int cgc_format_text_device(cgc_context *ctx, const float *values, int64_t rows, int num_col, int to_cm, char *out,
                           int64_t cap, int64_t *nbytes, int *needs_host);

/* Same computation on text that is already in DEVICE memory (d_text 16-byte aligned; atoms_offset[f] = byte offset of
 * the first atom line of frame f, HOST array).  d_dE_kcal: DEVICE float32 [n_frames][sites][numTot].  Work is enqueued
 * on `stream` (a cudaStream_t, NULL = default stream) and NOT synchronised.  Any n_frames >= 1 (the scratch buffers grow
 * to the largest batch seen).  The decode kernel fetches whole 16-byte units: the allocation behind d_text must extend to
 * the next multiple of 16 bytes after text_bytes.  cgc_run_device and cgc_decode_device share one set of scratch buffers
 * per context: use ONE stream per context for them and do not interleave them with cgc_run. */
int cgc_run_device(cgc_context *ctx, const void *d_text, int64_t text_bytes, const int64_t *atoms_offset,
                   int n_frames, float *d_dE_kcal, void *stream);

/* Decode only (K1): float32 [n_frames][n_atoms][3] = x,y,z into caller-owned DEVICE memory; for parity tests of the
 * decoder against (float)atof (src/grom2atomFMO.cpp:448-450). */
int cgc_decode_device(cgc_context *ctx, const void *d_text, int64_t text_bytes, const int64_t *atoms_offset,
                      int n_frames, float *d_xyz, void *stream);

 /* With option "profile" = 1 every hot-path kernel launch is bracketed by CUDA events on its own stream; this returns the
 * accumulated device time in ms and the launch count per kernel: [0] K1 decode, [1] K2 CDC pair sum, [2] FP64 bins ->
 * float32 rows, [3] K3 Correlate.  Synchronises on the recorded events.  reset != 0 clears the totals. */
int cgc_kernel_times(cgc_context *ctx, double *ms4, int64_t *launches4, int reset);
/* With option "profile" = 1 the host->device copies of the trajectory text (cgc_run's copy stream) are bracketed by events
 * too: accumulated device time of those copies in ms and their count.  Copy time / wall time of a run = how busy the link was. */
int cgc_copy_times(cgc_context *ctx, double *ms, int64_t *copies, int reset);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t cgc_launch_count(const cgc_context *ctx);
/* Developer aid.  With CGC_GUARD=1 in the environment (read once per process) every device buffer the contexts own carries
 * 4 KiB of a known byte pattern in front of it and behind it; this reads all bands back: *n_buffers guarded buffers alive,
 * *n_damaged of them with a changed band, i.e. some kernel wrote out of bounds.  CGC_ERR_STATE when guarding is off. */
int cgc_debug_guard_check(int *n_buffers, int *n_damaged);
/* device-side decode error flags accumulated so far (0 = every coordinate field was well-formed) */
int cgc_decode_errors(cgc_context *ctx, int *flags);

/* ---- Correlate ("MergeDE", src/FMOparse.cpp:168-221) ---------------------------------------------------------------
 * Partial sums live in one flat DEVICE buffer of doubles:  [ n | shift c (S*G) | sum(x-c) (S*G) | sum (x-c)(x-c)^T (S*G*G) ]
 * with x in kcal/mol and c the dE row of the trajectory's frame 0 (identical on every rank), so that buffers of
 * different ranks can be added element-wise except for the shift block, which is identical on all of them.
 * cgc_cov_bind lets the caller own that buffer (e.g. a torch tensor that is then all-reduced over NCCL). */
int64_t cgc_cov_size(const cgc_context *ctx);                       /* number of doubles; 0 when covariance is off */
int cgc_cov_bind(cgc_context *ctx, double *d_buffer);               /* DEVICE pointer, cgc_cov_size doubles, zeroed here */
int cgc_cov_reset(cgc_context *ctx);
int cgc_cov_partials(cgc_context *ctx, double *host_out);           /* copy the flat buffer to HOST */
/* The ONE collective of the hot path for a process that drives several GPUs itself (one context and one host thread per
 * GPU, each run over its own frame range — what the reference's authors did with one process per trajectory chunk,
 * scripts/2014-06-gro2dEcsv.sh:3-27): one grouped ncclAllReduce(sum, double) of n, sum(x-c) and sum (x-c)(x-c)^T; the shift
 * block, identical everywhere, is left alone.  Afterwards every context holds the sums of all frames.
 * cgc_comm_create = ncclCommInitAll over `devices` (in the order of the contexts handed to cgc_cov_allreduce later).  It
 * takes about a second: call it from a helper thread while the frames stream.  comm == NULL makes cgc_cov_allreduce create
 * (and destroy) a communicator itself.  NCCL is bound at run time (libnccl.so.2): CGC_ERR_UNSUPPORTED when it cannot be
 * loaded; the message of a failed cgc_comm_create is cgc_last_error(NULL) on the calling thread.  The float32 sums of the
 * literal (compat) `.mtx` are per context and are NOT combined: their sequential order is the point of that mode.  (One
 * process per GPU instead: all-reduce the cgc_cov_bind buffer with the caller's own communicator,
 * cgromcorrfmo_b200/dist.py.) */
int cgc_comm_create(const int *devices, int n, cgc_comm **out);
void cgc_comm_destroy(cgc_comm *comm);
int cgc_cov_allreduce(cgc_context **ctxs, int n, cgc_comm *comm);
/* Checkpoint / resume of the Correlate state (the reference loses its accumulators when a run is killed; its `.time`
 * prefix survives because the file only grows, src/FMOparse.cpp:48-50,158-165 — SURVEY.md section 5).  The blob holds
 * the flat partial sums and the two float32 sums of the literal `.mtx`; loading it into a context opened on the same
 * system continues the accumulation exactly where cgc_cov_save_state left off.  HOST buffers. */
int64_t cgc_cov_state_bytes(const cgc_context *ctx);
int cgc_cov_save_state(cgc_context *ctx, void *host_out, int64_t cap);
int cgc_cov_load_state(cgc_context *ctx, const void *host_in, int64_t nbytes);
/* Set the shift c from a DEVICE row float32 [sites][numTot] (the dE row of the trajectory's frame 0).  cgc_run does this
 * itself; cgc_cov_accumulate_device falls back to the first row it is given when no shift was set. */
int cgc_cov_set_shift_device(cgc_context *ctx, const float *d_row, void *stream);
/* Accumulate rows that are already on the device: d_dE float32 [n_frames][sites][numTot] */
int cgc_cov_accumulate_device(cgc_context *ctx, const float *d_dE_kcal, int n_frames, void *stream);
/* mean [S][G] and covariance [S][G][G] (HOST, double) from the partial sums; ddof 0 = population (the formula the
 * reference intends, src/FMOparse.cpp:189-206), 1 = np.cov as used by depca (scripts/depca/depca/sidechain_corr.py:67). */
int cgc_cov_finalize(cgc_context *ctx, int ddof, double *mean, double *cov);
/* Eigen-decomposition of symmetric matrices on the device (K6: batched one-sided Jacobi) — numpy.linalg.eigh in depca's
 * ComputeModes (scripts/depca/depca/sidechain_corr.py:91-96).  HOST in, HOST out: matrices float64 [n_mat][n][n];
 * evals [n_mat][n] ascending; evecs [n_mat][n][n] with ROW k the eigenvector of evals[k] (numpy returns them as columns:
 * depca transposes, sidechain_corr.py:97).  *sweeps (nullable) = Jacobi sweeps used.  Signs of eigenvectors are arbitrary,
 * as with LAPACK.  24 matrices of 1099^2: 0.55 s against 1.9 s for LAPACK on the box's 16 host cores. */
int cgc_eigh(cgc_context *ctx, const double *matrices, int n_mat, int n, double *evals, double *evecs, int *sweeps);
/* Mode projection, the step that follows the covariance in the pyPCA pipeline (depca.convert.ApplyPCA_hdf5,
 * scripts/depca/depca/convert.py:84-125, and the np.inner at the end of sidechain_corr.py's usage): per site,
 *   out[t][s][m] = sum_g (dE_rows[t][s][g] - mean[s][g]) * modes[s][m][g]
 * computed on FP64 tensor cores (K4).  All pointers are HOST memory: dE_rows float32 [n_frames][sites][numTot] (what
 * cgc_run returned, any unit), mean float64 [sites][numTot], modes float64 [sites][n_modes][numTot] (eigenvectors as
 * rows, as depca's ComputeModes returns them), out float64 [n_frames][sites][n_modes]. */
int cgc_project(cgc_context *ctx, const float *dE_rows, int64_t n_frames, const double *mean, const double *modes,
                int n_modes, double *out);
/* Time-correlation functions of dE series (K7), the last step of the pyPCA chain: depca's TimeCorr and HDFStoreCt_v1/v2/v3
 * (scripts/depca/depca/sidechain_corr.py:133-139, 170-182, 190-218, 221-242).  For every pair p,
 *   out[p][tau] = sum_t  s_first[t] * s_second[t + tau] / (n - tau),      tau = 0 .. corr_len - 1,
 * the sum over t = 0 .. n - tau - 1 (circular == 0: np.correlate / the np.inner loop of v2), or over t = 0 .. n - 1 with
 * t + tau taken modulo n (circular != 0: what v3's ifft(fft(f2) * conj(fft(f1))) evaluates).  v2 and v3 correlate
 * (first, second) = (f1, f2); TimeCorr / v1 is the same call with the pair swapped.  Evaluated directly in FP64 (n * corr_len
 * fused multiply-adds per pair), deterministic.  HOST in, HOST out: series float64 [n_series][n], already mean-free (the
 * scripts subtract the mean of the WHOLE stored series before windowing, :179-180); pair_first / pair_second int32
 * [n_pairs] indices into the series; out float64 [n_pairs][corr_len]; 1 <= corr_len <= n.  *kernel_ms (nullable) = device
 * time of the correlation kernels (CUDA events).  Needs no open trajectory. */
int cgc_time_corr(cgc_context *ctx, const double *series, int n_series, int64_t n, const int32_t *pair_first,
                  const int32_t *pair_second, int n_pairs, int64_t corr_len, int circular, double *out, double *kernel_ms);

/* `.time` / `.mtx` text back into numbers (K8, the inverse of the device formatter): what the reference's consumers do with
 * `[[float(x) for x in l.split(',')] for l in f.strip().split("\n")]` (scripts/2014-06-csv2hdf5.py:56-65) and genfromtxt
 * (scripts/depca/depca/dedata.py:96-120).  HOST in, HOST out.  text: rows of values separated by ',' and ended by '\n'
 * (surrounding whitespace is stripped, a missing final newline is supplied).  *num_col in: expected values per row, 0 = take
 * the count of the first row; out: the count used.  values: float64 [rows][*num_col], every entry BIT-IDENTICAL to strtod /
 * Python float of its token; NULL = count only (*rows_out is then the number of rows, for sizing the buffer).  Tokens of the
 * form [-+]digits[.digits][e[-+]digits] with <= 15 significant digits and |decimal exponent| <= 22 — everything "%g" prints
 * for finite values between 1e-17 and 1e27 — are converted on the device (one correctly rounded multiply or divide by an exact
 * power of ten); a chunk holding anything else (nan, inf, longer tokens) is re-read by the host with strtod.
 * CGC_ERR_INPUT_MALFORMED when a row has another number of values or a token is not a number (the message names the row).
 * *kernel_ms (nullable) = device time of the kernels.  Needs no open trajectory. */
int cgc_parse_text(cgc_context *ctx, const char *text, int64_t nbytes, int *num_col, double *values, int64_t cap_values,
                   int64_t *rows_out, double *kernel_ms);

/* ---- writers ----------------------------------------------------------------------------------------------------
 * cgc_write_time appends rows to <path> exactly as src/FMOparse.cpp:151-165 does: value = float(double(kcal)*349.7),
 * printed with "%g" (6 significant digits), ',' separated, one line per (frame, site).  truncate != 0 empties the file
 * first (src/FMOparse.cpp:48-50). */
int cgc_write_time(const char *path, int truncate, const float *dE_kcal, int64_t rows, int num_tot);
/* cgc_write_mtx writes sites*numTot lines of numTot values (src/FMOparse.cpp:208-221).  compat == 0: population
 * covariance in (kcal/mol)^2; compat != 0: the reference's literal arithmetic including its zero stride
 * (src/FMOparse.cpp:42,193,202), i.e. sum_t x_0 x_j / T^numTot - <x_i><x_j> in float32 — needs the option "mtx_compat" set
 * before the frames were run (CGC_ERR_ARG otherwise) and a single context (the float32 sums are sequential by nature). */
int cgc_write_mtx(cgc_context *ctx, const char *path, int compat);

#ifdef __cplusplus
}
#endif
#endif /* CGROMCORR_H */

// Device kernels of the traj2dEcsv hot path (sm_100a only) — launch wrappers.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace cgc {

#ifndef CGC_TILE_THREADS
#define CGC_TILE_THREADS 256          // overridden only by scripts/k2_variants.cu builds
#endif
constexpr int kTileThreads = CGC_TILE_THREADS;   // threads per CTA of the pair kernel
constexpr int kAtomsPerThread = 8;  // environment atoms held in registers by one thread
constexpr int kTileAtoms = kTileThreads * kAtomsPerThread;
constexpr int kDecodeThreads = 256; // atom lines per CTA of the decode kernel

// decode error flags (device -> host)
enum : int {
  kDecBadChar = 1,      // a coordinate field holds something other than [ -.0-9]
  kDecBadPoint = 2,     // the decimal point is not where the frame-0 layout says
  kDecBadNewline = 4,   // an atom line does not end where the frame-0 layout says
  kDecBadXtc = 8,       // an .xtc frame's header or bit stream is impossible (wrong atom count, a run past the last atom, ...)
};

struct DeviceSystem {
  // static, built once per trajectory
  int n_atoms = 0, n_pad = 0;           // n_pad: n_atoms rounded up to kTileAtoms
  int n_chromo = 0, n_groups = 0, num_tot = 0;
  int n_sites = 0;                      // sites computed (<= n_chromo)
  int site_cap = 0;                     // P: dQ atoms per site, padded to the largest site
  int line_bytes = 0, x_col = 20, field_w = 8, ndec = 3;
  int traj_mode = 0;                    // 0 .gro text; 1 / 2: .trr coordinates, single / double (line_bytes = 12 / 24);
                                        // 3: .xtc compressed coordinates (frames of varying size, launch_xtc_decode)
  double* kq = nullptr;                 // [n_pad] (double)33.2f * (double)q
  int32_t* col = nullptr;               // [n_pad] dE column, -1 = contributes nowhere (pads)
  int32_t* slot = nullptr;              // [n_pad] s*P + c for the c-th dQ atom of site s, else -1
  float* slot_dq = nullptr;             // [n_sites*P] dQ per slot (0 for padding slots)
};

// K1: fixed-column text -> xyz8 (3 floats per atom in blocks of 8 atoms [x0..x7][y0..y7][z0..z7], n_pad atoms per
// frame) + per-frame site blocks.
// text: device pointer, 16-byte aligned.  atoms_off: DEVICE int64[n_frames].
void launch_decode(const DeviceSystem& sys, const char* text, int64_t text_bytes, const int64_t* atoms_off,
                   int n_frames, float* xyz8, float4* sites, int* err_flags, cudaStream_t st);
// K1x (xtc_kernel.cu): the same outputs from GROMACS .xtc frames.  coords_off: DEVICE int64[n_frames], offset of every
// frame's coordinate block (its atom-count field; a multiple of 4) inside `text`.  scratch: xtc_scratch_bytes() bytes.
size_t xtc_scratch_bytes(const DeviceSystem& sys, int n_frames);
void launch_xtc_decode(const DeviceSystem& sys, const char* text, int64_t text_bytes, const int64_t* coords_off, int n_frames,
                       void* scratch, float* xyz8, float4* sites, int* err_flags, cudaStream_t st);
// ... and its two passes on their own: the index pass (checkpoints into `scratch`) needs nothing but the frames' bytes, so
// the streaming loop runs it on a stream of its own as soon as a batch has been copied — beside the previous batch's kernels.
void launch_xtc_index(const DeviceSystem& sys, const char* text, int64_t text_bytes, const int64_t* coords_off, int n_frames,
                      void* scratch, int* err_flags, cudaStream_t st);
void launch_xtc_unpack(const DeviceSystem& sys, const char* text, int64_t text_bytes, const int64_t* coords_off, int n_frames,
                       const void* scratch, float* xyz8, float4* sites, int* err_flags, cudaStream_t st);
// xyz8 -> plain float32 [n_frames][n_atoms][3]
void launch_unblock(const DeviceSystem& sys, const float* xyz8, float* xyz, int n_frames, cudaStream_t st);

// fill every slot of the per-frame site blocks with the neutral padding entry
void launch_init_sites(const DeviceSystem& sys, float4* sites, int n_frames, cudaStream_t st);

// K2: CDC pair sums.  acc: double [n_frames][n_sites][num_tot], must be zero on entry.  The partials are rounded to a grid
// of 2^-36 kcal/mol before they are added, so the adds are exact and the bins do not depend on the order of the atomics.
void launch_cdc(const DeviceSystem& sys, const float* xyz8, const float4* sites, double* acc, int n_frames,
                int sm_count, cudaStream_t st);
size_t cdc_smem_bytes(const DeviceSystem& sys);
int cdc_pad_site_cap(int p);  // dQ atoms per site, padded to what the kernel's loop structure wants

// acc (double) -> dE (float32), and re-zero acc for the next batch.
void launch_finish(double* acc, float* dE, int64_t n, cudaStream_t st);

// K3: per-site shifted sums.  buf = [ n | c (S*G) | sum (S*G) | sumsq (S*G*G) ]
void launch_cov_set_shift(const float* dE_row0, double* buf, int S, int G, cudaStream_t st);
// centred_scratch: cov_centred_doubles(n_frames, S, G) doubles (the centred rows the SYRK consumes, columns padded to 16,
// and partial column sums).  compat == null: the reference's literal float32 sums are not kept.
// Returns an error when the TMA tensor map cannot be built (driver entry point missing).
int64_t cov_padded_cols(int G);
int64_t cov_centred_doubles(int n_frames, int S, int G);
cudaError_t launch_cov_accumulate(const float* dE, int n_frames, double* buf, float* compat /* float32 [2][S*G] or null */,
                                  double* centred_scratch, int S, int G, cudaStream_t st);
// K4: out[t][s][m] = sum_g (x[t][s][g] - mean[s][g]) * modes[s][m][g]  (all DEVICE; centred: n_frames*S*G doubles scratch)
void launch_project(const float* x, int n_frames, const double* mean, const double* modes, int n_modes, double* centred,
                    double* out, int S, int G, cudaStream_t st);
// K5: rows of float32 -> the reference's `.time` text on the device (format_kernel.cu).  d_row_off: rows + 1 int64 (scratch;
// [rows] = total bytes afterwards); d_out: format_time_capacity(rows, ncol) bytes; *d_flags |= 1 when a value needs the host.
int64_t format_time_capacity(int64_t rows, int ncol);
void launch_format_time(const float* d_x, int64_t rows, int ncol, bool to_cm, int64_t* d_row_off, char* d_out, int* d_flags,
                        cudaStream_t st);
// K6 (eigh_kernel.cu): batched one-sided Jacobi.  Eigenvectors end up as the rows of d_Vt; d_evals is unsorted.
int eigh_jacobi(const double* d_A, int n_mat, int n, double* d_Wt, double* d_Vt, double* d_evals, unsigned int* d_rot,
                int max_sweeps, cudaStream_t st, int64_t* launches);
// K7 (timecorr_kernel.cu): out[p][tau] = sum_t f_first[t] * f_second[t + tau] / (n - tau), tau < corr_len (t + tau modulo n
// when circular).  d_series: rows of time_corr_series_stride(n) doubles, ZERO from element n on.  d_scratch: the doubles
// time_corr_plan asks for.  Returns the number of kernels launched.
int64_t time_corr_series_stride(int64_t n);
void time_corr_plan(int64_t n, int64_t corr_len, int n_pairs, bool circular, int sm_count, int* splits, int* wrap_splits,
                    int64_t* scratch_doubles);
int launch_time_corr(const double* d_series, int64_t n, const int32_t* d_first, const int32_t* d_second, int n_pairs,
                     int64_t corr_len, bool circular, int sm_count, double* d_scratch, double* d_out, cudaStream_t st);
// K8 (parse_kernel.cu): "%g" CSV text -> float64, bit-exact with strtod.  d_text: parse_text_padded_bytes(n_bytes) bytes,
// ZERO behind the text, 16-byte aligned.  launch_parse_count fills d_tile_counts (parse_text_tiles ints) and d_totals
// ([0] delimiters, [1] newlines); launch_parse writes value i = token i into d_out (cap doubles) and sets *d_flags (1: a
// token needs the host's strtod, 2: a row does not have num_col values, *d_bad_row = the first such row).
int64_t parse_text_padded_bytes(int64_t n_bytes);
int64_t parse_text_tiles(int64_t n_bytes);
int64_t parse_text_scan_bytes(int64_t n_bytes);   // scratch of launch_parse (d_tile_offsets)
void launch_parse_count(const char* d_text, int64_t n_bytes, int* d_tile_counts, unsigned long long* d_totals, int sm_count,
                        cudaStream_t st);
void launch_parse(const char* d_text, int64_t n_bytes, const int* d_tile_counts, int64_t* d_tile_offsets, int num_col,
                  int64_t cap, double* d_out, int* d_flags, unsigned long long* d_bad_row, int sm_count, cudaStream_t st);
void launch_narrow(const double* in, float* out, int64_t n, cudaStream_t st);   // out[i] = (float)in[i]
void launch_cov_finalize(const double* buf, int S, int G, int ddof, double* mean, double* cov, cudaStream_t st);

}  // namespace cgc

// K7 — time-correlation functions of dE series on the device ("the following step").
//
// depca correlates two mean-free series f1, f2 of one chromophore's dE columns over the lag tau:
//   TimeCorr / HDFStoreCt_v1   np.correlate(f1, f2, 'full')[n-1:] / (n - tau)        (scripts/depca/depca/sidechain_corr.py:133-139, 170-182)
//   HDFStoreCt_v2              np.inner(f1[0:n-tau], f2[tau:n]) / (n - tau)          (:190-218, a Python loop over tau)
//   HDFStoreCt_v3              ifft(fft(f2) * conj(fft(f1)))[0:corr_len] / (n - tau) (:221-242, i.e. the CIRCULAR sum)
// All three are   C[tau] = sum_t f1[t] * f2[t + tau] / (n - tau)   with different index conventions (v1 swaps the roles
// of f1 and f2) and, for v3, indices taken modulo n.  Here the sum is evaluated directly in FP64: n * corr_len fused
// multiply-adds per pair of series (8e9 FLOP for the scripts' defaults n = 200 000, corr_len = 20 000), which the FP64
// pipe finishes in a fraction of a millisecond — no FFT, no O(n log n) rounding pattern, and the linear variants are not
// an approximation of anything.
//
// Kernel: a CTA owns 1024 consecutive lags of one pair and a range of time chunks; a thread owns 8 consecutive lags.
// For a block of 8 time steps the thread needs f1[t..t+8) (the same for every thread: broadcast LDS.128) and the 15
// values f2[t + tau0 + 8j .. +15), i.e. two 8-blocks of f2, of which the second is the first of the next step.  f2 is staged
// in shared memory TRANSPOSED by blocks (element k of block b at [k][b]) so that the 128 threads of the CTA, whose windows
// are 8 elements apart, read consecutive addresses: 8 conflict-free LDS.64 + 4 broadcast LDS.128 feed 64 DFMAs.  Series are
// zero-padded behind their end, so the range of t needs no per-lag bound: the padded products add exactly 0.  Partial
// sums of the time ranges are written to a scratch array and added in a fixed order by the finish kernel (deterministic).
// The circular variant adds the wrapped terms  sum_{u<tau} f2[u] * f1[n - tau + u]  — the same kernel on the last
// corr_len samples of f1 against the first corr_len of f2.
#include "kernels.cuh"

#include <algorithm>
#include <cstdint>

namespace cgc {

namespace {

constexpr int kTcThreads = 128;
constexpr int kTcLags = 8;                                   // lags per thread
constexpr int kTcLagBlock = kTcThreads * kTcLags;            // lags per CTA
constexpr int kTcChunk = 512;                                // time steps per shared-memory fill
constexpr int kTcBlocks = (kTcChunk + kTcLagBlock) / 8 + 1;  // 8-blocks of f2 a chunk touches
constexpr int kTcBlockStride = 194;                          // >= kTcBlocks and = 2 mod 16: conflict-free transposed fill
static_assert(kTcBlockStride >= kTcBlocks && kTcBlockStride % 16 == 2, "shared-memory layout of the f2 window");

// acc[q] += sum_i f[i] * w[i + q], the window w being lo[0..7] followed by hi[0..7]
__device__ __forceinline__ void mac_block(const double (&f)[8], const double (&lo)[8], const double (&hi)[8],
                                          double (&acc)[kTcLags]) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
#pragma unroll
    for (int q = 0; q < kTcLags; q++) {
      const int k = i + q;
      acc[q] = fma(f[i], k < 8 ? lo[k] : hi[k - 8], acc[q]);
    }
  }
}

__device__ __forceinline__ void load_f2_block(const double* __restrict__ s2, int b, double (&w)[8]) {
#pragma unroll
  for (int k = 0; k < 8; k++) w[k] = s2[k * kTcBlockStride + b];
}

__device__ __forceinline__ void load_f1_block(const double* __restrict__ s1, int s, double (&f)[8]) {
  const double2* p = reinterpret_cast<const double2*>(s1 + 8 * s);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const double2 v = p[k];
    f[2 * k] = v.x;
    f[2 * k + 1] = v.y;
  }
}

// grid (lag blocks, time splits, pairs).  first/second: index lists into `series` (rows of `stride` doubles, zero behind
// their data); the second series is read from `second_off` on.  n_t: time steps of the first series that carry data.
// part: [pair][split][lag_pad] unnormalised sums.
__global__ void __launch_bounds__(kTcThreads)
time_corr_kernel(const double* __restrict__ series, int64_t stride, const int32_t* __restrict__ first,
                 const int32_t* __restrict__ second, int64_t second_off, int64_t n_t, int64_t lag_pad,
                 double* __restrict__ part) {
  __shared__ __align__(16) double s1[kTcChunk];
  __shared__ double s2[8 * kTcBlockStride];
  const int j = threadIdx.x;
  const int64_t tau0 = (int64_t)blockIdx.x * kTcLagBlock;
  const int64_t pair = blockIdx.z;
  const double* __restrict__ f1 = series + (int64_t)first[pair] * stride;
  const double* __restrict__ f2 = series + (int64_t)second[pair] * stride + second_off + tau0;

  // time steps that can pair with a lag of this block: t < n_t - tau0; cut into chunks, chunks dealt out to the splits
  const int64_t n_eff = n_t - tau0;
  const int64_t chunks = n_eff > 0 ? (n_eff + kTcChunk - 1) / kTcChunk : 0;
  const int64_t per_split = (chunks + gridDim.y - 1) / gridDim.y;
  const int64_t c_begin = per_split * blockIdx.y;
  const int64_t c_end = min(chunks, c_begin + per_split);

  double acc[kTcLags];
#pragma unroll
  for (int q = 0; q < kTcLags; q++) acc[q] = 0.0;

  for (int64_t c = c_begin; c < c_end; c++) {
    const int64_t t0 = c * kTcChunk;
    __syncthreads();   // the previous chunk has been consumed
    for (int e = j; e < kTcChunk; e += kTcThreads) s1[e] = f1[t0 + e];
    for (int e = j; e < kTcBlocks * 8; e += kTcThreads) s2[(e & 7) * kTcBlockStride + (e >> 3)] = f2[t0 + e];
    __syncthreads();
    double f[8], wa[8], wb[8];
    load_f2_block(s2, j, wa);
#pragma unroll 1
    for (int s = 0; s < kTcChunk / 8; s += 2) {
      load_f2_block(s2, s + j + 1, wb);
      load_f1_block(s1, s, f);
      mac_block(f, wa, wb, acc);
      load_f2_block(s2, s + j + 2, wa);
      load_f1_block(s1, s + 1, f);
      mac_block(f, wb, wa, acc);
    }
  }
  double* out = part + (pair * gridDim.y + blockIdx.y) * lag_pad + tau0 + (int64_t)j * kTcLags;
#pragma unroll
  for (int q = 0; q < kTcLags; q += 2) *reinterpret_cast<double2*>(out + q) = make_double2(acc[q], acc[q + 1]);
}

// out[pair][tau] = (sum over splits of part + wrapped term) / (n - tau)
__global__ void time_corr_finish_kernel(const double* __restrict__ part, int splits, const double* __restrict__ wrap,
                                        int wrap_splits, int64_t lag_pad, int64_t wrap_pad, int64_t n, int64_t corr_len,
                                        int64_t total, double* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (; i < total; i += step) {
    const int64_t pair = i / corr_len, tau = i - pair * corr_len;
    double v = 0.0;
    for (int s = 0; s < splits; s++) v += part[(pair * splits + s) * lag_pad + tau];
    if (wrap != nullptr && tau > 0) {
      double w = 0.0;
      for (int s = 0; s < wrap_splits; s++) w += wrap[(pair * wrap_splits + s) * wrap_pad + (corr_len - tau)];
      v += w;
    }
    out[i] = v / (double)(n - tau);
  }
}

int pick_splits(int64_t n_t, int64_t lag_blocks, int n_pairs, int sm_count) {
  const int64_t chunks = (n_t + kTcChunk - 1) / kTcChunk;
  const int64_t want = (int64_t)sm_count * 8;   // CTAs that fill the machine twice over at ~4 resident per SM
  const int64_t have = lag_blocks * n_pairs;
  int64_t s = (want + have - 1) / have;
  s = std::min<int64_t>(s, std::max<int64_t>(1, chunks / 4));   // at least four chunks per CTA
  return (int)std::max<int64_t>(1, std::min<int64_t>(s, 65535));
}

}  // namespace

int64_t time_corr_series_stride(int64_t n) { return (n + kTcChunk + kTcLagBlock + 16 + 7) & ~(int64_t)7; }

int64_t time_corr_lag_pad(int64_t corr_len) { return (corr_len + kTcLagBlock - 1) / kTcLagBlock * kTcLagBlock; }

void time_corr_plan(int64_t n, int64_t corr_len, int n_pairs, bool circular, int sm_count, int* splits, int* wrap_splits,
                    int64_t* scratch_doubles) {
  const int64_t lb = time_corr_lag_pad(corr_len) / kTcLagBlock;
  *splits = pick_splits(n, lb, n_pairs, sm_count);
  int64_t need = (int64_t)n_pairs * *splits * lb * kTcLagBlock;
  *wrap_splits = 0;
  if (circular) {
    const int64_t wl = time_corr_lag_pad(corr_len + 1);
    *wrap_splits = pick_splits(corr_len, wl / kTcLagBlock, n_pairs, sm_count);
    need += (int64_t)n_pairs * *wrap_splits * wl;
  }
  *scratch_doubles = need;
}

// d_series: n_series rows of time_corr_series_stride(n) doubles, zero from element n on.  d_first/d_second: n_pairs series
// indices.  d_out: [n_pairs][corr_len].  d_scratch: as sized by time_corr_plan.  out[p][tau] = sum_t f1[t] f2[t + tau] /
// (n - tau), t + tau taken modulo n when circular.  Returns the number of kernels launched.
int launch_time_corr(const double* d_series, int64_t n, const int32_t* d_first, const int32_t* d_second, int n_pairs,
                     int64_t corr_len, bool circular, int sm_count, double* d_scratch, double* d_out, cudaStream_t st) {
  const int64_t stride = time_corr_series_stride(n);
  const int64_t lag_pad = time_corr_lag_pad(corr_len);
  int splits = 1, wrap_splits = 0;
  int64_t scratch = 0;
  time_corr_plan(n, corr_len, n_pairs, circular, sm_count, &splits, &wrap_splits, &scratch);
  int launches = 0;
  double* part = d_scratch;
  time_corr_kernel<<<dim3((unsigned)(lag_pad / kTcLagBlock), (unsigned)splits, (unsigned)n_pairs), kTcThreads, 0, st>>>(
      d_series, stride, d_first, d_second, 0, n, lag_pad, part);
  launches++;
  double* wrap = nullptr;
  int64_t wrap_pad = 0;
  if (circular) {
    // wrapped terms: W[lambda] = sum_{u < corr_len} f2[u] * f1[n - corr_len + u + lambda], lambda = corr_len - tau in [1, corr_len]
    wrap_pad = time_corr_lag_pad(corr_len + 1);
    wrap = part + (int64_t)n_pairs * splits * lag_pad;
    time_corr_kernel<<<dim3((unsigned)(wrap_pad / kTcLagBlock), (unsigned)wrap_splits, (unsigned)n_pairs), kTcThreads, 0, st>>>(
        d_series, stride, d_second, d_first, n - corr_len, corr_len, wrap_pad, wrap);
    launches++;
  }
  const int64_t total = (int64_t)n_pairs * corr_len;
  time_corr_finish_kernel<<<(unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count * 16), 256, 0, st>>>(
      part, splits, wrap, wrap_splits, lag_pad, wrap_pad, n, corr_len, total, d_out);
  launches++;
  return launches;
}

}  // namespace cgc

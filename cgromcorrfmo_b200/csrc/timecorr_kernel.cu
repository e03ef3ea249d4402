// K7 — time-correlation functions of dE series on FP64 tensor cores ("the following step").
//
// depca correlates two mean-free series f1, f2 of one chromophore's dE columns over the lag tau:
//   TimeCorr / HDFStoreCt_v1   np.correlate(f1, f2, 'full')[n-1:] / (n - tau)        (scripts/depca/depca/sidechain_corr.py:133-139, 170-182)
//   HDFStoreCt_v2              np.inner(f1[0:n-tau], f2[tau:n]) / (n - tau)          (:190-218, a Python loop over tau)
//   HDFStoreCt_v3              ifft(fft(f2) * conj(fft(f1)))[0:corr_len] / (n - tau) (:221-242, i.e. the CIRCULAR sum)
// All three are   C[tau] = sum_t f1[t] * f2[t + tau] / (n - tau)   with different index conventions (v1 swaps the roles
// of f1 and f2) and, for v3, indices taken modulo n.  Here the sum is evaluated directly in FP64: n * corr_len fused
// multiply-adds per pair of series (8e9 FLOP for the scripts' defaults n = 200 000, corr_len = 20 000) — no FFT, no
// O(n log n) rounding pattern, and the linear variants are not an approximation of anything.
//
// As a tensor-core contraction.  Cut time into rows of 8: X[p][i] = f1[8p + i], and for a block lag 8M
// Y_M[p][j] = f2[8(p + M) + j].  The 8 x 8 Gram tile  G_M = X^T Y_M = sum_p X[p][i] Y_M[p][j]  holds every product of lag
// 8M + j - i exactly once, so
//     C[8M + d] = sum_{i + d < 8} G_M[i][i + d]  +  sum_{i + d >= 8} G_{M+1}[i][i + d - 8]          (diagonals of two tiles)
// and G_M accumulates with one DMMA m8n8k4 per 4 rows of p: A = 4 rows of X (the same for every M), B = rows p + M .. p + M + 3
// of f2 seen as an 8-wide matrix.  A warp owns 16 consecutive tiles (128 lags); the B fragment of tile M at step k is the
// fragment of tile M + 4 at step k - 1, so a ring of 16 fragment registers is refilled with 4 shared-memory loads per step
// of 16 DMMAs, and a lane's 32 fragment addresses are 32 consecutive doubles (conflict-free without padding).  Why the
// tensor pipe and not DFMA: scripts/pipe_bench6.cu — DFMA is bound by register-file bandwidth (42 FMA lanes/clk/SM with
// three fresh operands, 56 with one reused; the first, DFMA version of this kernel reached 0.72-0.77 of the pipe), DMMA
// m8n8k4 sustains 63.7 of 64.
// Series are zero-padded behind their end, so the range of t needs no per-lag bound: the padded products add exactly 0.
// The 15 diagonal sums of every tile go to a scratch array per time split; the finish kernel adds them in a fixed order
// (deterministic).  The circular variant adds the wrapped terms  sum_{u<tau} f2[u] * f1[n - tau + u]  — the same kernel on
// the last corr_len samples of f1 against the first corr_len of f2.
#include "kernels.cuh"

#include <algorithm>
#include <cstdint>
#include <cstdlib>

namespace cgc {

namespace {

constexpr int kTcWarps = 4;
constexpr int kTcThreads = kTcWarps * 32;
constexpr int kTcTiles = 16;                                 // Gram tiles per warp (8 lags each)
constexpr int kTcLagBlock = kTcWarps * kTcTiles * 8;         // lags per CTA
constexpr int kTcChunk = 1024;                               // time steps per shared-memory fill = 32 DMMA steps
constexpr int kTcSteps = kTcChunk / 32;
constexpr int kTcF2 = kTcChunk + kTcLagBlock + 32;           // f2 window of a chunk (+ the prefetch of the step after the last)
static_assert(kTcTiles == 16 && kTcSteps % 4 == 0, "the fragment ring has 16 slots and turns once in 4 steps");

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// grid (lag blocks, time splits, pairs).  first/second: index lists into `series` (rows of `stride` doubles, zero behind
// their data); the second series is read from `second_off` on.  n_t: time steps of the first series that carry data.
// part: [pair][split][tile][16]: slot d' + 7 of tile M = sum of G_M[i][j] over j - i = d' (slot 15 unused).
__global__ void __launch_bounds__(kTcThreads)
time_corr_kernel(const double* __restrict__ series, int64_t stride, const int32_t* __restrict__ first,
                 const int32_t* __restrict__ second, int64_t second_off, int64_t n_t, int64_t tiles_pad,
                 double* __restrict__ part) {
  __shared__ double s1[kTcChunk];
  __shared__ double s2[kTcF2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;   // DMMA fragment coordinates: A[gid][tig], B[tig][gid], C[gid][2 tig + {0, 1}]
  const int64_t tau0 = (int64_t)blockIdx.x * kTcLagBlock;
  const int64_t pair = blockIdx.z;
  const double* __restrict__ f1 = series + (int64_t)first[pair] * stride;
  const double* __restrict__ f2 = series + (int64_t)second[pair] * stride + second_off + tau0;

  // time steps that can pair with a lag of this block (its first tile reaches down to tau0 - 7): t < n_t - tau0 + 7
  const int64_t n_eff = n_t - tau0 + 7;
  const int64_t chunks = n_eff > 0 ? (n_eff + kTcChunk - 1) / kTcChunk : 0;
  const int64_t per_split = (chunks + gridDim.y - 1) / gridDim.y;
  const int64_t c_begin = per_split * blockIdx.y;
  const int64_t c_end = min(chunks, c_begin + per_split);

  double acc[kTcTiles][2];
#pragma unroll
  for (int m = 0; m < kTcTiles; m++) acc[m][0] = acc[m][1] = 0.0;

  const double* a_ptr = s1 + 8 * tig + gid;                           // A[i = gid][p' = tig] = f1[8 (p + tig) + gid]
  const double* b_ptr = s2 + 8 * tig + gid + warp * (kTcTiles * 8);   // B_m[p' = tig][j = gid] = row (p + m + tig) of f2, column gid
  for (int64_t c = c_begin; c < c_end; c++) {
    const int64_t t0 = c * kTcChunk;
    __syncthreads();   // the previous chunk has been consumed
    for (int e = tid; e < kTcChunk; e += kTcThreads) s1[e] = f1[t0 + e];
    for (int e = tid; e < kTcF2; e += kTcThreads) s2[e] = f2[t0 + e];
    __syncthreads();
    double ring[16];   // ring[r % 16] = fragment of f2 rows r .. r + 3; step k uses r = 4k .. 4k + 15
#pragma unroll
    for (int r = 0; r < 16; r++) ring[r] = b_ptr[8 * r];
    double a = a_ptr[0];
#pragma unroll 1
    for (int k0 = 0; k0 < kTcSteps; k0 += 4) {
#pragma unroll
      for (int kk = 0; kk < 4; kk++) {
        const int k = k0 + kk;
        const double a_next = a_ptr[32 * (k + 1) & (kTcChunk - 1)];   // (wraps on the last step: value unused)
#pragma unroll
        for (int m = 0; m < 4; m++) dmma_m8n8k4(acc[m][0], acc[m][1], a, ring[(4 * kk + m) & 15]);
        // rows 4k .. 4k + 3 are done with: their slots take rows 4k + 16 .. 4k + 19, first used at the END of the next step
#pragma unroll
        for (int m = 0; m < 4; m++) ring[(4 * kk + m) & 15] = b_ptr[8 * (4 * k + 16 + m)];
#pragma unroll
        for (int m = 4; m < kTcTiles; m++) dmma_m8n8k4(acc[m][0], acc[m][1], a, ring[(4 * kk + m) & 15]);
        a = a_next;
      }
    }
  }

  // diagonal sums of every tile, in a fixed order: the tile goes through shared memory, lane d' + 7 adds its diagonal
  __syncthreads();
  double* tile = s2 + warp * 64;
  double* out = part + ((pair * gridDim.y + blockIdx.y) * tiles_pad + tau0 / 8 + warp * kTcTiles) * 16;
#pragma unroll
  for (int m = 0; m < kTcTiles; m++) {
    tile[gid * 8 + 2 * tig] = acc[m][0];
    tile[gid * 8 + 2 * tig + 1] = acc[m][1];
    __syncwarp();
    if (lane < 16) {
      const int d = lane - 7;   // j - i
      double v = 0.0;
      if (lane < 15) {
        for (int i = 0; i < 8; i++) {
          const int jj = i + d;
          if (jj >= 0 && jj < 8) v += tile[i * 8 + jj];
        }
      }
      out[m * 16 + lane] = v;
    }
    __syncwarp();
  }
}

// out[pair][tau] = (sum over splits of the two tile diagonals holding lag tau, + wrapped term) / (n - tau)
__device__ __forceinline__ double lag_sum(const double* __restrict__ part, int64_t row0, int splits, int64_t tiles_pad,
                                          int64_t lag) {
  const int64_t M = lag >> 3;
  const int d = (int)(lag & 7);
  double v = 0.0;
  for (int s = 0; s < splits; s++) {
    const double* t = part + ((row0 + s) * tiles_pad + M) * 16;
    v += t[d + 7];                  // G_M, diagonal j - i = d
    if (d > 0) v += t[16 + d - 1];  // G_{M+1}, diagonal j - i = d - 8
  }
  return v;
}

__global__ void time_corr_finish_kernel(const double* __restrict__ part, int splits, const double* __restrict__ wrap,
                                        int wrap_splits, int64_t tiles_pad, int64_t wrap_tiles_pad, int64_t n,
                                        int64_t corr_len, int64_t total, double* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (; i < total; i += step) {
    const int64_t pair = i / corr_len, tau = i - pair * corr_len;
    double v = lag_sum(part, pair * splits, splits, tiles_pad, tau);
    if (wrap != nullptr && tau > 0) v += lag_sum(wrap, pair * wrap_splits, wrap_splits, wrap_tiles_pad, corr_len - tau);
    out[i] = v / (double)(n - tau);
  }
}

// Time splits per (pair, lag block): several times more CTAs than the machine holds at once (4 per SM), so that every SM
// is refilled until the end, but at least four chunks (4096 time steps) per CTA.
int pick_splits(int64_t n_t, int64_t lag_blocks, int n_pairs, int sm_count) {
  const int64_t chunks = (n_t + kTcChunk - 1) / kTcChunk;
  int64_t per_sm = 32;
  if (const char* e = getenv("CGC_TIMECORR_CTAS_PER_SM")) per_sm = std::max<int64_t>(1, atoll(e));   // tuning hook
  const int64_t want = (int64_t)sm_count * per_sm;
  const int64_t have = lag_blocks * n_pairs;
  int64_t s = (want + have - 1) / have;
  s = std::min<int64_t>(s, std::max<int64_t>(1, chunks / 4));
  return (int)std::max<int64_t>(1, std::min<int64_t>(s, 65535));
}

// lags 0 .. n_lags - 1 need tiles 0 .. (n_lags - 1) / 8 + 1, in whole CTAs
int64_t lag_blocks_for(int64_t n_lags) { return ((n_lags - 1) / 8 + 2 + kTcWarps * kTcTiles - 1) / (kTcWarps * kTcTiles); }

}  // namespace

int64_t time_corr_series_stride(int64_t n) { return (n + kTcChunk + kTcF2 + 64 + 7) & ~(int64_t)7; }

void time_corr_plan(int64_t n, int64_t corr_len, int n_pairs, bool circular, int sm_count, int* splits, int* wrap_splits,
                    int64_t* scratch_doubles) {
  const int64_t lb = lag_blocks_for(corr_len);
  *splits = pick_splits(n, lb, n_pairs, sm_count);
  int64_t need = (int64_t)n_pairs * *splits * lb * kTcLagBlock * 2;   // 16 slots per tile of 8 lags
  *wrap_splits = 0;
  if (circular) {
    const int64_t wb = lag_blocks_for(corr_len + 1);
    *wrap_splits = pick_splits(corr_len, wb, n_pairs, sm_count);
    need += (int64_t)n_pairs * *wrap_splits * wb * kTcLagBlock * 2;
  }
  *scratch_doubles = need;
}

// d_series: n_series rows of time_corr_series_stride(n) doubles, zero from element n on.  d_first/d_second: n_pairs series
// indices.  d_out: [n_pairs][corr_len].  d_scratch: as sized by time_corr_plan.  out[p][tau] = sum_t f1[t] f2[t + tau] /
// (n - tau), t + tau taken modulo n when circular.  Returns the number of kernels launched.
int launch_time_corr(const double* d_series, int64_t n, const int32_t* d_first, const int32_t* d_second, int n_pairs,
                     int64_t corr_len, bool circular, int sm_count, double* d_scratch, double* d_out, cudaStream_t st) {
  const int64_t stride = time_corr_series_stride(n);
  const int64_t lb = lag_blocks_for(corr_len);
  const int64_t tiles_pad = lb * kTcWarps * kTcTiles;
  int splits = 1, wrap_splits = 0;
  int64_t scratch = 0;
  time_corr_plan(n, corr_len, n_pairs, circular, sm_count, &splits, &wrap_splits, &scratch);
  int launches = 0;
  double* part = d_scratch;
  time_corr_kernel<<<dim3((unsigned)lb, (unsigned)splits, (unsigned)n_pairs), kTcThreads, 0, st>>>(
      d_series, stride, d_first, d_second, 0, n, tiles_pad, part);
  launches++;
  double* wrap = nullptr;
  int64_t wrap_tiles_pad = 0;
  if (circular) {
    // wrapped terms: W[lambda] = sum_{u < corr_len} f2[u] * f1[n - corr_len + u + lambda], lambda = corr_len - tau in [1, corr_len]
    const int64_t wb = lag_blocks_for(corr_len + 1);
    wrap_tiles_pad = wb * kTcWarps * kTcTiles;
    wrap = part + (int64_t)n_pairs * splits * tiles_pad * 16;
    time_corr_kernel<<<dim3((unsigned)wb, (unsigned)wrap_splits, (unsigned)n_pairs), kTcThreads, 0, st>>>(
        d_series, stride, d_second, d_first, n - corr_len, corr_len, wrap_tiles_pad, wrap);
    launches++;
  }
  const int64_t total = (int64_t)n_pairs * corr_len;
  time_corr_finish_kernel<<<(unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count * 16), 256, 0, st>>>(
      part, splits, wrap, wrap_splits, tiles_pad, wrap_tiles_pad, n, corr_len, total, d_out);
  launches++;
  return launches;
}

}  // namespace cgc

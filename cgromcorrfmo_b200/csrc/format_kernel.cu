// K5 — the `.time` text on the device.
//
// The reference prints every dE value with `stream << float` (default precision: "%g", 6 significant digits) after the
// unit conversion float(double(kcal) * 349.7) (reference src/FMOparse.cpp:151-165).  At GPU speeds that text is 240 kB per
// C2 frame — 2.9e8 conversions per second at 11k frames/s — and formatting it on the host costs as many core-seconds as
// staging the trajectory does (the CLI was CPU-bound at 5.2k frames/s with both on 16 cores).  Here one thread formats
// one value, byte for byte what printf("%g") prints:
//   * decimal exponent from the binary one, corrected against power-of-ten thresholds;
//   * the 6 digits as round(a * 10^k) in double; the product is rounded once (2^-53), so the digits can only be wrong
//     within ~1e-10 of a rounding tie: there the tie is broken exactly, comparing M * 10^k * 2^(E+1) with 2r + 1 in
//     128-bit integers (a = M * 2^E is a float32: M < 2^24) and rounding half to even like glibc;
//   * values outside [1e-16, 1e16) and non-finite ones raise a flag and the batch is formatted by the host instead
//     (dE in cm^-1 never gets there; exact zeros print as "0" / "-0").
// Rows are laid out by a two-pass scheme: lengths per row (format_len_kernel) -> exclusive scan over the rows
// (format_scan_kernel) -> format again, scan inside the row, stage 256 values in shared memory and copy them out as
// coalesced bytes (format_write_kernel).
#include "kernels.cuh"

#include <cstdint>

namespace cgc {

namespace {

__constant__ double kFmtPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                     1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
__constant__ double kFmtThr[37] = {1e-18, 1e-17, 1e-16, 1e-15, 1e-14, 1e-13, 1e-12, 1e-11, 1e-10, 1e-9, 1e-8, 1e-7, 1e-6,
                                   1e-5,  1e-4,  1e-3,  1e-2,  1e-1,  1e0,   1e1,   1e2,   1e3,  1e4,  1e5,  1e6,
                                   1e7,   1e8,   1e9,   1e10,  1e11,  1e12,  1e13,  1e14,  1e15, 1e16, 1e17, 1e18};
__constant__ unsigned long long kFmtPow10u[20] = {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull,
                                                  100000000ull, 1000000000ull, 10000000000ull, 100000000000ull,
                                                  1000000000000ull, 10000000000000ull, 100000000000000ull,
                                                  1000000000000000ull, 10000000000000000ull, 100000000000000000ull,
                                                  1000000000000000000ull, 10000000000000000000ull};

constexpr int kFmtThreads = 256;
constexpr int kFmtMaxLen = 12;    // "-1.23456e-05", "-0.000123456"

// sign of (|f| * 10^k) - (r + 1/2), exactly.  |f| = M * 2^E with M < 2^24; -22 <= k <= 22.
__device__ int fmt_tie_cmp(float f, int k, long long r) {
  const unsigned bits = __float_as_uint(f) & 0x7fffffffu;
  const unsigned long long M = (bits & 0x7fffffu) | 0x800000u;   // normal numbers only (|f| >= 1e-16)
  const int E = (int)(bits >> 23) - 150;
  unsigned __int128 lhs = M, rhs = (unsigned __int128)(2 * r + 1);
  if (k >= 0) {
    const int k1 = k > 19 ? 19 : k;
    lhs *= kFmtPow10u[k1];
    lhs *= kFmtPow10u[k - k1];          // 10^k < 2^74, M * 10^k < 2^98
  } else {
    rhs *= kFmtPow10u[-k];              // (2r + 1) * 10^-k < 2^21 * 2^74
  }
  const int sh = E + 1;                 // compare lhs * 2^sh with rhs
  if (sh >= 0) lhs <<= sh; else rhs <<= -sh;
  return lhs > rhs ? 1 : (lhs < rhs ? -1 : 0);
}

// "%g" of f into p (at most kFmtMaxLen characters); returns the length.  `unsupported` is set when the host must do it.
__device__ int fmt_g6_dev(float f, char* p, bool& unsupported) {
  if (f == 0.0f) {
    if (__float_as_uint(f) >> 31) { p[0] = '-'; p[1] = '0'; return 2; }
    p[0] = '0';
    return 1;
  }
  const double v = (double)f;
  const double a = fabs(v);
  if (!(a >= 1e-16 && a < 1e16)) {   // also NaN / Inf
    unsupported = true;
    p[0] = '?';
    return 1;
  }
  const int e2 = (int)((__double_as_longlong(a) >> 52) & 0x7ff) - 1023;   // a = 1.m * 2^e2
  int e10 = (e2 * 78913) >> 18;                                             // floor(e2 * log10(2))
  if (a < kFmtThr[e10 + 18]) e10--;
  else if (a >= kFmtThr[e10 + 19]) e10++;
  const int k = 5 - e10;                       // a * 10^k has 6 integer digits
  const double scaled = k >= 0 ? a * kFmtPow10[k] : a / kFmtPow10[-k];
  const double fl = floor(scaled);
  const double frac = scaled - fl;
  long long r = (long long)fl;
  if (fabs(frac - 0.5) < 1e-6) {
    const int c = fmt_tie_cmp(f, k, r);
    if (c > 0 || (c == 0 && (r & 1))) r++;     // exact tie: half to even
  } else if (frac > 0.5) {
    r++;
  }
  if (r >= 1000000) { r = 100000; e10++; }
  if (r < 100000 || r >= 1000000) {            // cannot happen for a consistent exponent; let the host decide
    unsupported = true;
    p[0] = '?';
    return 1;
  }
  char d[6];
  unsigned rr = (unsigned)r;
#pragma unroll
  for (int i = 5; i >= 0; i--) { d[i] = (char)('0' + rr % 10u); rr /= 10u; }
  int nd = 6;
  while (nd > 1 && d[nd - 1] == '0') nd--;
  char* q = p;
  if (v < 0) *q++ = '-';
  if (e10 < -4 || e10 >= 6) {
    *q++ = d[0];
    if (nd > 1) {
      *q++ = '.';
      for (int i = 1; i < nd; i++) *q++ = d[i];
    }
    *q++ = 'e';
    int e = e10;
    if (e < 0) { *q++ = '-'; e = -e; } else { *q++ = '+'; }
    *q++ = (char)('0' + e / 10);
    *q++ = (char)('0' + e % 10);
  } else if (e10 >= 0) {
    const int ni = e10 + 1;   // digits before the point
    if (nd <= ni) {
      for (int i = 0; i < nd; i++) *q++ = d[i];
      for (int i = nd; i < ni; i++) *q++ = '0';
    } else {
      for (int i = 0; i < ni; i++) *q++ = d[i];
      *q++ = '.';
      for (int i = ni; i < nd; i++) *q++ = d[i];
    }
  } else {
    *q++ = '0';
    *q++ = '.';
    for (int i = 0; i < -e10 - 1; i++) *q++ = '0';
    for (int i = 0; i < nd; i++) *q++ = d[i];
  }
  return (int)(q - p);
}

__device__ __forceinline__ float fmt_value(const float* __restrict__ x, int64_t i, int to_cm) {
  const float v = x[i];
  return to_cm ? (float)((double)v * 349.7) : v;   // reference src/FMOparse.cpp:151
}

// block-wide exclusive scan of one int per thread (kFmtThreads threads); returns the exclusive prefix, *total = the sum
__device__ int block_exclusive_scan(int v, int* total) {
  __shared__ int warp_sum[kFmtThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();   // warp_sum may still be read by the previous call
  if (lane == 31) warp_sum[warp] = inc;
  __syncthreads();
  int base = 0, all = 0;
#pragma unroll
  for (int w = 0; w < kFmtThreads / 32; w++) {
    const int s = warp_sum[w];
    if (w < warp) base += s;
    all += s;
  }
  *total = all;
  return base + inc - v;
}

// one CTA per row: bytes of the formatted row, separators (ncol - 1 commas and the newline) included
__global__ void __launch_bounds__(kFmtThreads)
format_len_kernel(const float* __restrict__ x, int ncol, int to_cm, int64_t* __restrict__ row_len, int* __restrict__ flags) {
  const int64_t row = blockIdx.x;
  int mine = 0;
  bool bad = false;
  char buf[16];
  for (int j = threadIdx.x; j < ncol; j += kFmtThreads) mine += fmt_g6_dev(fmt_value(x, row * ncol + j, to_cm), buf, bad) + 1;
  int total;
  block_exclusive_scan(mine, &total);
  if (threadIdx.x == 0) row_len[row] = total;
  if (bad) atomicOr(flags, 1);
}

// in place: row_len[0..rows) -> exclusive offsets, row_len[rows] = total bytes.  One CTA.
__global__ void __launch_bounds__(1024) format_scan_kernel(int64_t* __restrict__ row_len, int64_t rows) {
  __shared__ long long warp_sum[32];
  __shared__ long long carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t base = 0; base < rows; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const long long v = i < rows ? row_len[i] : 0;
    long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sum[warp] = inc;
    __syncthreads();
    long long before = carry_s, all = 0;
    for (int w = 0; w < 32; w++) {
      if (w < warp) before += warp_sum[w];
      all += warp_sum[w];
    }
    if (i < rows) row_len[i] = before + inc - v;
    __syncthreads();
    if (threadIdx.x == 0) carry_s += all;
    __syncthreads();
  }
  if (threadIdx.x == 0) row_len[rows] = carry_s;
}

__global__ void __launch_bounds__(kFmtThreads)
format_write_kernel(const float* __restrict__ x, int ncol, int to_cm, const int64_t* __restrict__ row_off, char* __restrict__ out) {
  __shared__ char stage[kFmtThreads * (kFmtMaxLen + 1)];
  const int64_t row = blockIdx.x;
  char* dst = out + row_off[row];
  for (int j0 = 0; j0 < ncol; j0 += kFmtThreads) {
    const int j = j0 + threadIdx.x;
    char buf[16];
    int len = 0;
    bool bad = false;
    if (j < ncol) {
      len = fmt_g6_dev(fmt_value(x, row * ncol + j, to_cm), buf, bad);
      buf[len++] = (j == ncol - 1) ? '\n' : ',';
    }
    int total;
    const int off = block_exclusive_scan(len, &total);
    for (int i = 0; i < len; i++) stage[off + i] = buf[i];
    __syncthreads();
    for (int i = threadIdx.x; i < total; i += kFmtThreads) dst[i] = stage[i];   // coalesced bytes
    dst += total;
    __syncthreads();   // stage is rewritten by the next chunk
  }
}

}  // namespace

int64_t format_time_capacity(int64_t rows, int ncol) { return rows * ncol * (int64_t)(kFmtMaxLen + 1); }

void launch_format_time(const float* d_x, int64_t rows, int ncol, bool to_cm, int64_t* d_row_off, char* d_out, int* d_flags,
                        cudaStream_t st) {
  if (rows <= 0 || ncol <= 0) return;
  format_len_kernel<<<(unsigned)rows, kFmtThreads, 0, st>>>(d_x, ncol, to_cm ? 1 : 0, d_row_off, d_flags);
  format_scan_kernel<<<1, 1024, 0, st>>>(d_row_off, rows);
  format_write_kernel<<<(unsigned)rows, kFmtThreads, 0, st>>>(d_x, ncol, to_cm ? 1 : 0, d_row_off, d_out);
}

}  // namespace cgc

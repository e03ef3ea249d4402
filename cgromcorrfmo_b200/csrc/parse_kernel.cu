// K8 — `.time` / `.mtx` text back into numbers on the device (the inverse of K5).
//
// The reference's consumers read csv_out.time back as text: scripts/2014-06-csv2hdf5.py:56-65 does
// `[[float(x) for x in l.split(',')] for l in f.strip().split("\n")]` (Python float == correctly rounded strtod) and reshapes
// to (frames, 7, 357); scripts/depca/depca/dedata.py:96-120 does the same through genfromtxt.  Files written by earlier
// runs of the reference (or by this CLI without --npy) go through this kernel instead: HBM-bound byte work.
//
// Tokens have variable width, so the text is cut into 512-byte blocks, one per warp at a time (16 bytes per lane); there is
// no CTA-wide barrier anywhere:
//   pass 1  count_kernel   delimiters (',' and '\n') per block, newlines in total                      (SWAR byte tests)
//   scan    exclusive prefix over the block counts                                                     (one CTA)
//   pass 2  parse_kernel   the warp copies its block (+ a halo) into its own shared memory, lists the positions of the
//                          block's delimiters there (warp scan of the per-lane counts), and deals the tokens that follow
//                          them out round-robin, so every lane parses the same number of tokens whatever the spans hold.
//                          The delimiter's global index + 1 is the value's index.
// A token  [-+]digits[.digits][e[-+]digits]  with at most 15 significant digits and a decimal exponent within +-22 converts
// EXACTLY like strtod: D and 10^|e| are exact doubles, so D * 10^e or D / 10^e is one correctly rounded operation (Clinger's
// fast path; __dmul_rn / __ddiv_rn).  "%g" text (6 significant digits) is always of that form unless the value is outside
// 1e-17 .. 1e27 or not finite; such tokens ("nan", "inf", longer mantissas, larger exponents, anything malformed) raise a
// flag and the host re-parses that chunk with strtod.  Row shape is checked on the fly: delimiter k must be a newline
// exactly when (k + 1) is a multiple of the column count.
#include "kernels.cuh"

#include <algorithm>
#include <cstdint>

namespace cgc {

namespace {

constexpr int kPtThreads = 256;
constexpr int kPtWarps = kPtThreads / 32;
constexpr int kPtSpan = 16;                        // bytes per lane
constexpr int kPtTile = 32 * kPtSpan;              // bytes per warp iteration ("block")
constexpr int kPtHalo = 48;                        // longest token the device takes (40) + slack, a multiple of 16
constexpr int kPtMaxToken = 40;

__constant__ double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                  1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

// 0x80 in every byte of w that equals c (exact per byte, no borrow artefacts)
__device__ __forceinline__ uint32_t bytes_equal(uint32_t w, uint32_t c4) {
  const uint32_t v = w ^ c4;
  const uint32_t t = (v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
  return ~(t | v | 0x7F7F7F7Fu);
}

// bit i of the result: byte i of the 16-byte span is ',' (and/or '\n')
__device__ __forceinline__ void span_masks(const uint4& q, uint32_t& delim, uint32_t& newline) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
  delim = newline = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const uint32_t nl = bytes_equal(w[k], 0x0A0A0A0Au);
    const uint32_t cm = bytes_equal(w[k], 0x2C2C2C2Cu);
    // gather the four 0x80 flags into 4 bits: bit 7 -> 0, 15 -> 1, 23 -> 2, 31 -> 3
    const uint32_t nl4 = ((nl >> 7) & 1u) | ((nl >> 14) & 2u) | ((nl >> 21) & 4u) | ((nl >> 28) & 8u);
    const uint32_t cm4 = ((cm >> 7) & 1u) | ((cm >> 14) & 2u) | ((cm >> 21) & 4u) | ((cm >> 28) & 8u);
    newline |= nl4 << (4 * k);
    delim |= (nl4 | cm4) << (4 * k);
  }
}

__global__ void __launch_bounds__(kPtThreads)
parse_count_kernel(const uint4* __restrict__ text, int64_t n_tiles, int* __restrict__ tile_counts,
                   unsigned long long* __restrict__ totals /* [0] delimiters, [1] newlines */) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kPtWarps + (threadIdx.x >> 5), n_warps = (int64_t)gridDim.x * kPtWarps;
  unsigned long long my_delims = 0, my_newlines = 0;
  for (int64_t tile = warp0; tile < n_tiles; tile += n_warps) {
    const uint4 q = text[tile * 32 + lane];
    uint32_t d, nl;
    span_masks(q, d, nl);
    const int cnt = __reduce_add_sync(0xffffffffu, __popc(d));
    if (lane == 0) tile_counts[tile] = cnt;
    my_delims += __popc(d);
    my_newlines += __popc(nl);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    my_delims += __shfl_xor_sync(0xffffffffu, my_delims, o);
    my_newlines += __shfl_xor_sync(0xffffffffu, my_newlines, o);
  }
  if (lane == 0) {
    atomicAdd(&totals[0], my_delims);
    atomicAdd(&totals[1], my_newlines);
  }
}

// exclusive prefix of the block counts in two levels: inside groups of 1024 blocks (coalesced, one CTA per group), then
// over the group totals (one CTA)
constexpr int kPtGroup = 1024;

__global__ void __launch_bounds__(kPtGroup)
parse_scan_groups_kernel(const int* __restrict__ counts, int64_t n, int* __restrict__ local, int64_t* __restrict__ group_total) {
  __shared__ int warp_sums[kPtGroup / 32];
  const int64_t i = (int64_t)blockIdx.x * kPtGroup + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int v = i < n ? counts[i] : 0;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = warp_sums[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    warp_sums[lane] = w;   // inclusive over the warps
  }
  __syncthreads();
  const int before = warp > 0 ? warp_sums[warp - 1] : 0;
  if (i < n) local[i] = before + incl - v;
  if (threadIdx.x == kPtGroup - 1) group_total[blockIdx.x] = before + incl;
}

__global__ void __launch_bounds__(1024)
parse_scan_totals_kernel(int64_t* __restrict__ group_total, int64_t n_groups) {   // in place: totals -> exclusive offsets
  __shared__ int64_t part[1024];
  const int64_t per = (n_groups + 1023) / 1024;
  const int64_t b = per * threadIdx.x, e = min(n_groups, b + per);
  int64_t s = 0;
  for (int64_t i = b; i < e; i++) s += group_total[i];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int64_t run = 0;
    for (int i = 0; i < 1024; i++) { const int64_t v = part[i]; part[i] = run; run += v; }
  }
  __syncthreads();
  int64_t run = part[threadIdx.x];
  for (int64_t i = b; i < e; i++) { const int64_t v = group_total[i]; group_total[i] = run; run += v; }
}

enum : int { kParseNeedsHost = 1, kParseBadShape = 2 };

// strtod of the token at s (shared memory; ends at ',' '\n' or a NUL of the padding).  false: not a fast-path token.
__device__ __forceinline__ bool parse_token(const unsigned char* __restrict__ s, double& out) {
  int i = 0;
  bool neg = false;
  if (s[0] == '-') { neg = true; i = 1; }
  else if (s[0] == '+') i = 1;
  unsigned long long D = 0;
  int nd = 0, frac = 0;
  bool dot = false, any = false, ok = true;
  unsigned char c;
  for (;; i++) {
    c = s[i];
    if (c >= '0' && c <= '9') {
      any = true;
      if (D != 0 || c != '0') {
        if (nd < 15) { D = D * 10 + (c - '0'); nd++; }
        else ok = false;   // more significant digits than a double takes exactly
      }
      if (dot) frac++;
    } else if (c == '.' && !dot) {
      dot = true;
    } else {
      break;
    }
    if (i >= kPtMaxToken) { ok = false; break; }
  }
  int ex = 0;
  if (c == 'e' || c == 'E') {
    i++;
    bool eneg = false;
    if (s[i] == '-') { eneg = true; i++; }
    else if (s[i] == '+') i++;
    bool edig = false;
    for (; i < kPtMaxToken + 6; i++) {
      c = s[i];
      if (c < '0' || c > '9') break;
      edig = true;
      if (ex < 100000) ex = ex * 10 + (c - '0');
    }
    if (!edig) ok = false;
    if (eneg) ex = -ex;
  }
  c = s[i];
  if (!(c == ',' || c == '\n' || c == 0) || !any || !ok || i > kPtMaxToken) return false;
  double v = 0.0;
  if (D != 0) {
    const int e10 = ex - frac;
    if (e10 > 22 || e10 < -22) return false;
    v = e10 >= 0 ? __dmul_rn((double)D, kPow10[e10]) : __ddiv_rn((double)D, kPow10[-e10]);
  }
  out = neg ? -v : v;
  return true;
}

__global__ void __launch_bounds__(kPtThreads)
parse_kernel(const uint4* __restrict__ text, int64_t n_bytes, int64_t n_tiles, const int64_t* __restrict__ group_offsets,
             const int* __restrict__ local_offsets, int num_col, int64_t cap, double* __restrict__ out, int* __restrict__ flags,
             unsigned long long* __restrict__ first_bad_row) {
  __shared__ __align__(16) unsigned char bytes_all[kPtWarps][kPtTile + kPtHalo];
  __shared__ uint16_t pos_all[kPtWarps][kPtTile];   // positions of the block's delimiters, bit 15 = it is a newline
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* bytes = bytes_all[warp];
  uint16_t* pos = pos_all[warp];
  const int64_t warp0 = (int64_t)blockIdx.x * kPtWarps + warp, n_warps = (int64_t)gridDim.x * kPtWarps;
  int my_flags = 0;
  for (int64_t tile = warp0; tile < n_tiles; tile += n_warps) {
    __syncwarp();   // the previous block has been consumed
    const uint4 q = text[tile * 32 + lane];
    reinterpret_cast<uint4*>(bytes)[lane] = q;
    if (lane < kPtHalo / 16) reinterpret_cast<uint4*>(bytes)[32 + lane] = text[(tile + 1) * 32 + lane];
    uint32_t d, nl;
    span_masks(q, d, nl);
    const int cnt = __popc(d);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int at = incl - cnt;
    while (d) {
      const int p = __ffs(d) - 1;
      d &= d - 1;
      pos[at++] = (uint16_t)((lane * kPtSpan + p) | (((nl >> p) & 1u) << 15));
    }
    __syncwarp();   // bytes and positions are in shared memory
    const int64_t k0 = group_offsets[tile / kPtGroup] + local_offsets[tile];   // global index of the block's first delimiter
    const int64_t base = tile * kPtTile;
    if (tile == 0 && lane == 0 && n_bytes > 0 && cap > 0) {
      double v;
      if (parse_token(bytes, v)) out[0] = v;
      else my_flags |= kParseNeedsHost;
    }
    for (int j = lane; j < total; j += 32) {
      const uint32_t e = pos[j];
      const int p = e & 0x7fff;
      const int64_t k = k0 + j;
      if ((bool)(e >> 15) != ((k + 1) % num_col == 0)) {
        my_flags |= kParseBadShape;
        atomicMin(first_bad_row, (unsigned long long)(k / num_col));
      }
      if (base + p + 1 < n_bytes && k + 1 < cap) {
        double v;
        if (parse_token(bytes + p + 1, v)) out[k + 1] = v;
        else my_flags |= kParseNeedsHost;
      }
    }
  }
  if (my_flags) atomicOr(flags, my_flags);
}

}  // namespace

// bytes of device text buffer needed for n_bytes of text: whole 512-byte blocks, one more for the halo of the last, zero-filled
int64_t parse_text_padded_bytes(int64_t n_bytes) { return ((n_bytes + kPtTile - 1) / kPtTile + 1) * kPtTile; }
int64_t parse_text_tiles(int64_t n_bytes) { return (n_bytes + kPtTile - 1) / kPtTile; }
int64_t parse_text_scan_bytes(int64_t n_bytes) {
  const int64_t tiles = parse_text_tiles(n_bytes);
  return 8 * ((tiles + kPtGroup - 1) / kPtGroup) + 4 * tiles + 16;
}

// d_text: parse_text_padded_bytes(n_bytes) bytes, zero behind the text.  d_tile_counts: tiles ints; d_totals: 2 x uint64
// (zeroed here): delimiters, newlines.
void launch_parse_count(const char* d_text, int64_t n_bytes, int* d_tile_counts, unsigned long long* d_totals, int sm_count,
                        cudaStream_t st) {
  const int64_t tiles = parse_text_tiles(n_bytes);
  cudaMemsetAsync(d_totals, 0, 2 * sizeof(unsigned long long), st);
  if (tiles == 0) return;
  const unsigned grid = (unsigned)std::min<int64_t>((tiles + kPtWarps - 1) / kPtWarps, (int64_t)sm_count * 8);
  parse_count_kernel<<<grid, kPtThreads, 0, st>>>(reinterpret_cast<const uint4*>(d_text), tiles, d_tile_counts, d_totals);
}

// d_tile_offsets: parse_text_scan_bytes(n_bytes) bytes of scratch.  d_out: cap doubles; value i = token i of the text.  d_flags: int, d_bad_row: uint64
// (both set up here).  flags & 1: a token needs the host's strtod; flags & 2: a row does not have num_col values
// (*d_bad_row = the first such row).
void launch_parse(const char* d_text, int64_t n_bytes, const int* d_tile_counts, int64_t* d_tile_offsets, int num_col,
                  int64_t cap, double* d_out, int* d_flags, unsigned long long* d_bad_row, int sm_count, cudaStream_t st) {
  const int64_t tiles = parse_text_tiles(n_bytes);
  cudaMemsetAsync(d_flags, 0, sizeof(int), st);
  cudaMemsetAsync(d_bad_row, 0xff, sizeof(unsigned long long), st);
  if (tiles == 0) return;
  const int64_t groups = (tiles + kPtGroup - 1) / kPtGroup;
  int64_t* group_off = d_tile_offsets;
  int* local = reinterpret_cast<int*>(d_tile_offsets + groups);
  parse_scan_groups_kernel<<<(unsigned)groups, kPtGroup, 0, st>>>(d_tile_counts, tiles, local, group_off);
  parse_scan_totals_kernel<<<1, 1024, 0, st>>>(group_off, groups);
  const unsigned grid = (unsigned)std::min<int64_t>((tiles + kPtWarps - 1) / kPtWarps, (int64_t)sm_count * 8);
  parse_kernel<<<grid, kPtThreads, 0, st>>>(reinterpret_cast<const uint4*>(d_text), n_bytes, tiles, group_off, local, num_col, cap,
                                            d_out, d_flags, d_bad_row);
}

}  // namespace cgc

// Device kernels of the traj2dEcsv hot path, written for sm_100a (B200).
//
//   K1  decode_kernel     fixed-column .gro text -> float32 coordinates (blocks of 8 atoms, SoA) + per-frame chromophore site blocks
//                         (replaces the per-frame text parsing of ReadOneTimeGrom2Atom, reference
//                         src/grom2atomFMO.cpp:396-455, and the per-site charge substitution of AtomDataLookup_v2, :574-609)
//   K2  cdc_kernel        Coulomb sum dQ(site atoms) x q(all other atoms), binned per residue group
//                         (replaces ComputeCDC_v1, reference src/grom2atomFMO.cpp:625-739)
//   K2b finish_kernel     FP64 bins -> float32 dE rows (the reference's cdc_kcal vector)
//   K3  cov_* kernels     n, sum(x-c), sum (x-c)(x-c)^T per site on FP64 tensor cores (DMMA)
//                         (replaces the Correlate block, reference src/FMOparse.cpp:168-178, 188-206)
#include "kernels.cuh"
#include "ptx_helpers.cuh"
#include "cdc_kernel.cuh"

#include <cuda.h>   // CUtensorMap and its enums; cuTensorMapEncodeTiled itself is bound through cudaGetDriverEntryPoint

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

namespace cgc {

// cudaFuncSetAttribute is per device: remember, per device, the largest dynamic shared memory size already granted.
template <typename K>
static void ensure_dynamic_smem(K kernel, size_t bytes, size_t* granted /* [16] */) {
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 15;
  if (bytes > 48 * 1024 && bytes > granted[dev]) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    granted[dev] = bytes;
  }
}

// cudaOccupancyMaxActiveBlocksPerMultiprocessor and the device attributes cost tens of microseconds per query: the
// streaming loop launches every kernel once per batch, so the answers are remembered per (device, kernel, shared memory).
static int device_index() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev & 15;
}
static int cached_sm_count() {
  static int count[16] = {};
  const int dev = device_index();
  if (count[dev] == 0) {
    int v = 148;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    count[dev] = v;
  }
  return count[dev];
}
template <typename K>
static int cached_occupancy(K kernel, int threads, size_t smem, int* slots /* [16][8][2] */) {
  int* mine = slots + device_index() * 16;
  for (int i = 0; i < 8; i++) {
    if (mine[2 * i + 1] != 0 && mine[2 * i] == (int)smem) return mine[2 * i + 1];
  }
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
  if (per_sm < 1) per_sm = 1;
  for (int i = 0; i < 8; i++) {
    if (mine[2 * i + 1] == 0) {
      mine[2 * i] = (int)smem;
      mine[2 * i + 1] = per_sm;
      break;
    }
  }
  return per_sm;
}

// The neutral padding entry of a site block: dQ = 0 at a far-away point (stored negated), so it adds exactly 0.
#define CGC_PAD_NEG_COORD 1.0e6f

// ---------------------------------------------------------------------------------------------------------------------
// K1  decode
// ---------------------------------------------------------------------------------------------------------------------
// One CTA stages kDecodeThreads atom lines (a contiguous byte range) into shared memory with one bulk TMA copy, then
// each thread parses the three coordinate fields of its line.  "%W.Df" text -> integer N -> N / 10^D rounded once
// to float32, which equals (float)atof(text) (reference src/grom2atomFMO.cpp:448-450) for every N < 2^24
// (no double rounding is possible); wider fields take the double-precision division.
__device__ __forceinline__ float parse_field(const unsigned char* s, int w, int ndec, float p10f, double p10d, bool wide,
                                             int& err) {
  unsigned long long v = 0;
  bool neg = false;
#pragma unroll 4
  for (int k = 0; k < w; k++) {
    unsigned int c = s[k];
    unsigned int d = c - '0';
    if (d <= 9u) v = v * 10ull + d;
    else if (c == '-') neg = true;
    else if (c != ' ' && c != '.' && c != '+') err |= kDecBadChar;
  }
  if (s[w - 1 - ndec] != '.') err |= kDecBadPoint;
  float r;
  if (!wide) r = __fdiv_rn((float)(unsigned int)v, p10f);
  else r = (float)__ddiv_rn((double)v, p10d);
  return neg ? -r : r;
}

// Fast path for the standard GROMACS layout "%8.3f": the 24 coordinate bytes are fetched as aligned 32-bit words,
// realigned with PRMT, and each field is converted with SWAR arithmetic.  A legal field is blanks, an optional '-',
// up to four digits (word lo), then '.', three digits (word hi): digit bytes are 0x3X, ' ' and '-' are 0x20 / 0x2D.
// ~25 instructions per field for the value + ~10 for the validation (first version, byte loop: ~250 per field; ncu
// showed the kernel issue-bound).
__device__ __forceinline__ unsigned swar4(unsigned d) {   // digit values c0 c1 c2 c3 (c0 lowest byte, most significant)
  const unsigned p = (d * 10u + (d >> 8)) & 0x00FF00FFu;  // (c0*10+c1) | (c2*10+c3) << 16
  return (p * 0x00640001u) >> 16;                         // 100*(c0*10+c1) + (c2*10+c3): no carries, both halves <= 99
}
// N / 1000 rounded once to float32 without the division sequence: q = N*r, q += (N - q*1000)*r with r = RN(1/1000).
// Equal to __fdiv_rn((float)N, 1000.f) — and to the reference's (float)atof — for EVERY integer N < 2^24
// (checked exhaustively: oracle/cdc_oracle.c cgo_div1000_mismatches, tests/test_oracle.py).
__device__ __forceinline__ float div1000(float n) {
  const float r = 1.0f / 1000.0f;
  const float q = __fmul_rn(n, r);
  return __fmaf_rn(__fmaf_rn(-q, 1000.0f, n), r, q);
}
__device__ __forceinline__ float parse_field_8_3(unsigned lo, unsigned hi, unsigned& bad) {
  const unsigned isd = lo & 0x10101010u;           // bit 4: a digit byte
  const unsigned m = isd - (isd >> 4);             // 0x0F under every digit byte of lo
  const unsigned t = lo & 0x0F0F0F0Fu;
  const unsigned dlo = t & m;                      // digit values; blanks and the sign -> 0
  const unsigned y = t & ~m;                       // low nibbles of the non-digit bytes: 0x0 (' ') or 0xD ('-')
  const unsigned dhi = hi & 0x0F0F0F00u;           // 0, c5, c6, c7
  const unsigned n = swar4(dlo) * 1000u + swar4(dhi);
  // validation: lo bytes 0x2X/0x3X with digits <= 9 and non-digits ' ' or '-' (a few other 0x2X bytes slip through:
  // those with low nibble 1, 4, 5, 8, 9, C); hi exactly '.', digit, digit, digit
  bad |= ((lo & 0xE0E0E0E0u) ^ 0x20202020u) | ((dlo + 0x06060606u) & 0x10101010u) | (y & 0x02020202u) |
         ((hi & 0xF0F0F0FFu) ^ 0x3030302Eu) | ((dhi + 0x06060600u) & 0x10101000u);
  const float r = div1000((float)n);
  return y ? -r : r;                               // sign applied last, so "-0.000" keeps its sign bit
}

// Persistent CTAs: a CTA strides over the chunks of kDecodeThreads lines; the bulk copy of its next chunk is in flight
// (second shared-memory buffer) while the current one is parsed, so the HBM read stream does not stop while a CTA computes.
__host__ __device__ inline uint32_t decode_buf_stride(int line_bytes) {
  return ((uint32_t)kDecodeThreads * line_bytes + 32u + 127u) & ~127u;
}

// kMode: kDecText generic "%W.Df" text, kDecText83 the "%8.3f" fast path, kDecTrr32 / kDecTrr64 big-endian binary
// coordinates of a .trr frame (a "line" is then the 12 or 24 bytes of one atom; nothing to validate).
enum : int { kDecText = 0, kDecText83 = 1, kDecTrr32 = 2, kDecTrr64 = 3 };

template <int kMode>
__global__ void __launch_bounds__(kDecodeThreads)
decode_kernel(const char* __restrict__ text, int64_t text_bytes, const int64_t* __restrict__ atoms_off, int n_atoms,
              int n_pad, int line_bytes, int x_col, int field_w, int ndec,
              const int32_t* __restrict__ slot, const float* __restrict__ slot_dq, int sites_stride /* float4 per frame */,
              float* __restrict__ xyz8, float4* __restrict__ sites, int* __restrict__ err_flags, int chunks_per_frame,
              int n_chunks) {
  extern __shared__ __align__(128) unsigned char sm_raw[];
  __shared__ __align__(8) uint64_t bar[2];
  const uint32_t buf_stride = decode_buf_stride(line_bytes);
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();

  // (frame, chunk in frame) -> first atom, lines present, source range rounded out to 16 bytes
  struct Chunk { int f, a0, n_here; uint32_t head, bytes; int64_t src_al; bool tail_open; };
  auto locate = [&](int f, int cf) {
    Chunk k;
    k.f = f;
    k.a0 = cf * kDecodeThreads;
    int n_here = n_atoms - k.a0;
    k.n_here = n_here < 0 ? 0 : (n_here > kDecodeThreads ? kDecodeThreads : n_here);
    k.head = 0; k.bytes = 0; k.src_al = 0; k.tail_open = false;
    if (k.n_here > 0) {
      const int64_t src = atoms_off[f] + (int64_t)k.a0 * line_bytes;
      k.src_al = src & ~(int64_t)15;
      k.head = (uint32_t)(src - k.src_al);
      const int64_t end = src + (int64_t)k.n_here * line_bytes;
      k.tail_open = end - 1 >= text_bytes;   // the very last line of the text may lack its newline
      int64_t bytes = (int64_t)k.head + (int64_t)k.n_here * line_bytes;
      bytes = (bytes + 15) & ~(int64_t)15;
      const int64_t lim = ((text_bytes + 15) & ~(int64_t)15) - k.src_al;  // never read past the (16-byte padded) text
      k.bytes = (uint32_t)(bytes > lim ? lim : bytes);
    }
    return k;
  };
  auto issue = [&](const Chunk& k, int b) {   // thread 0
    if (k.n_here > 0) {
      mbar_expect_tx(&bar[b], k.bytes);
      tma_bulk_g2s(sm_raw + b * buf_stride, text + k.src_al, k.bytes, &bar[b]);
    }
  };
  // the chunk index advances by gridDim.x per trip: keep (frame, chunk in frame) incrementally instead of dividing
  const int step_f = (int)gridDim.x / chunks_per_frame, step_c = (int)gridDim.x - step_f * chunks_per_frame;
  auto advance = [&](int& f, int& cf) {
    f += step_f;
    cf += step_c;
    if (cf >= chunks_per_frame) { cf -= chunks_per_frame; f++; }
  };

  int c = blockIdx.x;
  if (c >= n_chunks) return;
  int f_cur = c / chunks_per_frame, cf_cur = c - f_cur * chunks_per_frame;
  Chunk k = locate(f_cur, cf_cur);
  if (threadIdx.x == 0) issue(k, 0);
  uint32_t phase_bits = 0;
  for (int i = 0; c < n_chunks; c += gridDim.x, i++) {
    const int b = i & 1;
    // the other buffer was last read one iteration ago; the __syncthreads that ended that iteration orders those
    // reads before this copy
    int f_next = f_cur, cf_next = cf_cur;
    advance(f_next, cf_next);
    Chunk k_next = k;
    const bool has_next = c + (int)gridDim.x < n_chunks;
    if (has_next) {
      k_next = locate(f_next, cf_next);
      if (threadIdx.x == 0) issue(k_next, b ^ 1);
    }
    if (k.n_here > 0) {
      mbar_wait(&bar[b], (phase_bits >> b) & 1u);
      phase_bits ^= 1u << b;
    }
    const int f = k.f;
    const int a = k.a0 + threadIdx.x;
    if (a < n_pad) {
      float3 out;
      if (a < n_atoms) {
        const unsigned char* line = sm_raw + b * buf_stride + k.head + threadIdx.x * line_bytes;
        int err = 0;
        if constexpr (kMode == kDecTrr32) {
          const unsigned* wp = reinterpret_cast<const unsigned*>(line);   // 4-byte aligned: every .trr field is
          out.x = __uint_as_float(__byte_perm(wp[0], 0, 0x0123));
          out.y = __uint_as_float(__byte_perm(wp[1], 0, 0x0123));
          out.z = __uint_as_float(__byte_perm(wp[2], 0, 0x0123));
        } else if constexpr (kMode == kDecTrr64) {
          const unsigned* wp = reinterpret_cast<const unsigned*>(line);
          float v[3];
#pragma unroll
          for (int d = 0; d < 3; d++) {
            const unsigned hi = __byte_perm(wp[2 * d], 0, 0x0123), lo = __byte_perm(wp[2 * d + 1], 0, 0x0123);
            v[d] = (float)__hiloint2double((int)hi, (int)lo);   // rounded once, like (float)atof of the text would be
          }
          out.x = v[0]; out.y = v[1]; out.z = v[2];
        } else if constexpr (kMode == kDecText83) {
          const unsigned char* fld = line + x_col;
          const unsigned mis = (unsigned)(smem_addr(fld) & 3u);
          const unsigned* wp = reinterpret_cast<const unsigned*>(fld - mis);
          unsigned w[7];
#pragma unroll
          for (int j = 0; j < 7; j++) w[j] = wp[j];
          const unsigned sel = 0x3210u + 0x1111u * mis;
          unsigned u[6];
#pragma unroll
          for (int j = 0; j < 6; j++) u[j] = __byte_perm(w[j], w[j + 1], sel);
          unsigned bad = 0;
          out.x = parse_field_8_3(u[0], u[1], bad);
          out.y = parse_field_8_3(u[2], u[3], bad);
          out.z = parse_field_8_3(u[4], u[5], bad);
          if (bad) {  // rare: say which rule broke
            const bool point = (u[1] & 0xFFu) != 0x2Eu || (u[3] & 0xFFu) != 0x2Eu || (u[5] & 0xFFu) != 0x2Eu;
            err |= point ? kDecBadPoint : kDecBadChar;
          }
        } else {
          float p10f = 1.f;
          double p10d = 1.0;
          for (int q = 0; q < ndec; q++) { p10f *= 10.f; p10d *= 10.0; }
          const bool wide = field_w > 8;  // more than 7 digits could exceed 2^24
          out.x = parse_field(line + x_col, field_w, ndec, p10f, p10d, wide, err);
          out.y = parse_field(line + x_col + field_w, field_w, ndec, p10f, p10d, wide, err);
          out.z = parse_field(line + x_col + 2 * field_w, field_w, ndec, p10f, p10d, wide, err);
        }
        if constexpr (kMode == kDecText || kMode == kDecText83) {
          const bool last_unterminated = k.tail_open && (int)threadIdx.x == k.n_here - 1;
          if (!last_unterminated && line[line_bytes - 1] != '\n') err |= kDecBadNewline;
          if (err) atomicOr(err_flags, err);
        }
        const int sl = slot[a];
        if (sl >= 0) {
          // site block entry for the pair loop: negated coordinates (so dx = x + (-xc) is one FADD2) and dQ
          sites[(size_t)f * sites_stride + sl] = make_float4(-out.x, -out.y, -out.z, slot_dq[sl]);
        }
      } else {
        out = make_float3(0.f, 0.f, 0.f);  // tile padding: its kq is 0 and its column is -1
      }
      // blocks of 8 atoms, SoA inside a block: [x0..x7][y0..y7][z0..z7] — the pair kernel loads these as packed f32x2
      float* dst = xyz8 + ((size_t)f * n_pad + (size_t)(a & ~7)) * 3 + (a & 7);
      dst[0] = out.x;
      dst[8] = out.y;
      dst[16] = out.z;
    }
    __syncthreads();  // everyone is done with buffer b before the copy two chunks ahead overwrites it
    k = k_next;
    f_cur = f_next;
    cf_cur = cf_next;
  }
}

// xyz8 blocks -> plain [frame][atom][3] (only used by the decode-parity entry point)
__global__ void unblock_kernel(const float* __restrict__ xyz8, float* __restrict__ xyz, int n_atoms, int n_pad) {
  const int f = blockIdx.y;
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_atoms) return;
  const float* src = xyz8 + ((size_t)f * n_pad + (size_t)(a & ~7)) * 3 + (a & 7);
  float* dst = xyz + ((size_t)f * n_atoms + a) * 3;
  dst[0] = src[0];
  dst[1] = src[8];
  dst[2] = src[16];
}

void launch_unblock(const DeviceSystem& sys, const float* xyz8, float* xyz, int n_frames, cudaStream_t st) {
  dim3 grid((sys.n_atoms + 255) / 256, n_frames);
  unblock_kernel<<<grid, 256, 0, st>>>(xyz8, xyz, sys.n_atoms, sys.n_pad);
}

__global__ void init_sites_kernel(float4* sites, int64_t n_entries) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_entries) return;
  sites[i] = make_float4(CGC_PAD_NEG_COORD, CGC_PAD_NEG_COORD, CGC_PAD_NEG_COORD, 0.f);
}

void launch_init_sites(const DeviceSystem& sys, float4* sites, int n_frames, cudaStream_t st) {
  int64_t n = (int64_t)n_frames * sys.n_sites * sys.site_cap;
  if (n == 0) return;
  init_sites_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sites, n);
}

void launch_decode(const DeviceSystem& sys, const char* text, int64_t text_bytes, const int64_t* atoms_off,
                   int n_frames, float* xyz8, float4* sites, int* err_flags, cudaStream_t st) {
  const int chunks_per_frame = (sys.n_pad + kDecodeThreads - 1) / kDecodeThreads;
  const long long n_chunks = (long long)chunks_per_frame * n_frames;
  if (n_chunks == 0) return;
  const size_t smem = 2 * (size_t)decode_buf_stride(sys.line_bytes);
  const int mode = sys.traj_mode == 1 ? kDecTrr32 : sys.traj_mode == 2 ? kDecTrr64
                   : (sys.field_w == 8 && sys.ndec == 3) ? kDecText83 : kDecText;
  auto kern = mode == kDecTrr32 ? decode_kernel<kDecTrr32> : mode == kDecTrr64 ? decode_kernel<kDecTrr64>
              : mode == kDecText83 ? decode_kernel<kDecText83> : decode_kernel<kDecText>;
  static size_t granted[4][16] = {};
  ensure_dynamic_smem(kern, smem, granted[mode]);
  static int occ[4][16 * 16] = {};
  const int sm_count = cached_sm_count();
  const int per_sm = cached_occupancy(kern, kDecodeThreads, smem, occ[mode]);
  const long long grid = std::min<long long>(n_chunks, (long long)sm_count * per_sm);
  kern<<<(unsigned)grid, kDecodeThreads, smem, st>>>(text, text_bytes, atoms_off, sys.n_atoms, sys.n_pad, sys.line_bytes,
                                                     sys.x_col, sys.field_w, sys.ndec, sys.slot, sys.slot_dq,
                                                     sys.n_sites * sys.site_cap, xyz8, sites, err_flags,
                                                     chunks_per_frame, (int)n_chunks);
}

// ---------------------------------------------------------------------------------------------------------------------
// K2  CDC pair kernel: see cdc_kernel.cuh
// ---------------------------------------------------------------------------------------------------------------------
// Sites per work unit (chunk).  A unit of cs sites costs about cs + 0.3 site-times (ring hand-over, pipeline fill and
// drain); a launch takes ceil(units / CTAs) of them, and a launch with fewer units than resident CTAs leaves SMs with one
// CTA (8 warps), which runs the pair loop about 15 % slower.  Pick the cs <= 8 that minimises that estimate: 8 for any
// launch of a few frames or more (C2: 8 frames = 1176 units over 296 CTAs, 0.7 % tail), 2 for a single frame.
static int cdc_chunk_sites(const DeviceSystem& sys, int n_frames, int resident_ctas) {
  const long long items = (long long)(sys.n_pad / kTileAtoms) * n_frames;
  const int cs_max = std::min(sys.n_sites, kCdcMaxChunkSites);
  int best = cs_max;
  double best_t = 1e300;
  for (int cs = cs_max; cs >= 1; cs--) {
    const long long units = items * ((sys.n_sites + cs - 1) / cs);
    const double unit = cs + 0.3;
    const double t = units >= resident_ctas ? (double)((units + resident_ctas - 1) / resident_ctas) * unit : 1.15 * unit;
    if (t < best_t * 0.999) { best_t = t; best = cs; }
  }
  return best;
}

size_t cdc_smem_bytes(const DeviceSystem& sys) {
  return cdc_smem_layout(std::min(sys.n_sites, kCdcMaxChunkSites), sys.site_cap, nullptr);
}

// dQ atoms per site as the kernel wants them: at least 2 (the carried pipeline peels two entries), and P - 2 a multiple
// of 7 so that the unrolled main loop covers every entry (pads are neutral: dQ = 0 at a far-away point).
int cdc_pad_site_cap(int p) {
  if (p < 2) p = 2;
  return 2 + (p - 2 + 6) / 7 * 7;
}

template <int kUnroll>
static void launch_cdc_t(const DeviceSystem& sys, const float* xyz8, const float4* sites, double* acc, int n_frames,
                         int sm_count, cudaStream_t st) {
  const int tiles = sys.n_pad / kTileAtoms;
  auto kern = cdc_kernel<kCdcMinBlocks, kUnroll>;
  static size_t granted[16] = {};
  ensure_dynamic_smem(kern, 200 * 1024, granted);
  static int occ[16 * 16] = {};
  int per_sm = cached_occupancy(kern, kTileThreads, cdc_smem_bytes(sys), occ);
  const int chunk_sites = cdc_chunk_sites(sys, n_frames, sm_count * per_sm);
  const int chunks_per_item = (sys.n_sites + chunk_sites - 1) / chunk_sites;
  const size_t smem = cdc_smem_layout(chunk_sites, sys.site_cap, nullptr);
  per_sm = cached_occupancy(kern, kTileThreads, smem, occ);
  const long long n_units = (long long)tiles * n_frames * chunks_per_item;
  long long grid = (long long)sm_count * per_sm;
  if (grid > n_units) grid = n_units;
  kern<<<(unsigned)grid, kTileThreads, smem, st>>>(xyz8, sites, sys.kq, sys.col, acc, sys.n_pad, tiles, n_units,
                                                   sys.n_sites, sys.site_cap, sys.num_tot, chunk_sites, chunks_per_item);
}

void launch_cdc(const DeviceSystem& sys, const float* xyz8, const float4* sites, double* acc, int n_frames,
                int sm_count, cudaStream_t st) {
  if (sys.n_pad / kTileAtoms * n_frames == 0 || sys.n_sites == 0) return;
  // 42 site atoms per trip when that divides P - 2 (the 44 dQ atoms of a BCL), else 7 (cdc_pad_site_cap)
  if ((sys.site_cap - 2) % 42 == 0) launch_cdc_t<42>(sys, xyz8, sites, acc, n_frames, sm_count, st);
  else launch_cdc_t<7>(sys, xyz8, sites, acc, n_frames, sm_count, st);
}

// ---------------------------------------------------------------------------------------------------------------------
// K2b  finish: FP64 bins -> float32 dE (the reference's cdc_kcal), bins re-zeroed for the next batch
// ---------------------------------------------------------------------------------------------------------------------
__global__ void finish_kernel(double* __restrict__ acc, float* __restrict__ dE, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    dE[i] = (float)acc[i];
    acc[i] = 0.0;
  }
}

void launch_finish(double* acc, float* dE, int64_t n, cudaStream_t st) {
  if (n == 0) return;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  finish_kernel<<<(unsigned)blocks, 256, 0, st>>>(acc, dE, n);
}

// ---------------------------------------------------------------------------------------------------------------------
// K3  Correlate on FP64 tensor cores
// ---------------------------------------------------------------------------------------------------------------------
// buf = [ n | c (S*G) | sum (S*G) | sumsq (S*G*G) ], all double.  d_t = x_t - c.
// sumsq_s += D_s^T D_s with D_s the (frames x G) matrix of one site: a SYRK.  Only elements whose 64-block row is <= their
// 64-block column are computed and stored (kCovBlk; finalize mirrors them).
//
// cov_centre_kernel writes the centred rows D = x - c in FP64 into a scratch matrix [frame][site][Gp], Gp = G rounded up to
// 16 with zero pads, and partial column sums that cov_sum_finish_kernel adds (with n) in a fixed order; cov_compat_kernel
// keeps the reference's two float32 running sums when the literal .mtx is wanted.
// cov_syrk_kernel is a persistent kernel, ONE CTA of 8 warps per SM.  Work unit = one 128 x 128 tile of one site's sumsq, or
// one 128 x 64 half of it; the host deals the units out longest-first (LPT) into one list per CTA, the cheapest tiles cut in
// halves so that the tail is short.  For every 32 frames the two operands — 128 (+4) columns of D each — arrive in shared
// memory as TWO TMA tiled copies (cp.async.bulk.tensor.3d, SASS UTMALDG; box = 132 columns x 1 site x 32 frames: rows of
// 1056 contiguous bytes) of a tensor map over D; frames beyond the batch and columns beyond Gp are zero-filled by the TMA
// unit, so the loop has no bounds.  The 132-column box makes the shared-memory row pitch 1056 B = 264 words == 8 (mod 32),
// which puts the four frames of a DMMA fragment load into four different bank groups: every LDS.64 is conflict-free without
// any swizzle.  (History, all at 19-21 TFLOP/s: 64 x 64 tiles staged by cp.async; 128 x 64 and 128 x 128 tiles fed by
// sixteen 16-column boxes with 128-byte swizzle per stage — ncu put 26 % of the stall samples on the full-barrier wait with
// L2 at 6 %: 256 row requests of 128 B per stage is more than the TMA unit turns around.)  A ring of 3 stages with a full /
// empty mbarrier pair per stage runs ahead across tile boundaries; one elected thread issues the copies, no CTA-wide
// barrier exists after start-up.  16 warps: warp w owns rows 32 (w & 3) .. +32 and columns 32 (w >> 2) .. +32 of the tile = 4 x 4
// DMMA m8n8k4 tiles (32 FP64 accumulators per thread): per k-step of 4 frames 8 shared-memory loads feed 16 DMMAs, and
// every sub-partition holds four warps to hide them behind (with 8 warps of 32 x 64 ncu showed the DMMA pipe 78 % busy: one
// or two warps per sub-partition do not cover their own LDS latency).  Warps whose piece lies under the diagonal 64-block,
// beyond G, or in the half the unit does not cover skip the math, and a warp whose columns end early (the last column
// block) runs a narrow instantiation of the k-step (1 or 2 of the 4 column tiles); the four warps of a sub-partition are
// the four column quarters of one row strip, so such savings shorten the unit.  The accumulators go back as fire-and-forget FP64 reductions (RED.E.ADD.F64): every element is touched by exactly one
// thread per launch and launches are stream-ordered, so the result is deterministic.
constexpr int kCovBlk = 64;        // granularity of the stored upper triangle
constexpr int kSyrkTile = 128;     // tile edge
// measured on C5 (TFLOP/s): 8 frames x 8 stages 25.7, 16 x 6 27.8, 16 x 5 27.7, 16 x 4 28.4, 24 x 4 28.6, 32 x 3 28.7, 32 x 2 28.9,
// 48 x 2 28.5 (profiles/r02_k3_variants.log): fewer, larger stages — fewer barrier round trips per DMMA
#ifndef CGC_SYRK_K
#define CGC_SYRK_K 32
#endif
#ifndef CGC_SYRK_STAGES
#define CGC_SYRK_STAGES 3
#endif
constexpr int kSyrkK = CGC_SYRK_K;             // frames per stage (tuning builds override the two macros)
constexpr int kSyrkStages = CGC_SYRK_STAGES;
constexpr int kSyrkBoxCols = 132;  // columns per TMA box: the tile's 128 + 4, for a bank-conflict-free row pitch
constexpr int kSyrkPitch = kSyrkBoxCols * 8;                                      // 1056 B
constexpr int kSyrkOperandBytes = kSyrkPitch * kSyrkK;                            // 16896
constexpr int kSyrkStageBytes = 2 * kSyrkOperandBytes;                            // 33792
constexpr int kSyrkThreads = 512;   // 16 warps: 4 row strips x 4 column quarters of 32 x 32
constexpr size_t kSyrkSmem = (size_t)kSyrkStages * kSyrkStageBytes + 128;         // + room to align the ring to 128 B
constexpr int kSyrkColPad = 16;    // D's columns are padded to a multiple of this (16-byte aligned rows for the tensor map)

int64_t cov_padded_cols(int G) { return ((int64_t)G + kSyrkColPad - 1) / kSyrkColPad * kSyrkColPad; }
constexpr int kCentreChunksMax = 16;
int64_t cov_centred_doubles(int n_frames, int S, int G) {   // D, then the partial column sums of cov_centre_kernel
  return (int64_t)n_frames * S * cov_padded_cols(G) + (int64_t)kCentreChunksMax * S * G;
}

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__global__ void cov_set_shift_kernel(const float* __restrict__ row0, double* __restrict__ buf, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) buf[1 + i] = (double)row0[i];
}

void launch_cov_set_shift(const float* dE_row0, double* buf, int S, int G, cudaStream_t st) {
  int n = S * G;
  cov_set_shift_kernel<<<(n + 255) / 256, 256, 0, st>>>(dE_row0, buf, n);
}

// D = x - c in FP64, [frame][site][Gp] with zero pads, and the column sums of D: thread (ip, y) walks the frames y, y + Y, ...
// of column ip and leaves its partial sum in part[y][column]; cov_sum_finish_kernel adds the Y partials in order.
__global__ void cov_centre_kernel(const float* __restrict__ dE, int n_frames, const double* __restrict__ buf, int S, int G,
                                  int Gp, double* __restrict__ centred, double* __restrict__ part) {
  const int ip = blockIdx.x * blockDim.x + threadIdx.x;   // index into the padded row
  const int SGp = S * Gp, SG = S * G;
  if (ip >= SGp) return;
  const int s = ip / Gp, g = ip - s * Gp;
  if (g >= G) {
    for (int t = blockIdx.y; t < n_frames; t += gridDim.y) centred[(size_t)t * SGp + ip] = 0.0;
    return;
  }
  const int i = s * G + g;
  const double c = buf[1 + i];
  double sum = 0.0;
  for (int t = blockIdx.y; t < n_frames; t += gridDim.y) {
    const double d = (double)dE[(size_t)t * SG + i] - c;
    centred[(size_t)t * SGp + ip] = d;
    sum += d;
  }
  part[(size_t)blockIdx.y * SG + i] = sum;
}

__global__ void cov_sum_finish_kernel(const double* __restrict__ part, int n_part, int n_frames, double* __restrict__ buf, int SG) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) buf[0] += (double)n_frames;
  if (i >= SG) return;
  double sum = 0.0;
  for (int y = 0; y < n_part; y++) sum += part[(size_t)y * SG + i];
  buf[1 + SG + i] += sum;
}

// The two float32 running sums the reference's literal .mtx arithmetic needs (src/FMOparse.cpp:173-175 with the zero stride
// of :193,202): sum_t x_i and sum_t x_0 x_j, added frame by frame in order like the reference does (the loads of 8 frames
// are issued together; the adds keep their order).  Only launched when the "mtx_compat" option is on.
__global__ void cov_compat_kernel(const float* __restrict__ dE, int n_frames, int SG, int G, float* __restrict__ compat /* [2][SG] */) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= SG) return;
  const int i0 = (i / G) * G;  // column 0 of this site
  float m32 = compat[i], p32 = compat[SG + i];
  int t = 0;
  for (; t + 8 <= n_frames; t += 8) {
    float x[8], x0[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      x[u] = dE[(size_t)(t + u) * SG + i];
      x0[u] = dE[(size_t)(t + u) * SG + i0];
    }
#pragma unroll
    for (int u = 0; u < 8; u++) {
      m32 += x[u];
      p32 += __fmul_rn(x0[u], x[u]);
    }
  }
  for (; t < n_frames; t++) {
    const float x = dE[(size_t)t * SG + i];
    m32 += x;
    p32 += __fmul_rn(dE[(size_t)t * SG + i0], x);
  }
  compat[i] = m32;
  compat[SG + i] = p32;
}

// one box of the tensor map -> shared memory; coordinates (column, site, frame)
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(smem_addr(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_addr(bar))
               : "memory");
}

// Which 32 x 32 piece (row strip, column quarter) of its unit a warp computes; false: none.  Off the diagonal the four warps
// of sub-partition k (warps k, k+4, k+8, k+12) are the four column quarters of row strip k.  On a diagonal tile the pieces
// under the diagonal 64-block do not exist — strips 2 and 3 have no quarters 0 and 1 — so the twelve that do are dealt out
// three to a sub-partition instead of 4/4/2/2.
__host__ __device__ inline bool syrk_piece(int halves, bool diag, int warp, int& strip, int& quarter) {
  if (!diag) {
    strip = warp & 3;
    quarter = warp >> 2;
    return ((halves >> (quarter >> 1)) & 1) != 0;
  }
  int k = (warp >> 2) * 4 + (warp & 3);
  if (halves & 1) {
    if (k < 4) {   // strips 0, 1 x quarters 0, 1
      strip = k >> 1;
      quarter = k & 1;
      return true;
    }
    k -= 4;
  }
  if ((halves & 2) && k < 8) {   // strips 0..3 x quarters 2, 3
    strip = k & 3;
    quarter = 2 + (k >> 2);
    return true;
  }
  strip = quarter = 0;
  return false;
}

// one k-step (4 frames) of a warp: 4 row tiles x kNJ column tiles.  a_addr / b_addr: this thread's element (frame tig of the
// k-step, column gid) of the warp's first row / column tile
template <int kNJ>
__device__ __forceinline__ void syrk_kstep(double (&acc)[4][4][2], uint32_t a_addr, uint32_t b_addr) {
  double a[4], b[kNJ];
#pragma unroll
  for (int i = 0; i < 4; i++) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a[i]) : "r"(a_addr + i * 64));
#pragma unroll
  for (int j = 0; j < kNJ; j++) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(b[j]) : "r"(b_addr + j * 64));
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < kNJ; j++) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
}

// unit = halves << 28 | s << 16 | bi << 8 | bj  (128-blocks, bi <= bj; halves: bit 0 = columns 0..63, bit 1 = columns 64..127)
__global__ void __launch_bounds__(kSyrkThreads, 1)
cov_syrk_kernel(const __grid_constant__ CUtensorMap dmap, int n_frames, double* __restrict__ buf, int S, int G,
                const int* __restrict__ sched_off, const int* __restrict__ sched_unit) {
  extern __shared__ unsigned char syrk_smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kSyrkStages], empty_bar[kSyrkStages];
  unsigned char* ring = syrk_smem_raw + ((128u - (smem_addr(syrk_smem_raw) & 127u)) & 127u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gid = lane >> 2, tig = lane & 3;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < kSyrkStages; i++) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], kSyrkThreads / 32);
    }
    mbar_fence_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&dmap) : "memory");
  }
  __syncthreads();

  const int u_begin = sched_off[blockIdx.x], my_units = sched_off[blockIdx.x + 1] - u_begin;
  const int k_stages = (n_frames + kSyrkK - 1) / kSyrkK;
  const long long total = (long long)my_units * k_stages;   // stages this CTA consumes
  if (total == 0) return;
  // producer (thread 0): stage n of this CTA's sequence
  auto issue = [&](long long n) {
    const int local = (int)(n / k_stages);
    const int ks = (int)(n - (long long)local * k_stages);
    const int unit = sched_unit[u_begin + local];
    const int halves = (unit >> 28) & 3, s = (unit >> 16) & 0xfff, bi = (unit >> 8) & 255, bj = unit & 255;
    const int st = (int)(n % kSyrkStages);
    unsigned char* dst = ring + (size_t)st * kSyrkStageBytes;
    const bool diag = bi == bj;   // B's columns are inside A's
    mbar_expect_tx(&full_bar[st], diag ? kSyrkOperandBytes : kSyrkStageBytes);
    tma_load_3d(dst, &dmap, bi * kSyrkTile, s, ks * kSyrkK, &full_bar[st]);
    if (!diag) tma_load_3d(dst + kSyrkOperandBytes, &dmap, bj * kSyrkTile + (halves == 2 ? 64 : 0), s, ks * kSyrkK, &full_bar[st]);
  };
  if (threadIdx.x == 0) {
    for (long long n = 0; n < kSyrkStages - 1 && n < total; n++) issue(n);
  }

  const uint32_t ring_addr = smem_addr(ring);
  const uint32_t t_off = (uint32_t)(tig * kSyrkPitch + gid * 8);   // this thread's element of a fragment: frame tig, column gid

  double acc[4][4][2];
  long long n = 0;
  for (int lu = 0; lu < my_units; lu++) {
    const int unit = sched_unit[u_begin + lu];
    const int halves = (unit >> 28) & 3, s = (unit >> 16) & 0xfff, bi = (unit >> 8) & 255, bj = unit & 255;
    int strip, quarter;
    const bool has_piece = syrk_piece(halves, bi == bj, warp, strip, quarter);
    const int wr0 = bi * kSyrkTile + 32 * strip, wc0 = bj * kSyrkTile + 32 * quarter;
    // this warp's 32 x 32 piece: in a half the unit covers, inside G, and not below the diagonal 64-block
    const bool active = has_piece && wr0 < G && wc0 < G && (wr0 / kCovBlk) <= (wc0 / kCovBlk);
    const int nj = active ? min(4, (G - wc0 + 7) / 8) : 0;   // column tiles of 8 that hold real columns
    // where this warp's B columns sit in the stage: inside the A operand on a diagonal tile, else in the B operand, whose
    // first column is the tile's column 64 when the unit covers only the second half
    const uint32_t b_off = bi == bj ? (uint32_t)(32 * quarter * 8)
                                    : (uint32_t)(kSyrkOperandBytes + (32 * quarter - (halves == 2 ? 64 : 0)) * 8);
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int ks = 0; ks < k_stages; ks++, n++) {
      if (threadIdx.x == 0) {
        const long long nxt = n + kSyrkStages - 1;
        if (nxt < total) {
          if (nxt >= kSyrkStages) mbar_wait(&empty_bar[nxt % kSyrkStages], (uint32_t)((nxt / kSyrkStages - 1) & 1));
          issue(nxt);
        }
      }
      __syncwarp();
      const int st = (int)(n % kSyrkStages);
      mbar_wait(&full_bar[st], (uint32_t)((n / kSyrkStages) & 1));
      const uint32_t a_addr = ring_addr + (uint32_t)st * kSyrkStageBytes + (uint32_t)(32 * strip * 8) + t_off;
      const uint32_t b_addr = ring_addr + (uint32_t)st * kSyrkStageBytes + b_off + t_off;
      if (nj > 2) {
#pragma unroll
        for (int q = 0; q < kSyrkK / 4; q++) syrk_kstep<4>(acc, a_addr + q * 4 * kSyrkPitch, b_addr + q * 4 * kSyrkPitch);
      } else if (nj > 1) {
#pragma unroll
        for (int q = 0; q < kSyrkK / 4; q++) syrk_kstep<2>(acc, a_addr + q * 4 * kSyrkPitch, b_addr + q * 4 * kSyrkPitch);
      } else if (nj > 0) {
#pragma unroll
        for (int q = 0; q < kSyrkK / 4; q++) syrk_kstep<1>(acc, a_addr + q * 4 * kSyrkPitch, b_addr + q * 4 * kSyrkPitch);
      }
      // every read of the stage is done (the values are in registers): hand it back to the producer
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[st]);
    }
    if (active) {
      double* sumsq = buf + 1 + 2 * (size_t)S * G + (size_t)s * G * G;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int r = wr0 + i * 8 + gid;
        if (r >= G) continue;
        double* row = sumsq + (size_t)r * G;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int c = wc0 + j * 8 + tig * 2;
          if (c < G) atomicAdd(row + c, acc[i][j][0]);
          if (c + 1 < G) atomicAdd(row + c + 1, acc[i][j][1]);
        }
      }
    }
  }
}

// cuTensorMapEncodeTiled, bound through the runtime (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess) {
      p = nullptr;
    }
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// The units of all sites dealt out to `ctas` lists, longest first (LPT): cost = the busiest sub-partition's DMMA count per
// k-step (in column tiles) + a fixed cost for walking the stages.  The cheapest tiles — about two rounds' worth — are cut
// into their two column halves, so the lists can be evened out to within half a tile.
// One schedule per (device, S, G), built on first use and kept on the device.
struct SyrkSchedule {
  int S = 0, G = 0, ctas = 0;
  int* d_off = nullptr;    // [ctas + 1]
  int* d_unit = nullptr;   // [n_units]
};
static int syrk_unit_cost(int G, int bi, int bj, int halves) {
  int per_sp[4] = {0, 0, 0, 0};   // column tiles per k-step each sub-partition's warps compute
  for (int warp = 0; warp < kSyrkThreads / 32; warp++) {
    int strip, quarter;
    if (!syrk_piece(halves, bi == bj, warp, strip, quarter)) continue;
    const int wr0 = bi * kSyrkTile + 32 * strip, wc0 = bj * kSyrkTile + 32 * quarter;
    if (wr0 < G && wc0 < G && wr0 / kCovBlk <= wc0 / kCovBlk) {
      const int nj = std::min(4, (G - wc0 + 7) / 8);
      per_sp[warp & 3] += nj > 2 ? 4 : nj > 1 ? 2 : 1;
    }
  }
  return std::max(std::max(per_sp[0], per_sp[1]), std::max(per_sp[2], per_sp[3]));
}
static const SyrkSchedule* syrk_schedule(int S, int G, int ctas) {
  static std::mutex mu;
  static std::vector<SyrkSchedule> cache[16];
  std::lock_guard<std::mutex> lock(mu);
  auto& mine = cache[device_index()];
  for (auto& e : mine)
    if (e.S == S && e.G == G && e.ctas == ctas) return &e;
  const int nb = (G + kSyrkTile - 1) / kSyrkTile;
  struct U { int id; int cost; };
  std::vector<U> units;
  for (int s = 0; s < S; s++)
    for (int bi = 0; bi < nb; bi++)
      for (int bj = bi; bj < nb; bj++) {
        const int c = syrk_unit_cost(G, bi, bj, 3);
        if (c > 0) units.push_back({(3 << 28) | (s << 16) | (bi << 8) | bj, c + 1});
      }
  std::stable_sort(units.begin(), units.end(), [](const U& a, const U& b) { return a.cost > b.cost; });
  // cut the cheapest tiles in halves
  // half a round's worth, measured (C5, TFLOP/s): none 27.4, 0.5 rounds 27.8, 1 round 27.6, 2 rounds 26.0, 3 rounds 25.5 — a half
  // unit leaves two of a sub-partition's four warps without work, so every split costs pipe utilisation while it runs
  size_t n_split = std::min(units.size(), (size_t)ctas / 2);
  if (const char* e = getenv("CGC_SYRK_SPLIT_PCT")) n_split = std::min(units.size(), (size_t)(atoll(e) * ctas / 100));   // tuning hook
  std::vector<U> tail(units.end() - (long)n_split, units.end());
  units.resize(units.size() - n_split);
  for (const U& t : tail) {
    const int bi = (t.id >> 8) & 255, bj = t.id & 255;
    for (int h = 1; h <= 2; h++) {
      const int c = syrk_unit_cost(G, bi, bj, h);
      if (c > 0) units.push_back({(t.id & 0x0fffffff) | (h << 28), c + 1});
    }
  }
  std::stable_sort(units.begin(), units.end(), [](const U& a, const U& b) { return a.cost > b.cost; });
  std::vector<std::vector<int>> lists((size_t)ctas);
  std::vector<long long> load((size_t)ctas, 0);
  for (const U& u : units) {
    int best = 0;
    for (int c = 1; c < ctas; c++)
      if (load[(size_t)c] < load[(size_t)best]) best = c;
    lists[(size_t)best].push_back(u.id);
    load[(size_t)best] += u.cost;
  }
  std::vector<int> off((size_t)ctas + 1, 0), flat;
  for (int c = 0; c < ctas; c++) {
    off[(size_t)c + 1] = off[(size_t)c] + (int)lists[(size_t)c].size();
    flat.insert(flat.end(), lists[(size_t)c].begin(), lists[(size_t)c].end());
  }
  SyrkSchedule e;
  e.S = S; e.G = G; e.ctas = ctas;
  if (cudaMalloc(&e.d_off, sizeof(int) * off.size()) != cudaSuccess || cudaMalloc(&e.d_unit, sizeof(int) * std::max<size_t>(1, flat.size())) != cudaSuccess ||
      cudaMemcpy(e.d_off, off.data(), sizeof(int) * off.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(e.d_unit, flat.data(), sizeof(int) * flat.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
    return nullptr;
  }
  mine.push_back(e);
  return &mine.back();
}

cudaError_t launch_cov_accumulate(const float* dE, int n_frames, double* buf, float* compat, double* centred_scratch, int S,
                                  int G, cudaStream_t st) {
  if (n_frames <= 0) return cudaSuccess;
  if (S >= 4096 || (G + kSyrkTile - 1) / kSyrkTile > 255) return cudaErrorInvalidValue;
  const int Gp = (int)cov_padded_cols(G);
  const int SG = S * G;
  const int sm_count = cached_sm_count();
  dim3 cgrid((S * Gp + 255) / 256, std::max(1, std::min(std::min(n_frames, kCentreChunksMax), 8 * sm_count / std::max(1, (S * Gp + 255) / 256))));
  double* part = centred_scratch + (size_t)n_frames * S * Gp;
  cov_centre_kernel<<<cgrid, 256, 0, st>>>(dE, n_frames, buf, S, G, Gp, centred_scratch, part);
  cov_sum_finish_kernel<<<(SG + 255) / 256, 256, 0, st>>>(part, (int)cgrid.y, n_frames, buf, SG);
  if (compat) cov_compat_kernel<<<(SG + 127) / 128, 128, 0, st>>>(dE, n_frames, SG, G, compat);
  // tensor map over D [frame][site][Gp] (innermost first), box = 132 columns x 1 site x 16 frames, no swizzle,
  // out-of-range elements read as zero
  EncodeTiledFn encode = encode_tiled_fn();
  if (!encode) return cudaErrorNotSupported;
  CUtensorMap map;
  const cuuint64_t dims[3] = {(cuuint64_t)Gp, (cuuint64_t)S, (cuuint64_t)n_frames};
  const cuuint64_t strides[2] = {(cuuint64_t)Gp * 8, (cuuint64_t)S * Gp * 8};
  const cuuint32_t box[3] = {(cuuint32_t)kSyrkBoxCols, 1, (cuuint32_t)kSyrkK};
  const cuuint32_t estr[3] = {1, 1, 1};
  if (encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, centred_scratch, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    return cudaErrorInvalidValue;
  }
  const int nb = (G + kSyrkTile - 1) / kSyrkTile;
  const int n_tiles = S * nb * (nb + 1) / 2;
  const int grid = std::min(n_tiles, sm_count);
  const SyrkSchedule* sched = syrk_schedule(S, G, grid);
  if (!sched) return cudaErrorMemoryAllocation;
  static size_t granted[16] = {};
  ensure_dynamic_smem(cov_syrk_kernel, kSyrkSmem, granted);
  cov_syrk_kernel<<<grid, kSyrkThreads, kSyrkSmem, st>>>(map, n_frames, buf, S, G, sched->d_off, sched->d_unit);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------------
// K4  mode projection (depca's ApplyPCA, reference scripts/depca/depca/convert.py:84-125) on FP64 tensor cores
// ---------------------------------------------------------------------------------------------------------------------
// out[t][s][m] = sum_g (x[t][s][g] - mean[s][g]) * modes[s][m][g] : per site a (frames x G) by (G x M) GEMM.
// centre_rows_kernel writes D = x - mean in FP64; project_kernel computes 64 x 64 tiles of D_s W_s^T with
// mma.sync m8n8k4 f64 (4 warps, each a 32 x 32 quarter), both operands K-contiguous in global memory and staged by
// cp.async into double-buffered tiles whose 20-double row stride makes the fragment loads conflict-free.
__global__ void centre_rows_kernel(const float* __restrict__ x, const double* __restrict__ mean, double* __restrict__ d,
                                   int64_t n, int SG) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) d[i] = (double)x[i] - mean[i % SG];
}

__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gsrc, bool valid) {
  const int n = valid ? 8 : 0;   // src-size 0 -> the 8 bytes are zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_addr(smem_dst)), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kPrjBlk = 64;    // tile edge (frames and modes)
constexpr int kPrjK = 16;      // columns of G per stage
constexpr int kPrjLd = 20;     // row stride in doubles: (gid * 20 + tig) mod 16 is a permutation of 0..15 for gid, tig < 4

__global__ void __launch_bounds__(128)
project_kernel(const double* __restrict__ D, const double* __restrict__ W, double* __restrict__ out, int T, int S, int G,
               int M) {
  __shared__ __align__(16) double As[2][kPrjBlk][kPrjLd];   // frames x k
  __shared__ __align__(16) double Bs[2][kPrjBlk][kPrjLd];   // modes x k
  const int t0 = blockIdx.x * kPrjBlk, m0 = blockIdx.y * kPrjBlk, s = blockIdx.z;
  const int SG = S * G;
  const double* Ds = D + (size_t)s * G;               // row t at Ds + t * SG
  const double* Ws = W + (size_t)s * M * G;           // row m at Ws + m * G
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wi = (warp >> 1) * 32, wj = (warp & 1) * 32;
  const int gid = lane >> 2, tig = lane & 3;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  auto stage = [&](int b, int k0) {
    for (int idx = threadIdx.x; idx < kPrjBlk * kPrjK; idx += 128) {
      const int r = idx / kPrjK, k = idx - r * kPrjK;
      const int g = k0 + k, t = t0 + r, m = m0 + r;
      const bool gv = g < G;
      cp_async_8(&As[b][r][k], Ds + (size_t)(t < T ? t : 0) * SG + (gv ? g : 0), gv && t < T);
      cp_async_8(&Bs[b][r][k], Ws + (size_t)(m < M ? m : 0) * G + (gv ? g : 0), gv && m < M);
    }
    cp_async_commit();
  };

  const int n_chunks = (G + kPrjK - 1) / kPrjK;
  stage(0, 0);
  for (int ch = 0; ch < n_chunks; ch++) {
    const int cb = ch & 1;
    if (ch + 1 < n_chunks) {
      stage(cb ^ 1, (ch + 1) * kPrjK);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
#pragma unroll
    for (int k0 = 0; k0 < kPrjK; k0 += 4) {
      double a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = As[cb][wi + i * 8 + gid][k0 + tig];
#pragma unroll
      for (int j = 0; j < 4; j++) bb[j] = Bs[cb][wj + j * 8 + gid][k0 + tig];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], bb[j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int t = t0 + wi + i * 8 + gid;
    if (t >= T) continue;
    double* row = out + ((size_t)t * S + s) * M;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int m = m0 + wj + j * 8 + tig * 2;
      if (m < M) row[m] = acc[i][j][0];
      if (m + 1 < M) row[m + 1] = acc[i][j][1];
    }
  }
}

void launch_project(const float* x, int n_frames, const double* mean, const double* modes, int n_modes, double* centred,
                    double* out, int S, int G, cudaStream_t st) {
  if (n_frames <= 0 || n_modes <= 0) return;
  const int64_t n = (int64_t)n_frames * S * G;
  int64_t blocks = std::min<int64_t>((n + 255) / 256, 148 * 16);
  centre_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, mean, centred, n, S * G);
  dim3 grid((n_frames + kPrjBlk - 1) / kPrjBlk, (n_modes + kPrjBlk - 1) / kPrjBlk, S);
  project_kernel<<<grid, 128, 0, st>>>(centred, modes, out, n_frames, S, G, n_modes);
}

// mean = c + sum/n ; cov_ij = (sumsq_ij - sum_i sum_j / n) / (n - ddof), upper blocks mirrored
__global__ void cov_finalize_kernel(const double* __restrict__ buf, int S, int G, int ddof, double* __restrict__ mean,
                                    double* __restrict__ cov) {
  const int s = blockIdx.z;
  const int i = blockIdx.y * blockDim.y + threadIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G || j >= G) return;
  const int SG = S * G;
  const double n = buf[0];
  const double* sum = buf + 1 + SG + (size_t)s * G;
  const double* sumsq = buf + 1 + 2 * (size_t)SG + (size_t)s * G * G;
  // the stored half: 64x64 blocks with block-row <= block-column
  const bool stored = (i / kCovBlk) <= (j / kCovBlk);
  const double ss = stored ? sumsq[(size_t)i * G + j] : sumsq[(size_t)j * G + i];
  cov[(size_t)s * G * G + (size_t)i * G + j] = (ss - sum[i] * sum[j] / n) / (n - (double)ddof);
  if (i == 0) mean[(size_t)s * G + j] = buf[1 + (size_t)s * G + j] + sum[j] / n;
}

__global__ void narrow_kernel(const double* __restrict__ in, float* __restrict__ out, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = (float)in[i];
}

void launch_narrow(const double* in, float* out, int64_t n, cudaStream_t st) {
  if (n <= 0) return;
  narrow_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16), 256, 0, st>>>(in, out, n);
}

void launch_cov_finalize(const double* buf, int S, int G, int ddof, double* mean, double* cov, cudaStream_t st) {
  dim3 block(32, 8);
  dim3 grid((G + 31) / 32, (G + 7) / 8, S);
  cov_finalize_kernel<<<grid, block, 0, st>>>(buf, S, G, ddof, mean, cov);
}

}  // namespace cgc

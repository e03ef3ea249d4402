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

#include <cstdio>

namespace cgc {

// ---------------------------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk TMA (cp.async.bulk -> SASS UBLKCP), packed FP32 (FFMA2/FADD2/FMUL2), MUFU.RSQ
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// global -> shared bulk copy through the TMA unit; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}

// Packed FP32 pairs live in 64-bit registers for their whole life (so ptxas never re-packs them with MOVs).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 fadd2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 fmul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
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

__global__ void __launch_bounds__(kDecodeThreads)
decode_kernel(const char* __restrict__ text, int64_t text_bytes, const int64_t* __restrict__ atoms_off, int n_atoms,
              int n_pad, int line_bytes, int x_col, int field_w, int ndec,
              const int32_t* __restrict__ slot, const float* __restrict__ slot_dq, int sites_stride /* float4 per frame */,
              float* __restrict__ xyz8, float4* __restrict__ sites, int* __restrict__ err_flags) {
  extern __shared__ __align__(128) unsigned char sm_raw[];
  __shared__ __align__(8) uint64_t bar;
  const int f = blockIdx.y;
  const int a0 = blockIdx.x * kDecodeThreads;
  const int a = a0 + threadIdx.x;
  int n_here = n_atoms - a0;
  n_here = n_here < 0 ? 0 : (n_here > kDecodeThreads ? kDecodeThreads : n_here);
  uint32_t head = 0;
  if (n_here > 0) {
    const int64_t src = atoms_off[f] + (int64_t)a0 * line_bytes;
    const int64_t src_al = src & ~(int64_t)15;
    head = (uint32_t)(src - src_al);
    int64_t bytes = (int64_t)head + (int64_t)n_here * line_bytes;
    bytes = (bytes + 15) & ~(int64_t)15;
    const int64_t lim = ((text_bytes + 15) & ~(int64_t)15) - src_al;  // never read past the (16-byte padded) text
    if (bytes > lim) bytes = lim;
    if (threadIdx.x == 0) {
      mbar_init(&bar, 1);
      mbar_fence_init();
      mbar_expect_tx(&bar, (uint32_t)bytes);
      tma_bulk_g2s(sm_raw, text + src_al, (uint32_t)bytes, &bar);
    }
    __syncthreads();  // barrier initialised before anyone polls it
    mbar_wait(&bar, 0);
  }
  if (a >= n_pad) return;
  float3 out;
  if (a < n_atoms) {
    const unsigned char* line = sm_raw + head + threadIdx.x * line_bytes;
    int err = 0;
    float p10f = 1.f;
    double p10d = 1.0;
    for (int k = 0; k < ndec; k++) { p10f *= 10.f; p10d *= 10.0; }
    const bool wide = field_w > 8;  // more than 7 digits could exceed 2^24
    out.x = parse_field(line + x_col, field_w, ndec, p10f, p10d, wide, err);
    out.y = parse_field(line + x_col + field_w, field_w, ndec, p10f, p10d, wide, err);
    out.z = parse_field(line + x_col + 2 * field_w, field_w, ndec, p10f, p10d, wide, err);
    const bool last_unterminated = (atoms_off[f] + (int64_t)(a + 1) * line_bytes - 1) >= text_bytes;
    if (!last_unterminated && line[line_bytes - 1] != '\n') err |= kDecBadNewline;
    if (err) atomicOr(err_flags, err);
    const int sl = slot[a];
    if (sl >= 0) {
      // site block entry, laid out for the packed pair loop: (-x,-x,-y,-y) (-z,-z,dq,dq)
      const float dq = slot_dq[sl];
      float4* dst = sites + (size_t)f * sites_stride + (size_t)sl * 2;
      dst[0] = make_float4(-out.x, -out.x, -out.y, -out.y);
      dst[1] = make_float4(-out.z, -out.z, dq, dq);
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
  sites[2 * i + 0] = make_float4(CGC_PAD_NEG_COORD, CGC_PAD_NEG_COORD, CGC_PAD_NEG_COORD, CGC_PAD_NEG_COORD);
  sites[2 * i + 1] = make_float4(CGC_PAD_NEG_COORD, CGC_PAD_NEG_COORD, 0.f, 0.f);
}

void launch_init_sites(const DeviceSystem& sys, float4* sites, int n_frames, cudaStream_t st) {
  int64_t n = (int64_t)n_frames * sys.n_sites * sys.site_cap;
  if (n == 0) return;
  init_sites_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sites, n);
}

void launch_decode(const DeviceSystem& sys, const char* text, int64_t text_bytes, const int64_t* atoms_off,
                   int n_frames, float* xyz8, float4* sites, int* err_flags, cudaStream_t st) {
  dim3 grid((sys.n_pad + kDecodeThreads - 1) / kDecodeThreads, n_frames);
  size_t smem = (size_t)kDecodeThreads * sys.line_bytes + 32;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaFuncSetAttribute(decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  decode_kernel<<<grid, kDecodeThreads, smem, st>>>(text, text_bytes, atoms_off, sys.n_atoms, sys.n_pad, sys.line_bytes,
                                                    sys.x_col, sys.field_w, sys.ndec, sys.slot, sys.slot_dq,
                                                    sys.n_sites * sys.site_cap * 2, xyz8, sites, err_flags);
}

// ---------------------------------------------------------------------------------------------------------------------
// K2  CDC pair kernel
// ---------------------------------------------------------------------------------------------------------------------
// Work item = (frame, tile of kTileAtoms consecutive atoms).  Persistent CTAs stride over the items.  Each thread keeps
// E = 8 consecutive atoms in registers as 4 packed pairs; the frame's site block (S x P entries of 32 B) arrives in
// shared memory by one bulk TMA copy, double-buffered so the next item's block lands while this one is computed.
// For every site s and site atom c, all threads read the same 32 B (broadcast LDS.128 x2) and evaluate 8 pairs:
//   dx = x - xc (FADD2 x3), r2 = dx*dx + dy*dy + dz*dz (FMUL2 + 2 FFMA2), rinv = MUFU.RSQ (x2), V += dQ*rinv (FFMA2)
// i.e. 7 packed FP32 instructions + 2 MUFU per 2 pairs.  V_s(a) = sum_c dQ_c / r_ac is FP32 over <= P terms; everything
// after that is FP64: w = V * (33.2f * q_a), summed per residue group (atoms of a group are consecutive in the file,
// so a thread's 8 atoms form at most a few runs), warp-shuffle reduced when the whole warp sits in one group, and
// added to the FP64 bin acc[frame][site][column].
__global__ void __launch_bounds__(kTileThreads, 2)
cdc_kernel(const float* __restrict__ xyz8, const float4* __restrict__ sites, const double* __restrict__ kq_all,
           const int32_t* __restrict__ col_all, double* __restrict__ acc, int n_pad, int tiles_per_frame, int n_items,
           int S, int P, int G) {
  extern __shared__ __align__(128) unsigned char sm_raw[];
  __shared__ __align__(8) uint64_t bars[2];
  const uint32_t block_bytes = (uint32_t)S * P * 32u;

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncthreads();

  int item = blockIdx.x;
  if (item >= n_items) return;
  if (threadIdx.x == 0) {
    const int f0 = item / tiles_per_frame;
    mbar_expect_tx(&bars[0], block_bytes);
    tma_bulk_g2s(sm_raw, sites + (size_t)f0 * S * P * 2, block_bytes, &bars[0]);
  }
  uint32_t phase_bits = 0;  // bit b = parity to wait for on bars[b]
  int cur = 0;
  const int lane = threadIdx.x & 31;

  for (; item < n_items; item += gridDim.x) {
    const int f = item / tiles_per_frame;
    const int tile = item - f * tiles_per_frame;
    const int next = item + gridDim.x;
    if (threadIdx.x == 0 && next < n_items) {
      // buf[cur^1] was last read two iterations ago; the __syncthreads at the end of the previous one orders that
      const int fn = next / tiles_per_frame;
      mbar_expect_tx(&bars[cur ^ 1], block_bytes);
      tma_bulk_g2s(sm_raw + (cur ^ 1) * block_bytes, sites + (size_t)fn * S * P * 2, block_bytes, &bars[cur ^ 1]);
    }

    // ---- this thread's 8 atoms
    const int a0 = tile * kTileAtoms + threadIdx.x * kAtomsPerThread;
    const ulonglong2* ap = reinterpret_cast<const ulonglong2*>(xyz8 + ((size_t)f * n_pad + a0) * 3);
    f32x2 X[4], Y[4], Z[4];
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const ulonglong2 vx = __ldg(ap + j), vy = __ldg(ap + 2 + j), vz = __ldg(ap + 4 + j);
      X[2 * j] = vx.x; X[2 * j + 1] = vx.y;
      Y[2 * j] = vy.x; Y[2 * j + 1] = vy.y;
      Z[2 * j] = vz.x; Z[2 * j + 1] = vz.y;
    }
    int cols[8];
    {
      const int4 c0 = __ldg(reinterpret_cast<const int4*>(col_all + a0));
      const int4 c1 = __ldg(reinterpret_cast<const int4*>(col_all + a0 + 4));
      cols[0] = c0.x; cols[1] = c0.y; cols[2] = c0.z; cols[3] = c0.w;
      cols[4] = c1.x; cols[5] = c1.y; cols[6] = c1.z; cols[7] = c1.w;
    }
    bool same = true;
#pragma unroll
    for (int e = 1; e < 8; e++) same = same && (cols[e] == cols[0]);
    const int c_lane0 = __shfl_sync(0xffffffffu, cols[0], 0);
    // whole warp inside one environment group (columns < 0 are padding; chromophore columns never fill a warp)
    const bool warp_uniform = __all_sync(0xffffffffu, same && cols[0] == c_lane0) && c_lane0 >= 0;

    mbar_wait(&bars[cur], (phase_bits >> cur) & 1u);
    phase_bits ^= 1u << cur;
    const ulonglong2* sblk = reinterpret_cast<const ulonglong2*>(sm_raw + cur * block_bytes);
    double* acc_frame = acc + (size_t)f * S * G;

    for (int s = 0; s < S; s++) {
      f32x2 V[4];
#pragma unroll
      for (int j = 0; j < 4; j++) V[j] = 0ull;
      const ulonglong2* sp = sblk + (size_t)s * P * 2;
#pragma unroll 2
      for (int c = 0; c < P; c++) {
        const ulonglong2 A = sp[2 * c];      // (-x,-x) (-y,-y)
        const ulonglong2 B = sp[2 * c + 1];  // (-z,-z) (dq,dq)
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const f32x2 dx = fadd2(X[j], A.x), dy = fadd2(Y[j], A.y), dz = fadd2(Z[j], B.x);
          f32x2 r2 = fmul2(dx, dx);
          r2 = ffma2(dy, dy, r2);
          r2 = ffma2(dz, dz, r2);
          float r2a, r2b;
          unpack2(r2, r2a, r2b);
          const f32x2 ri = pack2(rsqrt_approx(r2a), rsqrt_approx(r2b));
          V[j] = ffma2(B.y, ri, V[j]);
        }
      }
      float Vf[8];
#pragma unroll
      for (int j = 0; j < 4; j++) unpack2(V[j], Vf[2 * j], Vf[2 * j + 1]);
      // ---- FP64 from here: w_e = V_e * 33.2f * q_e, binned per column
      const double* kq = kq_all + a0;
      double* acc_row = acc_frame + (size_t)s * G;
      if (warp_uniform) {
        double t = 0.0;
#pragma unroll
        for (int e = 0; e < 8; e++) t += (double)Vf[e] * __ldg(kq + e);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) atomicAdd(acc_row + c_lane0, t);
      } else {
        double run = 0.0;
        int cc = cols[0];
#pragma unroll
        for (int e = 0; e < 8; e++) {
          if (cols[e] != cc) {
            if (cc >= 0 && cc != s) atomicAdd(acc_row + cc, run);
            cc = cols[e];
            run = 0.0;
          }
          run += (double)Vf[e] * __ldg(kq + e);
        }
        if (cc >= 0 && cc != s) atomicAdd(acc_row + cc, run);
      }
    }
    __syncthreads();  // everyone is done with buf[cur] before it is refilled
    cur ^= 1;
  }
}

size_t cdc_smem_bytes(const DeviceSystem& sys) { return (size_t)2 * sys.n_sites * sys.site_cap * 32; }

void launch_cdc(const DeviceSystem& sys, const float* xyz8, const float4* sites, double* acc, int n_frames,
                int sm_count, cudaStream_t st) {
  const int tiles = sys.n_pad / kTileAtoms;
  const int n_items = tiles * n_frames;
  if (n_items == 0) return;
  size_t smem = cdc_smem_bytes(sys);
  static size_t configured = 0;
  if (smem > configured) {
    cudaFuncSetAttribute(cdc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  int per_sm = 2;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cdc_kernel, kTileThreads, smem);
  if (per_sm < 1) per_sm = 1;
  int grid = sm_count * per_sm;
  if (grid > n_items) grid = n_items;
  cdc_kernel<<<grid, kTileThreads, smem, st>>>(xyz8, sites, sys.kq, sys.col, acc, sys.n_pad, tiles, n_items, sys.n_sites,
                                               sys.site_cap, sys.num_tot);
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
// sumsq_s += D_s^T D_s with D_s the (frames x G) matrix of one site: a SYRK.  The G x G result is cut into 64 x 64
// blocks; only blocks with bi <= bj are computed (finalize mirrors them).  One CTA of 4 warps owns one block of one
// site for the whole batch; each warp owns a 32 x 32 quarter = 4 x 4 DMMA m8n8k4 tiles (32 FP64 accumulators/thread).
constexpr int kCovBlk = 64;
constexpr int kCovChunk = 32;   // frames staged in shared memory per step
constexpr int kCovLd = 68;      // row stride in doubles: 136 words == 8 (mod 32) -> conflict-free DMMA fragment loads

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

// Also keeps the two float32 running sums the reference's literal .mtx arithmetic needs (src/FMOparse.cpp:173-175 with
// the zero stride of :193,202): sum_t x_i and sum_t x_0 x_j, added frame by frame in order like the reference does.
__global__ void cov_sum_kernel(const float* __restrict__ dE, int n_frames, double* __restrict__ buf, int SG, int G,
                               float* __restrict__ compat /* [2][SG] or null */) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) buf[0] += (double)n_frames;
  if (i >= SG) return;
  const double c = buf[1 + i];
  double s = 0.0;
  const int i0 = (i / G) * G;  // column 0 of this site
  float m32 = compat ? compat[i] : 0.f, p32 = compat ? compat[SG + i] : 0.f;
  for (int t = 0; t < n_frames; t++) {
    const float x = dE[(size_t)t * SG + i];
    s += (double)x - c;
    m32 += x;
    p32 += __fmul_rn(dE[(size_t)t * SG + i0], x);
  }
  buf[1 + SG + i] += s;
  if (compat) { compat[i] = m32; compat[SG + i] = p32; }
}

__global__ void __launch_bounds__(128)
cov_syrk_kernel(const float* __restrict__ dE, int n_frames, double* __restrict__ buf, int S, int G, int nblk) {
  __shared__ double Ai[kCovChunk][kCovLd];
  __shared__ double Aj[kCovChunk][kCovLd];
  // upper-triangular block index -> (bi, bj)
  int b = blockIdx.x, bi = 0;
  while (b >= nblk - bi) { b -= nblk - bi; bi++; }
  const int bj = bi + b;
  const int s = blockIdx.y;
  const int SG = S * G;
  const double* shift = buf + 1 + (size_t)s * G;
  double* sumsq = buf + 1 + 2 * (size_t)SG + (size_t)s * G * G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wi = (warp >> 1) * 32, wj = (warp & 1) * 32;  // this warp's quarter of the block
  const int gid = lane >> 2, tig = lane & 3;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int t0 = 0; t0 < n_frames; t0 += kCovChunk) {
    // stage (x - c) for both column blocks; zero beyond G or beyond the batch
    for (int idx = threadIdx.x; idx < kCovChunk * kCovBlk; idx += 128) {
      const int k = idx / kCovBlk, cidx = idx - k * kCovBlk;
      const int t = t0 + k;
      const int gi = bi * kCovBlk + cidx, gj = bj * kCovBlk + cidx;
      double vi = 0.0, vj = 0.0;
      if (t < n_frames) {
        const float* row = dE + (size_t)t * SG + (size_t)s * G;
        if (gi < G) vi = (double)row[gi] - shift[gi];
        if (gj < G) vj = (double)row[gj] - shift[gj];
      }
      Ai[k][cidx] = vi;
      Aj[k][cidx] = vj;
    }
    __syncthreads();
#pragma unroll
    for (int k0 = 0; k0 < kCovChunk; k0 += 4) {
      double a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = Ai[k0 + tig][wi + i * 8 + gid];
#pragma unroll
      for (int j = 0; j < 4; j++) bb[j] = Aj[k0 + tig][wj + j * 8 + gid];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], bb[j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int r = bi * kCovBlk + wi + i * 8 + gid;
      const int c = bj * kCovBlk + wj + j * 8 + tig * 2;
      if (r < G) {
        if (c < G) sumsq[(size_t)r * G + c] += acc[i][j][0];
        if (c + 1 < G) sumsq[(size_t)r * G + c + 1] += acc[i][j][1];
      }
    }
}

void launch_cov_accumulate(const float* dE, int n_frames, double* buf, float* compat, int S, int G, cudaStream_t st) {
  if (n_frames <= 0) return;
  const int SG = S * G;
  const int nblk = (G + kCovBlk - 1) / kCovBlk;
  dim3 grid(nblk * (nblk + 1) / 2, S);
  cov_syrk_kernel<<<grid, 128, 0, st>>>(dE, n_frames, buf, S, G, nblk);
  cov_sum_kernel<<<(SG + 255) / 256, 256, 0, st>>>(dE, n_frames, buf, SG, G, compat);
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

void launch_cov_finalize(const double* buf, int S, int G, int ddof, double* mean, double* cov, cudaStream_t st) {
  dim3 block(32, 8);
  dim3 grid((G + 31) / 32, (G + 7) / 8, S);
  cov_finalize_kernel<<<grid, block, 0, st>>>(buf, S, G, ddof, mean, cov);
}

}  // namespace cgc

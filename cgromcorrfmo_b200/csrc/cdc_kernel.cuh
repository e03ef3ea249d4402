// K2 — the CDC pair kernel (replaces ComputeCDC_v1, reference src/grom2atomFMO.cpp:625-739), sm_100a.
//
// Work unit = (frame, tile of kTileAtoms = 2048 consecutive atoms, chunk of <= 8 sites).  The units are numbered with
// the chunk index fastest, and every persistent CTA (256 threads, 2 per SM) takes one CONTIGUOUS range of
// n_units / gridDim units (+-1): the load is balanced to within one chunk (C2, 128 frames: 18816 units over 296 CTAs =
// 63.6 each, 0.7 % tail; whole (frame, tile) items strided over the CTAs left a 3.7 % tail), and a CTA reloads its
// atoms only when its range crosses into the next (frame, tile).
//
// Each thread keeps 8 consecutive atoms in registers as 4 packed pairs (one block of the xyz8 layout K1 writes, loaded
// as 64-bit values).  For every site s and site atom c, all threads read the same 16 B from shared memory (one
// broadcast LDS.128: -xc, -yc, -zc, dQ) and evaluate 8 pairs:
//   dx = x - xc (FADD2 x3), r2 = dx*dx + dy*dy + dz*dz (FMUL2 + 2 FFMA2), rinv = MUFU.RSQ (x2), V += dQ*rinv (FFMA2)
// i.e. 7 packed FP32 instructions + 2 MUFU per 2 pairs: the XU pipe (16 rsqrt/clk/SM) is the ceiling, the FP32 pipe
// runs at 7/8 of it.  The loop is software-pipelined in three stages (r2 of entry c, rsqrt of entry c-1, accumulate
// entry c-2), and the pipeline is carried across the sites of a chunk: the first two entries of site s+1 still
// accumulate into site s, whose epilogue follows them.  (16 atoms per thread — half the LDS.128 per pair — was built
// and measured: no faster in the bare loop and slower in the kernel, the per-atom epilogue constants no longer fit in
// registers.  scripts/pipe_bench5.cu, scripts/cdc_kernel_16atoms_experiment.cuh, profiles/.)
//
// Site blocks (sites x P entries of 16 B) arrive in a ring of kCdcRing shared-memory buffers, one bulk TMA copy per
// unit, with a full/empty mbarrier pair per buffer (full: the copy's complete_tx; empty: one arrival per warp once it
// has read the buffer).  There is no __syncthreads after start-up: the warps of a CTA run free.
//
// V_s(a) = sum_c dQ_c / r_ac is FP32 over <= P terms; everything after that is FP64: w = V * (33.2f * q_a), summed per
// residue group in a fixed order inside a thread / warp; the per-thread / per-warp partials are rounded to a fixed grid of
// 2^-36 kcal/mol before they are added to the FP64 bins, which makes those adds exact and therefore order-independent
// (bin_add below): the rows are bit-reproducible although the bins are filled by atomics (SURVEY.md section 7 asks for a
// reproducible scheme).  Atoms of a group are consecutive in the file, so a thread's 8 atoms form a few runs at most:
//   * warp entirely inside one environment group (the solvent): each lane parks its FP64 partial in shared memory; the
//     32 lane values of every site of the chunk are added after the chunk (4 lanes per site, 8 values each, two shuffle
//     steps) — no shuffle tree and no dependent chain inside the site loop;
//   * otherwise: in-thread runs, one fire-and-forget grid-rounded FP64 reduction (RED.E.ADD.F64) per run (the own-site column is skipped here,
//     reference src/grom2atomFMO.cpp:674).  The 8 column indices are re-read from L1 here instead of living in
//     registers across the pair loop: the loop sits at the 128-register cap, and freeing 8 was worth 0.7 % (freeing all
//     24 epilogue registers is worth 3 %: the no-epilogue variant; re-reading 33.2f*q as well costs more than it frees).
//
// Template knobs (chosen by measurement, scripts/k2_variants.cu; logs under profiles/):
//   kMinBlocks  CTAs per SM promised to ptxas (register cap 65536 / (256 * kMinBlocks))
//   kUnroll     site atoms per trip of the main loop; the host pads P so that kUnroll divides P - 2
//   kFlags      measurement only: bit 0 = skip the epilogue, bit 1 = integer-pipe float->double instead of F2F
#pragma once
#include "kernels.cuh"
#include "ptx_helpers.cuh"

namespace cgc {

constexpr int kCdcMinBlocks = 2;
constexpr int kCdcMaxChunkSites = 8;   // site rows one warp reduces per chunk (4 lanes per row)
constexpr int kCdcWarps = kTileThreads / 32;
constexpr int kCdcRing = 4;            // site-block buffers in flight
constexpr double kCdcBinGridShift = 98304.0;   // 1.5 * 2^16: adding it puts the unit in the last place at 2^-36

// One partial sum -> its bin, reproducibly.  The partial is first rounded to a multiple of 2^-36 kcal/mol — (w + C) - C with
// C = 1.5 * 2^16, two FP64 adds, no branch, no conversion instruction (F2I would run on the XU pipe the pair loop saturates) —
// and then added with a fire-and-forget FP64 reduction (RED.E.ADD.F64, the add happens in L2).  Every addend and every
// partial sum of a bin is a multiple of 2^-36, so while |bin| < 2^17 = 131072 kcal/mol each add is EXACT and the bin does
// not depend on the order in which the reductions land: the bins, and the float32 rows made from them, are bit-identical
// from run to run and for any cut of the frames into launches.  (Beyond 2^17 the adds round again: nothing breaks, only the
// guarantee lapses.  Infinity and NaN from coincident atoms propagate as they do in the reference.)  One add is off by at
// most 2^-37 = 7.3e-12 kcal/mol; a bin takes a few hundred adds: < 1e-12 of its sum of |terms|, five orders of magnitude
// under the float32 rounding of the row.
__device__ __forceinline__ void bin_add(double* bin, double w) {
  atomicAdd(bin, __dadd_rn(__dadd_rn(w, kCdcBinGridShift), -kCdcBinGridShift));
}

// sa = (-xc, -yc, -zc, dQ) of one site atom.  pack2(v, v) is folded by ptxas into the scalar-broadcast operand form of
// the packed instruction (SASS "FADD2 R, R.F32x2.HI_LO, R.F32"), so the broadcast costs no instruction.
__device__ __forceinline__ f32x2 pair_r2(f32x2 x, f32x2 y, f32x2 z, const float4& sa) {
  const f32x2 dx = fadd2(x, pack2(sa.x, sa.x)), dy = fadd2(y, pack2(sa.y, sa.y)), dz = fadd2(z, pack2(sa.z, sa.z));
  f32x2 r2 = fmul2(dx, dx);
  r2 = ffma2(dy, dy, r2);
  return ffma2(dz, dz, r2);
}

__device__ __forceinline__ f32x2 rsqrt2(f32x2 r2) {
  float a, b;
  unpack2(r2, a, b);
  return pack2(rsqrt_approx(a), rsqrt_approx(b));
}

// Measurement only (kFlags bit 1): float -> double on the integer pipe instead of F2F.F64.F32, which runs on the XU pipe
// this kernel saturates.  Exact for normal numbers.  It is NOT faster (0.851 vs 0.855 of the ceiling: the extra integer
// instructions and registers cost more than the 2 % of XU time the 8 conversions per site take), so the product uses cvt.
template <bool kBits>
__device__ __forceinline__ double f2d(float f) {
  if constexpr (kBits) {
    const unsigned b = __float_as_uint(f);
    const unsigned hi = (((b & 0x7fffffffu) >> 3) + 0x38000000u) | (b & 0x80000000u);
    return __hiloint2double((int)hi, (int)(b << 29));
  } else {
    return (double)f;
  }
}

// dynamic shared memory of one CTA: [ring of site blocks][per-warp lane partials: warps x 32 lanes x (chunk_sites + 1)]
__host__ __device__ inline int cdc_red_stride(int chunk_sites) { return chunk_sites + 1; }
__host__ __device__ inline size_t cdc_smem_layout(int chunk_sites, int P, size_t* red_off) {
  size_t off = (size_t)kCdcRing * chunk_sites * P * 16;
  off = (off + 15) & ~(size_t)15;
  if (red_off) *red_off = off;
  off += (size_t)kCdcWarps * 32 * cdc_red_stride(chunk_sites) * sizeof(double);
  return off;
}

template <int kMinBlocks, int kUnroll, int kFlags = 0>
__global__ void __launch_bounds__(kTileThreads, kMinBlocks)
cdc_kernel(const float* __restrict__ xyz8, const float4* __restrict__ sites, const double* __restrict__ kq_all,
           const int32_t* __restrict__ col_all, double* __restrict__ acc, int n_pad, int tiles_per_frame,
           long long n_units, int S, int P, int G, int chunk_sites, int chunks_per_item) {
  static_assert(kAtomsPerThread == 8, "the loads below are written for 8 atoms per thread");
  constexpr bool kBits = (kFlags & 2) != 0;
  extern __shared__ __align__(128) unsigned char sm_raw[];
  __shared__ __align__(8) uint64_t full_bar[kCdcRing], empty_bar[kCdcRing];
  const uint32_t buf_bytes = (uint32_t)chunk_sites * P * 16u;
  size_t red_off;
  cdc_smem_layout(chunk_sites, P, &red_off);
  const int red_stride = cdc_red_stride(chunk_sites);
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  double* red_w = reinterpret_cast<double*>(sm_raw + red_off) + (size_t)warp * 32 * red_stride;  // [32 lanes][red_stride]

  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < kCdcRing; i++) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], kCdcWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();

  // this CTA's contiguous range of units; unit -> (item = frame * tiles + tile, chunk), chunk fastest
  const long long u_lo = n_units * blockIdx.x / gridDim.x, u_hi = n_units * (blockIdx.x + 1) / gridDim.x;
  if (u_lo >= u_hi) return;

  // ---- producer (thread 0 only): kCdcRing-1 units ahead of its own compute
  long long p_u = u_lo;
  auto produce_one = [&]() {   // issue the copy for unit p_u, if the range has one
    if (p_u >= u_hi) return;
    const long long item = p_u / chunks_per_item;
    const int k = (int)(p_u - item * chunks_per_item);
    const int f = (int)(item / tiles_per_frame);
    const int s0 = k * chunk_sites;
    const uint32_t seq = (uint32_t)(p_u - u_lo), b = seq % kCdcRing, use = seq / kCdcRing;
    if (use > 0) mbar_wait(&empty_bar[b], (use - 1) & 1u);   // every warp has read the buffer's previous contents
    const uint32_t bytes = (uint32_t)min(chunk_sites, S - s0) * P * 16u;
    mbar_expect_tx(&full_bar[b], bytes);
    tma_bulk_g2s(sm_raw + b * buf_bytes, sites + ((size_t)f * S + s0) * P, bytes, &full_bar[b]);
    p_u++;
  };
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < kCdcRing - 1; i++) produce_one();
  }

  // item state, reloaded when the range crosses into the next (frame, tile)
  f32x2 X[4], Y[4], Z[4];
  double kq[8];
  const int32_t* col_ptr = col_all;   // this thread's 8 columns (re-read from L1 by the mixed-group epilogue only)
  unsigned run_mask = 0;              // bit e: atom e ends a run of equal, real columns; bit 8+e: atom e starts a run
  bool warp_uniform = false;
  int c_lane0 = -1;
  double* acc_frame = acc;
  long long cur_item = -1;

  uint32_t seq = 0;
  for (long long u = u_lo; u < u_hi; u++, seq++) {
    const long long item = u / chunks_per_item;
    const int k = (int)(u - item * chunks_per_item);
    if (item != cur_item) {
      cur_item = item;
      const int f = (int)(item / tiles_per_frame);
      const int tile = (int)(item - (long long)f * tiles_per_frame);
      const int a0 = tile * kTileAtoms + threadIdx.x * kAtomsPerThread;
      {
        const ulonglong2* ap = reinterpret_cast<const ulonglong2*>(xyz8 + ((size_t)f * n_pad + a0) * 3);
#pragma unroll
        for (int j = 0; j < 2; j++) {
          const ulonglong2 vx = __ldg(ap + j), vy = __ldg(ap + 2 + j), vz = __ldg(ap + 4 + j);
          X[2 * j] = vx.x; X[2 * j + 1] = vx.y;
          Y[2 * j] = vy.x; Y[2 * j + 1] = vy.y;
          Z[2 * j] = vz.x; Z[2 * j + 1] = vz.y;
        }
      }
      col_ptr = col_all + a0;
      const int4 c0 = __ldg(reinterpret_cast<const int4*>(col_ptr));
      const int4 c1 = __ldg(reinterpret_cast<const int4*>(col_ptr + 4));
      {
        const double2* kp = reinterpret_cast<const double2*>(kq_all + a0);
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const double2 v = __ldg(kp + j);
          kq[2 * j] = v.x;
          kq[2 * j + 1] = v.y;
        }
      }
      const bool same = c0.y == c0.x && c0.z == c0.x && c0.w == c0.x && c1.x == c0.x && c1.y == c0.x && c1.z == c0.x &&
                        c1.w == c0.x;
      {
        const int ce[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
        run_mask = 1u << 8;
#pragma unroll
        for (int e = 0; e < 8; e++) {
          const bool end = e == 7 || ce[e + 1] != ce[e];
          if (end && ce[e] >= 0) run_mask |= 1u << e;
          if (end && e < 7) run_mask |= 1u << (9 + e);
        }
      }
      c_lane0 = __shfl_sync(0xffffffffu, c0.x, 0);
      // whole warp inside one environment group (columns < 0 are padding)
      warp_uniform = __all_sync(0xffffffffu, same && c0.x == c_lane0) && c_lane0 >= 0;
      acc_frame = acc + (size_t)f * S * G;
    }

    // FP64 from here: w_e = V_e * 33.2f * q_e, binned per column (sl: the site's position in the chunk)
    auto epilogue = [&](const f32x2 (&V)[4], int s, int sl) {
      float Vf[8];
#pragma unroll
      for (int j = 0; j < 4; j++) unpack2(V[j], Vf[2 * j], Vf[2 * j + 1]);
      double* acc_row = acc_frame + (size_t)s * G;
      if constexpr ((kFlags & 1) != 0) {
        float t = 0.f;
#pragma unroll
        for (int e = 0; e < 8; e++) t += Vf[e];
        if (t == 123.456f) bin_add(acc_row, (double)t);
      } else if (warp_uniform) {
        // two independent chains of four
        double t0 = f2d<kBits>(Vf[0]) * kq[0], t1 = f2d<kBits>(Vf[1]) * kq[1];
#pragma unroll
        for (int e = 2; e < 8; e += 2) {
          t0 += f2d<kBits>(Vf[e]) * kq[e];
          t1 += f2d<kBits>(Vf[e + 1]) * kq[e + 1];
        }
        red_w[lane * red_stride + sl] = t0 + t1;   // summed over the lanes after the chunk
      } else {
        // the columns are not kept in registers across the pair loop (they would cost 8 of its 128): L1 has them
        const int4 c0 = __ldg(reinterpret_cast<const int4*>(col_ptr));
        const int4 c1 = __ldg(reinterpret_cast<const int4*>(col_ptr + 4));
        const int cols[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
        // branch-free segmented sum: the run structure is a per-item bit mask, the flush a predicated reduction
        double run = 0.0;
#pragma unroll
        for (int e = 0; e < 8; e++) {
          const double w = f2d<kBits>(Vf[e]) * kq[e];
          run = ((run_mask >> (8 + e)) & 1u) ? w : run + w;
          // own-site column: reference src/grom2atomFMO.cpp:674
          if (((run_mask >> e) & 1u) && cols[e] != s) bin_add(acc_row + cols[e], run);
        }
      }
    };

    const int s_begin = k * chunk_sites;
    const int n_s = min(chunk_sites, S - s_begin);
    if (threadIdx.x == 0) produce_one();
    if (k == chunks_per_item - 1 && u + 1 < u_hi) {
      // the next item's 96 bytes of coordinates -> L1, so that its first loads do not stall on L2/DRAM
      const long long nitem = item + 1;
      const int nf = (int)(nitem / tiles_per_frame);
      const int nt = (int)(nitem - (long long)nf * tiles_per_frame);
      const char* np = reinterpret_cast<const char*>(
          xyz8 + ((size_t)nf * n_pad + nt * kTileAtoms + threadIdx.x * kAtomsPerThread) * 3);
      asm volatile("prefetch.global.L1 [%0];" ::"l"(np));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(np + 32));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(np + 64));
    }
    const uint32_t b = seq % kCdcRing;
    mbar_wait(&full_bar[b], (seq / kCdcRing) & 1u);
    const float4* sblk = reinterpret_cast<const float4*>(sm_raw + b * buf_bytes);

    // pipeline registers, carried across the sites of the chunk.  R = 1, dq = 0: the first two accumulates add 0.
    f32x2 V[4], R[4], RI[4];
    float dq1 = 0.f, dq2 = 0.f;
#pragma unroll
    for (int j = 0; j < 4; j++) { V[j] = 0ull; R[j] = pack2(1.f, 1.f); RI[j] = R[j]; }
    auto body = [&](const float4 sa) {   // sa: one broadcast LDS.128 (-xc, -yc, -zc, dQ)
#pragma unroll
      for (int j = 0; j < 4; j++) {
        V[j] = ffma2(pack2(dq2, dq2), RI[j], V[j]);      // entry c-2
        RI[j] = rsqrt2(R[j]);                            // entry c-1
        R[j] = pair_r2(X[j], Y[j], Z[j], sa);            // entry c
      }
      dq2 = dq1;
      dq1 = sa.w;
    };
    const int n_main = (P - 2) / kUnroll;   // the host pads P: P >= 2 and kUnroll divides P - 2
    for (int sl = 0; sl < n_s; sl++) {
      const float4* sp = sblk + (size_t)sl * P;
      body(sp[0]);   // these two still accumulate the previous site's last entries
      body(sp[1]);
      if (sl > 0) epilogue(V, s_begin + sl - 1, sl - 1);
#pragma unroll
      for (int j = 0; j < 4; j++) V[j] = 0ull;
      sp += 2;
#pragma unroll 1
      for (int m = 0; m < n_main; m++, sp += kUnroll) {
#pragma unroll
        for (int c = 0; c < kUnroll; c++) body(sp[c]);
      }
    }
    // all reads of the buffer are done (their values are in registers): hand it back to the producer
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[b]);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      V[j] = ffma2(pack2(dq2, dq2), RI[j], V[j]);
      V[j] = ffma2(pack2(dq1, dq1), rsqrt2(R[j]), V[j]);
    }
    epilogue(V, s_begin + n_s - 1, n_s - 1);

    if constexpr ((kFlags & 1) == 0) {
      if (warp_uniform) {
        // lane partials of this warp: 4 lanes per site row, 8 lane values each, then two shuffle steps
        __syncwarp();
        const int r = lane >> 2, qd = lane & 3;
        double t = 0.0;
        if (r < n_s) {
          const double* p = red_w + (size_t)(qd * 8) * red_stride + r;
#pragma unroll
          for (int l = 0; l < 8; l++) t += p[(size_t)l * red_stride];
        }
        t += __shfl_xor_sync(0xffffffffu, t, 1);
        t += __shfl_xor_sync(0xffffffffu, t, 2);
        if (qd == 0 && r < n_s && c_lane0 != s_begin + r) bin_add(acc_frame + (size_t)(s_begin + r) * G + c_lane0, t);
        __syncwarp();   // the rows are rewritten by the next chunk's epilogues
      }
    }
  }
}

}  // namespace cgc

// EXPERIMENT RECORD (not built): the pair kernel with 16 atoms per thread, 128-thread CTAs, 4 per SM, and the
// epilogue constants in a shared-memory stash.  Needs kTileThreads = 128, kAtomsPerThread = 16 in kernels.cuh.
// Measured 0.775 of the ceiling (0.859 without epilogue) against 0.85 for the 8-atom kernel: profiles/r01_k2_16atoms.log

// K2 — the CDC pair kernel (replaces ComputeCDC_v1, reference src/grom2atomFMO.cpp:625-739), sm_100a.
//
// Work item = (frame, tile of kTileAtoms = 2048 consecutive atoms, group of sites).  Persistent CTAs of 4 warps (4 CTAs
// per SM) stride over the items.  A group is all sites when the launch is large and shrinks down to 2 sites for small
// launches, so that even a single frame fills the machine.
//
// Each thread keeps 16 consecutive atoms in registers as 8 packed pairs (two blocks of the xyz8 layout K1 writes, loaded
// as 64-bit values).  For every site s and site atom c, all threads read the same 16 B from shared memory (one broadcast
// LDS.128: -xc, -yc, -zc, dQ) and evaluate 16 pairs:
//   dx = x - xc (FADD2 x3), r2 = dx*dx + dy*dy + dz*dz (FMUL2 + 2 FFMA2), rinv = MUFU.RSQ (x2), V += dQ*rinv (FFMA2)
// i.e. 7 packed FP32 instructions + 2 MUFU per 2 pairs: the XU pipe (16 rsqrt/clk/SM) is the ceiling, the FP32 pipe runs
// at 7/8 of it.  MUFU and LDS share the MIO queue, and the broadcast load turned out to be what costs most next to the
// MUFU stream: the bare loop reaches 0.98 of the ceiling without loads, 0.91 with one LDS.128 per 8 pairs, 0.95 per 12
// and 0.97 per 16 (scripts/pipe_bench5.cu, profiles/).  Hence 16 atoms per thread, which leaves no registers for the
// per-atom constants of the epilogue: 33.2f*q (FP64) and the dE column of each atom are parked in a thread-private,
// conflict-free shared-memory stash once per item and read back once per site.
//
// The loop is software-pipelined in three stages (r2 of entry c, rsqrt of entry c-1, accumulate entry c-2) and the
// pipeline is carried across the sites of a chunk: the first two entries of site s+1 still accumulate into site s.
//
// Site blocks (sites x P entries of 16 B) arrive in a ring of kCdcRing shared-memory buffers, one bulk TMA copy per
// chunk of <= 8 sites, with a full/empty mbarrier pair per buffer (full: the copy's complete_tx; empty: one arrival per
// warp once it has read the buffer).  There is no __syncthreads after start-up: the warps of a CTA run free.
//
// V_s(a) = sum_c dQ_c / r_ac is FP32 over <= P terms; everything after that is FP64: w = V * (33.2f * q_a), summed per
// residue group.  Atoms of a group are consecutive in the file, so a thread's 16 atoms form a few runs at most:
//   * warp entirely inside one environment group (the solvent): each lane parks its FP64 partial in shared memory; the
//     32 lane values of every site of the chunk are added after the chunk (4 lanes per site, 8 values each, two shuffle
//     steps) — no shuffle tree and no dependent chain inside the site loop;
//   * otherwise: in-thread runs, one fire-and-forget atomicAdd(double) per run (the own-site column is skipped here,
//     reference src/grom2atomFMO.cpp:674).
//
// kFlags (measurement only): bit 0 = skip the epilogue.
#pragma once
#include "kernels.cuh"
#include "ptx_helpers.cuh"

namespace cgc {

constexpr int kCdcMinBlocks = 4;
constexpr int kCdcMaxChunkSites = 8;   // site rows one warp reduces per chunk (4 lanes per row)
constexpr int kCdcWarps = kTileThreads / 32;
constexpr int kCdcRing = 3;            // site-block buffers in flight
constexpr int kNP = kAtomsPerThread / 2;   // packed pairs per thread

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

// dynamic shared memory of one CTA: [ring of site blocks][stash: kq double2 [8][threads], col int4 [4][threads]]
// [per-warp lane partials: warps x 32 lanes x (chunk_sites + 1) doubles]
__host__ __device__ inline int cdc_red_stride(int chunk_sites) { return chunk_sites + 1; }
__host__ __device__ inline size_t cdc_smem_layout(int chunk_sites, int P, size_t* stash_off, size_t* red_off) {
  size_t off = (size_t)kCdcRing * chunk_sites * P * 16;
  off = (off + 15) & ~(size_t)15;
  if (stash_off) *stash_off = off;
  off += (size_t)kTileThreads * kAtomsPerThread * (sizeof(double) + sizeof(int32_t));
  if (red_off) *red_off = off;
  off += (size_t)kCdcWarps * 32 * cdc_red_stride(chunk_sites) * sizeof(double);
  return off;
}

template <int kMinBlocks, int kUnroll, int kFlags = 0>
__global__ void __launch_bounds__(kTileThreads, kMinBlocks)
cdc_kernel(const float* __restrict__ xyz8, const float4* __restrict__ sites, const double* __restrict__ kq_all,
           const int32_t* __restrict__ col_all, double* __restrict__ acc, int n_pad, int tiles_per_frame, int n_items,
           int S, int P, int G, int n_groups, int group_sites, int chunk_sites) {
  static_assert(kAtomsPerThread == 16, "the stash layout and the loads below are written for 16 atoms per thread");
  extern __shared__ __align__(128) unsigned char sm_raw[];
  __shared__ __align__(8) uint64_t full_bar[kCdcRing], empty_bar[kCdcRing];
  const uint32_t buf_bytes = (uint32_t)chunk_sites * P * 16u;
  size_t stash_off, red_off;
  cdc_smem_layout(chunk_sites, P, &stash_off, &red_off);
  const int red_stride = cdc_red_stride(chunk_sites);
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  // stash, indexed [piece][thread] so that the 8 threads of a quarter-warp hit 8 different 16-byte bank groups
  double2* kq_st = reinterpret_cast<double2*>(sm_raw + stash_off) + threadIdx.x;                 // [8][kTileThreads]
  int4* col_st = reinterpret_cast<int4*>(sm_raw + stash_off + (size_t)kTileThreads * 128) + threadIdx.x;  // [4][kTileThreads]
  double* red_w = reinterpret_cast<double*>(sm_raw + red_off) + (size_t)warp * 32 * red_stride;  // [32 lanes][red_stride]

  // item -> (frame, tile, group of sites); the group index varies fastest so that neighbouring CTAs share atoms in L2
  auto decode_item = [&](int it, int& f, int& tile, int& s_lo, int& s_hi) {
    const int ft = it / n_groups;
    const int g = it - ft * n_groups;
    f = ft / tiles_per_frame;
    tile = ft - f * tiles_per_frame;
    s_lo = g * group_sites;
    s_hi = min(S, s_lo + group_sites);
  };

  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < kCdcRing; i++) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], kCdcWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();
  if ((int)blockIdx.x >= n_items) return;

  // ---- producer (thread 0 only): walks the CTA's (item, chunk) sequence kCdcRing-1 entries ahead of its own compute
  int p_item = blockIdx.x, p_k = 0;
  uint32_t p_seq = 0;
  auto produce_one = [&]() {   // issue the copy for sequence entry p_seq, if the sequence has one
    if (p_item >= n_items) return;
    int f, tile, s_lo, s_hi;
    decode_item(p_item, f, tile, s_lo, s_hi);
    const int s0 = s_lo + p_k * chunk_sites;
    const uint32_t b = p_seq % kCdcRing, use = p_seq / kCdcRing;
    if (use > 0) mbar_wait(&empty_bar[b], (use - 1) & 1u);   // every warp has read the buffer's previous contents
    const uint32_t bytes = (uint32_t)min(chunk_sites, s_hi - s0) * P * 16u;
    mbar_expect_tx(&full_bar[b], bytes);
    tma_bulk_g2s(sm_raw + b * buf_bytes, sites + ((size_t)f * S + s0) * P, bytes, &full_bar[b]);
    p_seq++;
    if (s0 + chunk_sites < s_hi) p_k++;
    else { p_k = 0; p_item += gridDim.x; }
  };
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < kCdcRing - 1; i++) produce_one();
  }

  uint32_t seq = 0;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    int f, tile, s_lo, s_hi;
    decode_item(item, f, tile, s_lo, s_hi);
    const int n_chunks = (s_hi - s_lo + chunk_sites - 1) / chunk_sites;
    const int a0 = tile * kTileAtoms + threadIdx.x * kAtomsPerThread;

    // ---- per-atom constants of the epilogue -> stash (thread-private slots: no synchronisation needed)
    bool same = true;
    int col0;
    {
      const int4* cp = reinterpret_cast<const int4*>(col_all + a0);
      int4 c[4];
#pragma unroll
      for (int j = 0; j < 4; j++) c[j] = __ldg(cp + j);
      col0 = c[0].x;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        same = same && c[j].x == col0 && c[j].y == col0 && c[j].z == col0 && c[j].w == col0;
        col_st[j * kTileThreads] = c[j];
      }
      const double2* kp = reinterpret_cast<const double2*>(kq_all + a0);
#pragma unroll
      for (int j = 0; j < 8; j++) kq_st[j * kTileThreads] = __ldg(kp + j);
    }
    const int c_lane0 = __shfl_sync(0xffffffffu, col0, 0);
    // whole warp inside one environment group (columns < 0 are padding)
    const bool warp_uniform = __all_sync(0xffffffffu, same && col0 == c_lane0) && c_lane0 >= 0;

    // ---- this thread's 16 consecutive atoms (two blocks of the xyz8 layout)
    f32x2 X[kNP], Y[kNP], Z[kNP];
    {
      const ulonglong2* ap = reinterpret_cast<const ulonglong2*>(xyz8 + ((size_t)f * n_pad + a0) * 3);
#pragma unroll
      for (int blk = 0; blk < 2; blk++)
#pragma unroll
        for (int j = 0; j < 2; j++) {
          const ulonglong2 vx = __ldg(ap + 6 * blk + j), vy = __ldg(ap + 6 * blk + 2 + j), vz = __ldg(ap + 6 * blk + 4 + j);
          X[4 * blk + 2 * j] = vx.x; X[4 * blk + 2 * j + 1] = vx.y;
          Y[4 * blk + 2 * j] = vy.x; Y[4 * blk + 2 * j + 1] = vy.y;
          Z[4 * blk + 2 * j] = vz.x; Z[4 * blk + 2 * j + 1] = vz.y;
        }
    }
    double* acc_frame = acc + (size_t)f * S * G;

    // FP64 from here: w_e = V_e * 33.2f * q_e, binned per column (sl: the site's position in the chunk)
    auto epilogue = [&](const f32x2 (&V)[kNP], int s, int sl) {
      if constexpr ((kFlags & 1) != 0) {
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < kNP; j++) { float a, b; unpack2(V[j], a, b); t += a + b; }
        if (t == 123.456f) atomicAdd(acc_frame + (size_t)s * G, (double)t);
      } else if (warp_uniform) {
        double t0 = 0.0, t1 = 0.0;   // two independent chains
#pragma unroll
        for (int j = 0; j < kNP; j++) {
          float a, b;
          unpack2(V[j], a, b);
          const double2 k2 = kq_st[j * kTileThreads];
          t0 += (double)a * k2.x;
          t1 += (double)b * k2.y;
        }
        red_w[lane * red_stride + sl] = t0 + t1;   // summed over the lanes after the chunk
      } else {
        double* acc_row = acc_frame + (size_t)s * G;
        double run = 0.0;
        int cc = col0;
#pragma unroll
        for (int j4 = 0; j4 < 4; j4++) {
          const int4 c4 = col_st[j4 * kTileThreads];
          const int ce[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
          for (int h = 0; h < 2; h++) {
            float a, b;
            unpack2(V[2 * j4 + h], a, b);
            const double2 k2 = kq_st[(2 * j4 + h) * kTileThreads];
            const float ve[2] = {a, b};
            const double ke[2] = {k2.x, k2.y};
#pragma unroll
            for (int u = 0; u < 2; u++) {
              const int c = ce[2 * h + u];
              if (c != cc) {
                if (cc >= 0 && cc != s) atomicAdd(acc_row + cc, run);   // own-site column: src/grom2atomFMO.cpp:674
                cc = c;
                run = 0.0;
              }
              run += (double)ve[u] * ke[u];
            }
          }
        }
        if (cc >= 0 && cc != s) atomicAdd(acc_row + cc, run);
      }
    };

    for (int k = 0; k < n_chunks; k++, seq++) {
      const int s_begin = s_lo + k * chunk_sites;
      const int n_s = min(chunk_sites, s_hi - s_begin);
      if (threadIdx.x == 0) produce_one();
      if (k == n_chunks - 1 && item + (int)gridDim.x < n_items) {
        // the next item's 192 bytes of coordinates -> L1, so that its first loads do not stall on L2/DRAM
        int nf, nt, nlo, nhi;
        decode_item(item + gridDim.x, nf, nt, nlo, nhi);
        const char* np = reinterpret_cast<const char*>(
            xyz8 + ((size_t)nf * n_pad + nt * kTileAtoms + threadIdx.x * kAtomsPerThread) * 3);
#pragma unroll
        for (int i = 0; i < 6; i++) asm volatile("prefetch.global.L1 [%0];" ::"l"(np + 32 * i));
      }
      const uint32_t b = seq % kCdcRing;
      mbar_wait(&full_bar[b], (seq / kCdcRing) & 1u);
      const float4* sblk = reinterpret_cast<const float4*>(sm_raw + b * buf_bytes);

      // pipeline registers, carried across the sites of the chunk.  R = 1, dq = 0: the first two accumulates add 0.
      f32x2 V[kNP], R[kNP], RI[kNP];
      float dq1 = 0.f, dq2 = 0.f;
#pragma unroll
      for (int j = 0; j < kNP; j++) { V[j] = 0ull; R[j] = pack2(1.f, 1.f); RI[j] = R[j]; }
      auto body = [&](const float4 sa) {   // sa: one broadcast LDS.128 (-xc, -yc, -zc, dQ)
#pragma unroll
        for (int j = 0; j < kNP; j++) {
          V[j] = ffma2(pack2(dq2, dq2), RI[j], V[j]);      // entry c-2
          RI[j] = rsqrt2(R[j]);                            // entry c-1
          R[j] = pair_r2(X[j], Y[j], Z[j], sa);            // entry c
        }
        dq2 = dq1;
        dq1 = sa.w;
      };
      const int n_main = (P - 2) / kUnroll;   // the host pads P so that the remainder is short
      for (int sl = 0; sl < n_s; sl++) {
        const float4* sp = sblk + (size_t)sl * P;
        body(sp[0]);   // P >= 2 (the host pads): these two still accumulate the previous site's last entries
        body(sp[1]);
        if (sl > 0) epilogue(V, s_begin + sl - 1, sl - 1);
#pragma unroll
        for (int j = 0; j < kNP; j++) V[j] = 0ull;
        sp += 2;
#pragma unroll 1
        for (int m = 0; m < n_main; m++, sp += kUnroll) {
#pragma unroll
          for (int c = 0; c < kUnroll; c++) body(sp[c]);
        }
#pragma unroll 1
        for (int c = 2 + n_main * kUnroll; c < P; c++, sp++) body(sp[0]);
      }
      // all reads of the buffer are done (their values are in registers): hand it back to the producer
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[b]);
#pragma unroll
      for (int j = 0; j < kNP; j++) {
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
          if (qd == 0 && r < n_s && c_lane0 != s_begin + r) atomicAdd(acc_frame + (size_t)(s_begin + r) * G + c_lane0, t);
          __syncwarp();   // the rows are rewritten by the next chunk's epilogues
        }
      }
    }
  }
}

}  // namespace cgc

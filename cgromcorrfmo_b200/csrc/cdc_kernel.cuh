// K2 — the CDC pair kernel (replaces ComputeCDC_v1, reference src/grom2atomFMO.cpp:625-739), sm_100a.
//
// Work item = (frame, tile of kTileAtoms consecutive atoms, group of sites).  Persistent CTAs stride over the items.  A
// group is all sites when the launch is large (atoms, columns and charges are then loaded once per frame and tile) and
// shrinks down to 2 sites for small launches, so that even a single frame fills the machine.  Each thread keeps 8
// consecutive atoms in registers as 4 packed pairs (loaded as 64-bit values straight from the blocked SoA layout K1
// writes).  Inside an item the sites pass through shared memory in chunks of <= 8 sites: a chunk's site block (sites x P
// entries of 16 B: -x,-y,-z,dQ) arrives by one bulk TMA copy, double-buffered so the next chunk (of this or the next
// item) lands while this one is computed.  For every site s and site atom c, all threads read the same 16 B (one broadcast LDS.128) and
// evaluate 8 pairs:
//   dx = x - xc (FADD2 x3), r2 = dx*dx + dy*dy + dz*dz (FMUL2 + 2 FFMA2), rinv = MUFU.RSQ (x2), V += dQ*rinv (FFMA2)
// i.e. 7 packed FP32 instructions + 2 MUFU per 2 pairs: the XU pipe (16 rsqrt/clk/SM) is the ceiling, the FP32 pipe
// runs at 7/8 of it.  The loop is software-pipelined in three stages (r2 of atom c, rsqrt of atom c-1, accumulate
// atom c-2) so that ptxas can spread the MUFUs between the packed FP32 instructions.
// V_s(a) = sum_c dQ_c / r_ac is FP32 over <= P terms; everything after that is FP64: w = V * (33.2f * q_a), summed per
// residue group.  Atoms of a group are consecutive in the file, so a thread's 8 atoms form a few runs at most:
//   * warp entirely inside one environment group (the solvent): each lane parks its FP64 partial in shared memory and
//     moves on; the 32 lane values of every (warp, site) are added after the item's site loop (no shuffle tree, and no
//     dependent chain, in the site loop);
//   * otherwise: in-thread runs, one fire-and-forget atomicAdd(double) per run (the own-site column is skipped here,
//     reference src/grom2atomFMO.cpp:674).
//
// Template knobs (chosen by measurement, scripts/k2_variants.cu; logs under profiles/):
//   kMinBlocks  CTAs per SM promised to ptxas (register cap 65536 / (256 * kMinBlocks))
//   kUnroll     unroll factor of the site-atom loop
//   kStages     2 or 3 software-pipeline stages
//   kFlags      measurement only: bit 0 = skip the epilogue, bit 1 = warp-shuffle tree instead of the deferred reduction
#pragma once
#include "kernels.cuh"
#include "ptx_helpers.cuh"

namespace cgc {

constexpr int kCdcMinBlocks = 2;
constexpr int kCdcUnroll = 21;  // divides P - 2 = 42 for the 44 dQ atoms of a BCL; other P values take the remainder loop
constexpr int kCdcStages = 3;
constexpr int kCdcMaxChunkSites = 8;   // bounds the shared memory of the deferred reduction
constexpr int kCdcWarps = kTileThreads / 32;

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

// dynamic shared memory of one CTA: two site-block buffers + two reduction buffers + per-warp columns
__host__ __device__ inline int cdc_red_stride(int chunk_sites) { return kCdcWarps * chunk_sites + 1; }
__host__ __device__ inline size_t cdc_smem_layout(int chunk_sites, int P, size_t* red_off, size_t* wcol_off) {
  size_t off = (size_t)2 * chunk_sites * P * 16;
  off = (off + 15) & ~(size_t)15;
  if (red_off) *red_off = off;
  off += (size_t)2 * 32 * cdc_red_stride(chunk_sites) * sizeof(double);
  if (wcol_off) *wcol_off = off;
  off += (size_t)2 * kCdcWarps * sizeof(int);
  return off;
}

template <int kMinBlocks, int kUnroll, int kStages, int kFlags = 0>
__global__ void __launch_bounds__(kTileThreads, kMinBlocks)
cdc_kernel(const float* __restrict__ xyz8, const float4* __restrict__ sites, const double* __restrict__ kq_all,
           const int32_t* __restrict__ col_all, double* __restrict__ acc, int n_pad, int tiles_per_frame, int n_items,
           int S, int P, int G, int n_groups, int group_sites, int chunk_sites) {
  extern __shared__ __align__(128) unsigned char sm_raw[];
  __shared__ __align__(8) uint64_t bars[2];
  const uint32_t buf_bytes = (uint32_t)chunk_sites * P * 16u;   // one buffer: the dQ atoms of up to chunk_sites sites
  size_t red_off, wcol_off;
  cdc_smem_layout(chunk_sites, P, &red_off, &wcol_off);
  double* red = reinterpret_cast<double*>(sm_raw + red_off);     // [2][32 lanes][red_stride]
  int* wcol = reinterpret_cast<int*>(sm_raw + wcol_off);         // [2][warps]: column of a single-group warp, else -1
  const int red_stride = cdc_red_stride(chunk_sites);
  // item -> (frame, tile, group of sites); the group index varies fastest so that neighbouring CTAs share atoms in L2.
  // Inside an item the group's sites are processed in shared-memory chunks of <= chunk_sites sites.
  auto decode_item = [&](int it, int& f, int& tile, int& s_lo, int& s_hi) {
    const int ft = it / n_groups;
    const int g = it - ft * n_groups;
    f = ft / tiles_per_frame;
    tile = ft - f * tiles_per_frame;
    s_lo = g * group_sites;
    s_hi = min(S, s_lo + group_sites);
  };
  // the (item, chunk) sequence this CTA walks; thread 0 keeps one TMA copy ahead of the compute
  auto issue_copy = [&](int it, int k, int buf_i) {
    int f, tile, s_lo, s_hi;
    decode_item(it, f, tile, s_lo, s_hi);
    const int s0 = s_lo + k * chunk_sites;
    const uint32_t bytes = (uint32_t)min(chunk_sites, s_hi - s0) * P * 16u;
    mbar_expect_tx(&bars[buf_i], bytes);
    tma_bulk_g2s(sm_raw + buf_i * buf_bytes, sites + ((size_t)f * S + s0) * P, bytes, &bars[buf_i]);
  };

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncthreads();

  int item = blockIdx.x;
  if (item >= n_items) return;
  if (threadIdx.x == 0) issue_copy(item, 0, 0);
  uint32_t phase_bits = 0;  // bit b = parity to wait for on bars[b]
  int cur = 0;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;

  for (; item < n_items; item += gridDim.x) {
    int f, tile, s_lo, s_hi;
    decode_item(item, f, tile, s_lo, s_hi);
    const int n_chunks = (s_hi - s_lo + chunk_sites - 1) / chunk_sites;

    // ---- this thread's 8 consecutive atoms (one block of the xyz8 layout), loaded once per item
    const int a0 = tile * kTileAtoms + threadIdx.x * kAtomsPerThread;
    f32x2 X[4], Y[4], Z[4];
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
    int cols[8];
    {
      const int4 c0 = __ldg(reinterpret_cast<const int4*>(col_all + a0));
      const int4 c1 = __ldg(reinterpret_cast<const int4*>(col_all + a0 + 4));
      cols[0] = c0.x; cols[1] = c0.y; cols[2] = c0.z; cols[3] = c0.w;
      cols[4] = c1.x; cols[5] = c1.y; cols[6] = c1.z; cols[7] = c1.w;
    }
    double kq[8];
    {
      const double2* kp = reinterpret_cast<const double2*>(kq_all + a0);
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const double2 v = __ldg(kp + j);
        kq[2 * j] = v.x;
        kq[2 * j + 1] = v.y;
      }
    }
    bool same = true;
#pragma unroll
    for (int e = 1; e < 8; e++) same = same && (cols[e] == cols[0]);
    const int c_lane0 = __shfl_sync(0xffffffffu, cols[0], 0);
    // whole warp inside one environment group (columns < 0 are padding; chromophore columns never fill a warp)
    const bool warp_uniform = __all_sync(0xffffffffu, same && cols[0] == c_lane0) && c_lane0 >= 0;
    double* acc_frame = acc + (size_t)f * S * G;

    for (int k = 0; k < n_chunks; k++) {
      const int s_begin = s_lo + k * chunk_sites;
      const int n_s = min(chunk_sites, s_hi - s_begin);
      if (threadIdx.x == 0) {
        // the copy for the next chunk of the sequence goes into buf[cur^1], last read one chunk ago; the __syncthreads
        // that ended that chunk orders those reads before this write
        if (k + 1 < n_chunks) issue_copy(item, k + 1, cur ^ 1);
        else if (item + (int)gridDim.x < n_items) issue_copy(item + gridDim.x, 0, cur ^ 1);
      }
      double* red_cur = red + (size_t)cur * 32 * red_stride;
      if (lane == 0) wcol[cur * kCdcWarps + warp] = warp_uniform ? c_lane0 : -1;

      mbar_wait(&bars[cur], (phase_bits >> cur) & 1u);
      phase_bits ^= 1u << cur;
      const float4* sblk = reinterpret_cast<const float4*>(sm_raw + cur * buf_bytes);

      for (int sl = 0; sl < n_s; sl++) {
        const int s = s_begin + sl;
        f32x2 V[4];
#pragma unroll
        for (int j = 0; j < 4; j++) V[j] = 0ull;
        const float4* sp = sblk + (size_t)sl * P;
        if constexpr (kStages == 3) {
          // three stages: r2 of atom c, rsqrt of atom c-1, accumulate atom c-2   (pads have dQ = 0)
          f32x2 R[4], RI[4];
          float dq1, dq2;
          {
            const float4 s0 = sp[0], s1 = sp[P > 1 ? 1 : 0];
#pragma unroll
            for (int j = 0; j < 4; j++) RI[j] = rsqrt2(pair_r2(X[j], Y[j], Z[j], s0));
#pragma unroll
            for (int j = 0; j < 4; j++) R[j] = pair_r2(X[j], Y[j], Z[j], s1);
            dq2 = s0.w;
            dq1 = P > 1 ? s1.w : 0.f;
          }
#pragma unroll kUnroll
          for (int c = 2; c < P; c++) {
            const float4 sa = sp[c];  // one broadcast LDS.128: (-xc, -yc, -zc, dQ)
#pragma unroll
            for (int j = 0; j < 4; j++) {
              V[j] = ffma2(pack2(dq2, dq2), RI[j], V[j]);      // atom c-2
              RI[j] = rsqrt2(R[j]);                            // atom c-1
              R[j] = pair_r2(X[j], Y[j], Z[j], sa);            // atom c
            }
            dq2 = dq1;
            dq1 = sa.w;
          }
#pragma unroll
          for (int j = 0; j < 4; j++) {
            V[j] = ffma2(pack2(dq2, dq2), RI[j], V[j]);
            V[j] = ffma2(pack2(dq1, dq1), rsqrt2(R[j]), V[j]);
          }
        } else {
          f32x2 R[4];
          float dq_prev;
          {
            const float4 sa = sp[0];
#pragma unroll
            for (int j = 0; j < 4; j++) R[j] = pair_r2(X[j], Y[j], Z[j], sa);
            dq_prev = sa.w;
          }
#pragma unroll kUnroll
          for (int c = 1; c < P; c++) {
            const float4 sa = sp[c];
#pragma unroll
            for (int j = 0; j < 4; j++) {
              const f32x2 ri = rsqrt2(R[j]);                  // atom c-1
              R[j] = pair_r2(X[j], Y[j], Z[j], sa);           // atom c, independent of ri
              V[j] = ffma2(pack2(dq_prev, dq_prev), ri, V[j]);
            }
            dq_prev = sa.w;
          }
#pragma unroll
          for (int j = 0; j < 4; j++) V[j] = ffma2(pack2(dq_prev, dq_prev), rsqrt2(R[j]), V[j]);
        }
        float Vf[8];
#pragma unroll
        for (int j = 0; j < 4; j++) unpack2(V[j], Vf[2 * j], Vf[2 * j + 1]);

        // ---- FP64 from here: w_e = V_e * 33.2f * q_e, binned per column
        double* acc_row = acc_frame + (size_t)s * G;
        if constexpr ((kFlags & 1) != 0) {  // measurement only: no binning
          float t = 0.f;
#pragma unroll
          for (int e = 0; e < 8; e++) t += Vf[e];
          if (t == 123.456f) atomicAdd(acc_row, (double)t);
        } else if (warp_uniform) {
          // two independent chains of four
          double t0 = (double)Vf[0] * kq[0], t1 = (double)Vf[1] * kq[1];
#pragma unroll
          for (int e = 2; e < 8; e += 2) {
            t0 += (double)Vf[e] * kq[e];
            t1 += (double)Vf[e + 1] * kq[e + 1];
          }
          double t = t0 + t1;
          if constexpr ((kFlags & 2) != 0) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (lane == 0) atomicAdd(acc_row + c_lane0, t);
          } else {
            red_cur[lane * red_stride + warp * chunk_sites + sl] = t;   // summed over lanes after the chunk's site loop
          }
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
            run += (double)Vf[e] * kq[e];
          }
          if (cc >= 0 && cc != s) atomicAdd(acc_row + cc, run);
        }
      }
      __syncthreads();  // everyone is done with buf[cur]; all lane partials of this chunk are in red[cur]
      if constexpr ((kFlags & 3) == 0) {
        // deferred lane reduction: 4 threads per (warp, site) row, 8 lane values each, then two shuffle steps.
        // red[cur] / wcol[cur] are rewritten two chunks from now, after the next __syncthreads — which every thread
        // reaches only after it has finished this pass.
        const int rows = kCdcWarps * chunk_sites;
        for (int task = threadIdx.x; task < ((4 * rows + 31) & ~31); task += kTileThreads) {
          const int r = task >> 2, q = task & 3;
          const int w = r / chunk_sites, sl = r - w * chunk_sites;
          double t = 0.0;
          int c = -1;
          if (r < rows && sl < n_s) {
            c = wcol[cur * kCdcWarps + w];
            if (c >= 0) {
              const double* p = red_cur + (size_t)(q * 8) * red_stride + r;
#pragma unroll
              for (int l = 0; l < 8; l++) t += p[(size_t)l * red_stride];
            }
          }
          t += __shfl_xor_sync(0xffffffffu, t, 1);
          t += __shfl_xor_sync(0xffffffffu, t, 2);
          if (q == 0 && c >= 0) atomicAdd(acc_frame + (size_t)(s_begin + sl) * G + c, t);
        }
      }
      cur ^= 1;
    }
  }
}

}  // namespace cgc

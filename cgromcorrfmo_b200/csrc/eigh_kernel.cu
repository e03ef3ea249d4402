// K6 — symmetric eigen-decomposition of the per-site covariance matrices on the device.
//
// depca diagonalises every site's Ncoarse x Ncoarse covariance with numpy.linalg.eigh
// (reference scripts/depca/depca/sidechain_corr.py:91-96) before sorting the modes.  Here: one-sided (Hestenes) Jacobi,
// batched over the matrices.  W starts as A and V as I, both stored transposed (row p = column p); a rotation of the pair
// (p, q) makes the columns w_p, w_q orthogonal and is applied to v_p, v_q as well; when all columns of W = A V are
// mutually orthogonal, V holds the eigenvectors and |lambda_p| = |w_p| with the sign of v_p . w_p.
// First version (jacobi_round_kernel, kept behind CGC_EIGH_UNBLOCKED for comparison): the n/2 disjoint column pairs of a
// round-robin round are independent, one CTA per (pair, matrix), one launch per round, n - 1 rounds per sweep.  Every round
// reads and writes W and V once (32 n^2 bytes per matrix): 24 matrices of 1099^2 took 14 sweeps = 14 TB of traffic = 2.65 s,
// against 1.9 s for LAPACK on the box's 16 host cores.
// Block version (jacobi_block_round_kernel, the default): pairs of 16-column BLOCKS, the 32 x 32 Gram matrix of a pair
// rotated in shared memory — 69 rounds per sweep instead of 1098, 1/16 of the traffic, Gram and apply passes on DMMA:
// 0.55 s for the same matrices (15 sweeps, transfers included; scripts/dev_eigh_check.py), eigenvalues within 4e-13 of
// lambda_max of LAPACK's.
#include "kernels.cuh"

#include <algorithm>
#include <cstdint>
#include <cstdlib>

namespace cgc {

namespace {

constexpr int kEighThreads = 256;

__device__ __forceinline__ double block_sum3(double& a, double& b, double& c) {
  __shared__ double red[3][kEighThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[0][warp] = a; red[1][warp] = b; red[2][warp] = c; }
  __syncthreads();
  a = b = c = 0.0;
#pragma unroll
  for (int w = 0; w < kEighThreads / 32; w++) { a += red[0][w]; b += red[1][w]; c += red[2][w]; }
  return a;
}

// round-robin tournament over m (even) players: pair k of round r
__device__ __forceinline__ void tournament_pair(int m, int r, int k, int& p, int& q) {
  if (k == 0) { p = m - 1; q = r; }
  else { p = (r + k) % (m - 1); q = (r - k + (m - 1)) % (m - 1); }
}

__global__ void __launch_bounds__(kEighThreads)
jacobi_round_kernel(double* __restrict__ Wt, double* __restrict__ Vt, int n, int m, int r, double tol,
                    unsigned int* __restrict__ rotations) {
  int p, q;
  tournament_pair(m, r, blockIdx.x, p, q);
  if (p >= n || q >= n) return;   // the bye of an odd n
  if (p > q) { const int t = p; p = q; q = t; }
  const size_t base = (size_t)blockIdx.y * n * n;
  double* wp = Wt + base + (size_t)p * n;
  double* wq = Wt + base + (size_t)q * n;
  double alpha = 0.0, beta = 0.0, gamma = 0.0;
  for (int i = threadIdx.x; i < n; i += kEighThreads) {
    const double a = wp[i], b = wq[i];
    alpha += a * a;
    beta += b * b;
    gamma += a * b;
  }
  block_sum3(alpha, beta, gamma);
  if (!(fabs(gamma) > tol * sqrt(alpha * beta))) return;   // already orthogonal (or a zero column)
  const double zeta = (beta - alpha) / (2.0 * gamma);
  const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
  double* vp = Vt + base + (size_t)p * n;
  double* vq = Vt + base + (size_t)q * n;
  for (int i = threadIdx.x; i < n; i += kEighThreads) {
    const double a = wp[i], b = wq[i];
    wp[i] = c * a - s * b;
    wq[i] = s * a + c * b;
    const double x = vp[i], y = vq[i];
    vp[i] = c * x - s * y;
    vq[i] = s * x + c * y;
  }
  if (threadIdx.x == 0) atomicAdd(rotations, 1u);
}

// ---- block version: one CTA per pair of 16-column blocks -------------------------------------------------------------
// The 32 columns of a block pair are orthogonalised against each other in one visit: the CTA forms their 32 x 32 Gram
// matrix G = [Wp Wq]^T [Wp Wq] (one pass over the columns), runs ONE cyclic Jacobi sweep on G in shared memory (31 rounds of
// 16 disjoint rotations, the same rotation formula as above — a one-sided rotation of two columns IS a two-sided rotation
// of their Gram matrix), accumulating the rotations in J, and applies J to the 32 rows of Wt and Vt (a second pass).
// A round-robin over the 69 blocks of n = 1099 has 69 rounds instead of 1098, each still touching W and V once: 1/16 of the
// memory traffic; the Gram matrix and the application of J are DMMA m8n8k4 contractions.
constexpr int kJb = 16;                 // columns per block
constexpr int kJc = 2 * kJb;            // columns per CTA
constexpr int kJChunk = 64;             // elements of every column staged per step (two such tiles: the copy of the next runs under the DMMAs of this one)
constexpr int kJLd = kJChunk + 4;       // row stride of the staged tile: DMMA fragment loads (8 rows x 4 elements, or 4 x 8) hit 16 banks
constexpr int kJLdJ = kJc + 4;          // row stride of J, for the same reason
constexpr int kJThreads = 256;
constexpr size_t kJSmemBytes =
    sizeof(double) * (2 * kJc * kJLd + kJc * (kJc + 1) + kJc * kJLdJ + 2 * 16) + sizeof(int) * (2 * 16 + 4);

__device__ __forceinline__ void jb_cp_async_8(void* smem_dst, const void* gsrc, bool valid) {
  const int nbytes = valid ? 8 : 0;   // src-size 0 -> the 8 bytes are zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc),
               "r"(nbytes)
               : "memory");
}
__device__ __forceinline__ void jb_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void jb_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma_884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(kJThreads, 4)
jacobi_block_round_kernel(double* __restrict__ Wt, double* __restrict__ Vt, int n, int nb_pad, int r, double tol,
                          int inner_sweeps, unsigned int* __restrict__ rotations) {
  extern __shared__ __align__(16) unsigned char jsm[];
  double* tiles = reinterpret_cast<double*>(jsm);                      // [2][kJc][kJLd]
  double (*G)[kJc + 1] = reinterpret_cast<double (*)[kJc + 1]>(tiles + 2 * kJc * kJLd);
  double (*J)[kJLdJ] = reinterpret_cast<double (*)[kJLdJ]>(&G[kJc][0]);
  double* rot_c = &J[kJc][0];
  double* rot_s = rot_c + 16;
  int* rot_p = reinterpret_cast<int*>(rot_s + 16);
  int* rot_q = rot_p + 16;
  int* any_rot = rot_q + 16;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;   // DMMA fragment coordinates: A[gid][tig], B[tig][gid], C[gid][2 tig + {0, 1}]
  int P, Q;
  tournament_pair(nb_pad, r, blockIdx.x, P, Q);
  if (P > Q) { const int t = P; P = Q; Q = t; }
  const size_t base = (size_t)blockIdx.y * n * n;
  auto gcol = [&](int l) { return l < kJb ? P * kJb + l : Q * kJb + (l - kJb); };   // global column of local column l
  const int n_chunks = (n + kJChunk - 1) / kJChunk;
  // asynchronous copy of chunk ch of the 32 rows into tile buffer ch & 1 (zero beyond the matrix)
  auto stage = [&](const double* __restrict__ M, int ch) {
    double* tile = tiles + (ch & 1) * kJc * kJLd;
    const int i0 = ch * kJChunk;
#pragma unroll
    for (int e = tid; e < kJc * kJChunk; e += kJThreads) {
      const int row = e / kJChunk, col = e - row * kJChunk;
      const int gc = gcol(row), i = i0 + col;
      const bool valid = gc < n && i < n;
      jb_cp_async_8(&tile[row * kJLd + col], M + base + (valid ? (size_t)gc * n + i : 0), valid);
    }
    jb_cp_async_commit();
  };
  // chunk ch is in its buffer and visible to the CTA; the copy of chunk ch + 1 is in flight
  auto advance = [&](const double* __restrict__ M, int ch) {
    if (ch + 1 < n_chunks) {
      stage(M, ch + 1);
      jb_cp_async_wait<1>();
    } else {
      jb_cp_async_wait<0>();
    }
    __syncthreads();
  };

  // 1. Gram matrix G = T T^T on the FP64 tensor pipe: 4 x 4 tiles of 8 x 8, warp w owns tiles (w / 2, 2 (w % 2) + {0, 1})
  {
    const int tr = warp >> 1, tc = (warp & 1) * 2;
    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    stage(Wt, 0);
    for (int ch = 0; ch < n_chunks; ch++) {
      advance(Wt, ch);
      const double* tile = tiles + (ch & 1) * kJc * kJLd;
      const double* pa = tile + (8 * tr + gid) * kJLd + tig;          // A[m = gid][k = tig] = T[8 tr + gid][i + tig]
      const double* pb0 = tile + (8 * tc + gid) * kJLd + tig;         // B[k = tig][n = gid] = T[8 tc + gid][i + tig]
      const double* pb1 = pb0 + 8 * kJLd;
#pragma unroll 8
      for (int i = 0; i < kJChunk; i += 4) {
        const double fa = pa[i], fb0 = pb0[i], fb1 = pb1[i];
        dmma_884(acc[0][0], acc[0][1], fa, fb0);
        dmma_884(acc[1][0], acc[1][1], fa, fb1);
      }
      __syncthreads();   // the buffer is free before the copy two chunks ahead lands in it
    }
#pragma unroll
    for (int t = 0; t < 2; t++) {
      G[8 * tr + gid][8 * (tc + t) + 2 * tig] = acc[t][0];
      G[8 * tr + gid][8 * (tc + t) + 2 * tig + 1] = acc[t][1];
    }
    for (int e = tid; e < kJc * kJc; e += kJThreads) {
      const int row = e / kJc, col = e - row * kJc;
      J[row][col] = row == col ? 1.0 : 0.0;
    }
    if (tid == 0) any_rot[0] = any_rot[1] = 0;
    __syncthreads();
  }

  // 2. cyclic sweep(s) over the 32 columns on G, rotations accumulated in J
  unsigned int my_rot = 0;
  for (int rr = 0; rr < inner_sweeps * (kJc - 1); rr++) {
    if (rr > 0 && rr % (kJc - 1) == 0) {   // a further sweep only if the last one still rotated something
      if (any_rot[1] == 0) break;
      __syncthreads();
      if (tid == 0) any_rot[1] = 0;
      __syncthreads();
    }
    if (tid < 16) {
      int p, q;
      tournament_pair(kJc, rr % (kJc - 1), tid, p, q);
      if (p > q) { const int t = p; p = q; q = t; }
      const double app = G[p][p], aqq = G[q][q], apq = G[p][q];
      double c = 1.0, s = 0.0;
      if (fabs(apq) > tol * sqrt(app * aqq)) {
        const double zeta = (aqq - app) / (2.0 * apq);
        const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        c = 1.0 / sqrt(1.0 + t * t);
        s = c * t;
        my_rot++;
        any_rot[0] = 1;
        any_rot[1] = 1;
      }
      rot_c[tid] = c; rot_s[tid] = s; rot_p[tid] = p; rot_q[tid] = q;
    }
    __syncthreads();
    const int k = tid >> 4, i = tid & 15;
    const double c = rot_c[k], s = rot_s[k];
    const int p = rot_p[k], q = rot_q[k];
    if (s != 0.0) {   // columns p, q of G and of J:  X <- X R
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int ii = i + 16 * h;
        const double gp = G[ii][p], gq = G[ii][q];
        G[ii][p] = c * gp - s * gq;
        G[ii][q] = s * gp + c * gq;
        const double jp = J[ii][p], jq = J[ii][q];
        J[ii][p] = c * jp - s * jq;
        J[ii][q] = s * jp + c * jq;
      }
    }
    __syncthreads();
    if (s != 0.0) {   // rows p, q of G:  G <- R^T G
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int jj = i + 16 * h;
        const double gp = G[p][jj], gq = G[q][jj];
        G[p][jj] = c * gp - s * gq;
        G[q][jj] = s * gp + c * gq;
      }
    }
    __syncthreads();
  }
  if (my_rot) atomicAdd(rotations, my_rot);
  if (any_rot[0] == 0) return;   // these 32 columns are already orthogonal: nothing to write

  // 3. rows <- J^T rows for Wt and Vt on the tensor pipe: out[c][i] = sum_k J[k][c] T[k][i], M = c (4 tiles), N = i (8 tiles
  //    per chunk, one per warp), K = k (8 steps of 4)
  const double* pj = &J[tig][gid];   // A[m = gid][k = tig] = J[4 ks + tig][8 ct + gid]
  for (int which = 0; which < 2; which++) {
    double* M = which == 0 ? Wt : Vt;
    stage(M, 0);
    for (int ch = 0; ch < n_chunks; ch++) {
      advance(M, ch);
      const double* tile = tiles + (ch & 1) * kJc * kJLd;
      double out[4][2];
#pragma unroll
      for (int ct = 0; ct < 4; ct++) out[ct][0] = out[ct][1] = 0.0;
      const double* pb = tile + tig * kJLd + 8 * warp + gid;   // B[k = tig][n = gid] = T[4 ks + tig][8 warp + gid]
#pragma unroll
      for (int ks = 0; ks < 8; ks++) {
        const double fb = pb[4 * ks * kJLd];
#pragma unroll
        for (int ct = 0; ct < 4; ct++) dmma_884(out[ct][0], out[ct][1], pj[4 * ks * kJLdJ + 8 * ct], fb);
      }
      const int i = ch * kJChunk + 8 * warp + 2 * tig;
#pragma unroll
      for (int ct = 0; ct < 4; ct++) {
        const int gc = gcol(8 * ct + gid);
        if (gc >= n) continue;
        double* row = M + base + (size_t)gc * n + i;
        if (i < n) row[0] = out[ct][0];
        if (i + 1 < n) row[1] = out[ct][1];
      }
      __syncthreads();   // the buffer is free before the copy two chunks ahead lands in it
    }
  }
}

__global__ void eigh_init_kernel(const double* __restrict__ A, double* __restrict__ Wt, double* __restrict__ Vt, int n,
                                 int64_t total) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int64_t e = i % ((int64_t)n * n);
    const int row = (int)(e / n), col = (int)(e % n);
    Wt[i] = A[i - e + (int64_t)col * n + row];   // transposed (A is symmetric up to rounding; take it literally)
    Vt[i] = row == col ? 1.0 : 0.0;
  }
}

// lambda_p = sign(v_p . w_p) |w_p|; one CTA per (column, matrix)
__global__ void __launch_bounds__(kEighThreads)
eigh_values_kernel(const double* __restrict__ Wt, const double* __restrict__ Vt, int n, double* __restrict__ evals) {
  const size_t row = ((size_t)blockIdx.y * n + blockIdx.x) * n;
  double ww = 0.0, vw = 0.0, dummy = 0.0;
  for (int i = threadIdx.x; i < n; i += kEighThreads) {
    const double w = Wt[row + i];
    ww += w * w;
    vw += w * Vt[row + i];
  }
  block_sum3(ww, vw, dummy);
  if (threadIdx.x == 0) evals[(size_t)blockIdx.y * n + blockIdx.x] = vw < 0.0 ? -sqrt(ww) : sqrt(ww);
}

}  // namespace

// d_A: n_mat symmetric n x n matrices; d_Wt, d_Vt: scratch / eigenvectors as rows (n_mat * n * n each); d_evals: n_mat * n
// (unsorted); d_rot: one counter.  Returns the number of sweeps, or -1 when max_sweeps was not enough.
int eigh_jacobi(const double* d_A, int n_mat, int n, double* d_Wt, double* d_Vt, double* d_evals, unsigned int* d_rot,
                int max_sweeps, cudaStream_t st, int64_t* launches) {
  const int64_t total = (int64_t)n_mat * n * n;
  eigh_init_kernel<<<(unsigned)std::min<int64_t>((total + 255) / 256, 148 * 16), 256, 0, st>>>(d_A, d_Wt, d_Vt, n, total);
  if (launches) (*launches)++;
  const double tol = 1e-15;
  int sweeps = -1;
  static const bool unblocked = getenv("CGC_EIGH_UNBLOCKED") != nullptr;   // the first version, kept for comparison
  if (unblocked) {
    const int m = (n + 1) & ~1;
    for (int sweep = 0; sweep < max_sweeps; sweep++) {
      cudaMemsetAsync(d_rot, 0, sizeof(unsigned int), st);
      if (m >= 2) {
        for (int r = 0; r < m - 1; r++) {
          jacobi_round_kernel<<<dim3(m / 2, n_mat), kEighThreads, 0, st>>>(d_Wt, d_Vt, n, m, r, tol, d_rot);
        }
        if (launches) *launches += m - 1;
      }
      unsigned int rot = 0;
      cudaMemcpyAsync(&rot, d_rot, sizeof(unsigned int), cudaMemcpyDeviceToHost, st);
      if (cudaStreamSynchronize(st) != cudaSuccess) return -2;
      if (rot == 0) { sweeps = sweep + 1; break; }
    }
  } else {
    cudaFuncSetAttribute(jacobi_block_round_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kJSmemBytes);
    cudaFuncSetAttribute(jacobi_block_round_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);   // four CTAs of 52 KB per SM
    const int nb = (n + kJb - 1) / kJb;
    const int nb_pad = std::max(2, (nb + 1) & ~1);
    static const int inner = getenv("CGC_EIGH_INNER_SWEEPS") ? std::max(1, atoi(getenv("CGC_EIGH_INNER_SWEEPS"))) : 1;   // tuning hook (more inner sweeps per visit were measured: no fewer outer sweeps, slower)
    for (int sweep = 0; sweep < max_sweeps; sweep++) {
      cudaMemsetAsync(d_rot, 0, sizeof(unsigned int), st);
      for (int r = 0; r < nb_pad - 1; r++) {
        jacobi_block_round_kernel<<<dim3(nb_pad / 2, n_mat), kJThreads, kJSmemBytes, st>>>(d_Wt, d_Vt, n, nb_pad, r, tol, inner, d_rot);
      }
      if (launches) *launches += nb_pad - 1;
      unsigned int rot = 0;
      cudaMemcpyAsync(&rot, d_rot, sizeof(unsigned int), cudaMemcpyDeviceToHost, st);
      if (cudaStreamSynchronize(st) != cudaSuccess) return -2;
      if (rot == 0) { sweeps = sweep + 1; break; }
    }
  }
  eigh_values_kernel<<<dim3(n, n_mat), kEighThreads, 0, st>>>(d_Wt, d_Vt, n, d_evals);
  if (launches) (*launches)++;
  return sweeps;
}

}  // namespace cgc

// K6 — symmetric eigen-decomposition of the per-site covariance matrices on the device.
//
// depca diagonalises every site's Ncoarse x Ncoarse covariance with numpy.linalg.eigh
// (reference scripts/depca/depca/sidechain_corr.py:91-96) before sorting the modes.  Here: one-sided (Hestenes) Jacobi,
// batched over the matrices.  W starts as A and V as I, both stored transposed (row p = column p); a rotation of the pair
// (p, q) makes the columns w_p, w_q orthogonal and is applied to v_p, v_q as well; when all columns of W = A V are
// mutually orthogonal, V holds the eigenvectors and |lambda_p| = |w_p| with the sign of v_p . w_p.  The n/2 disjoint
// pairs of a round-robin round are independent: one CTA per (pair, matrix), one launch per round, n - 1 rounds per sweep,
// sweeps until a whole sweep rotates nothing.  The work is memory traffic (every round reads and writes W and V once:
// 32 n^2 bytes per matrix), not arithmetic: 24 matrices of 1099^2 take 14 sweeps = 14 TB of traffic = 2.65 s at 5.4 TB/s,
// against 1.9 s for LAPACK's tridiagonalisation on the box's 16 host cores (scripts/dev_eigh_check.py).  A version that
// kept pairs of 2-6 column blocks in shared memory (1/b of the traffic) was slower still (2.8-3.9 s: the pair loop inside a
// CTA is a chain of block-wide reductions).  So this kernel is the device-only option, not the default of the pyPCA shim.
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
  const int m = (n + 1) & ~1;
  const double tol = 1e-15;
  int sweeps = -1;
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
  eigh_values_kernel<<<dim3(n, n_mat), kEighThreads, 0, st>>>(d_Wt, d_Vt, n, d_evals);
  if (launches) (*launches)++;
  return sweeps;
}

}  // namespace cgc

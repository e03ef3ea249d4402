// Developer microbenchmark: what the FP64 pipe of sm_100a sustains — DFMA with different operand patterns, and DMMA m8n8k4.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 scripts/pipe_bench6.cu -o scripts/pipe_bench6
//
// Patterns (8 independent accumulators per thread in all of them, 64 DFMAs per loop trip):
//   shared_a   acc[q] = fma(a_i, b[q], acc[q])      one multiplicand shared by 8 consecutive DFMAs (K7's source order)
//   distinct   acc[q] = fma(a[q], b[q], acc[q])     no operand shared between consecutive DFMAs
//   window     acc[q] = fma(f[i], w[i + q], acc[q]) K7's register-window pattern, operands refreshed from shared memory
//   dmma       mma.sync.m8n8k4.f64 with 8 accumulator tiles per warp (K3's instruction)
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1;} } while (0)

template <int MODE>
__global__ void __launch_bounds__(128) fp64_kernel(double* out, const double* in, int iters) {
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = in[i];
  __syncthreads();
  double acc[8], a[8], b[16];
#pragma unroll
  for (int q = 0; q < 8; q++) { acc[q] = 0.0; a[q] = in[q + threadIdx.x]; }
#pragma unroll
  for (int q = 0; q < 16; q++) b[q] = in[32 + q + threadIdx.x];
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int q = 0; q < 8; q++) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc[q]) : "d"(a[i]), "d"(b[q]));
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int q = 0; q < 8; q++) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc[q]) : "d"(a[q]), "d"(b[(q + i) & 15]));
    } else if (MODE == 2) {
      // K7's step: 8 f values (broadcast), 8 new window values per step from shared memory
      const int base = (it * 8) & 511;
#pragma unroll
      for (int k = 0; k < 8; k++) { a[k] = sm[base + k]; b[8 + k] = sm[512 + ((base + threadIdx.x) & 255) + k * 32]; }
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int q = 0; q < 8; q++) acc[q] = fma(a[i], b[i + q], acc[q]);
#pragma unroll
      for (int k = 0; k < 8; k++) b[k] = b[8 + k];
    } else {
      // 8 DMMA per trip x 8 trips = the FMA count of 64 DFMAs per thread x 4
#pragma unroll
      for (int r = 0; r < 8; r++)
#pragma unroll
        for (int t = 0; t < 4; t++)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                       : "+d"(acc[2 * t]), "+d"(acc[2 * t + 1]) : "d"(a[r]), "d"(b[t]));
    }
  }
  double s = 0;
#pragma unroll
  for (int q = 0; q < 8; q++) s += acc[q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
int bench(const char* name, int ctas_per_sm, double fma_per_thread_trip) {
  double *out, *in;
  const int threads = 128, grid = 148 * ctas_per_sm;
  CK(cudaMalloc(&out, sizeof(double) * grid * threads));
  CK(cudaMalloc(&in, sizeof(double) * 2048));
  CK(cudaMemset(in, 0, sizeof(double) * 2048));
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  fp64_kernel<MODE><<<grid, threads>>>(out, in, 10);
  cudaEventRecord(e0);
  fp64_kernel<MODE><<<grid, threads>>>(out, in, iters);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double fma = (double)grid * threads * iters * fma_per_thread_trip;
  printf("%-10s %2d warps/SM: %8.3f ms  %6.2f TFLOP/s  %6.2f FMA lanes/clk/SM at 1965 MHz\n", name, ctas_per_sm * 4, ms,
         2 * fma / ms / 1e9, fma / (ms * 1e-3 * 1.965e9) / 148);
  cudaFree(out); cudaFree(in);
  return 0;
}

int main() {
  for (int c : {1, 2, 4, 8}) {
    if (bench<0>("shared_a", c, 64)) return 1;
    if (bench<1>("distinct", c, 64)) return 1;
    if (bench<2>("window", c, 64)) return 1;
    if (bench<3>("dmma", c, 32 * 256.0 / 32)) return 1;   // 32 DMMAs x 256 FMA per warp, per thread: / 32
  }
  return 0;
}

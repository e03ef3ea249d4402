// Developer microbenchmark: sustained issue rates of MUFU.RSQ, packed FP32 (FFMA2) and their mix on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1;} } while (0)
typedef unsigned long long u64;

template <int NM, int NF2, int NF1>
__global__ void mix(float* out, int iters, long long* cyc) {
  float m[NM > 0 ? NM : 1];
  u64 f2[NF2 > 0 ? NF2 : 1];
  float f1[NF1 > 0 ? NF1 : 1];
  for (int i = 0; i < NM; i++) m[i] = 1.0f + threadIdx.x * 1e-3f + i;
  for (int i = 0; i < NF2; i++) f2[i] = 0x3f8000003f800000ull + i;
  for (int i = 0; i < NF1; i++) f1[i] = 1.0f + i;
  u64 c2 = 0x3f8000013f800001ull; float c1 = 1.0000001f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int i = 0; i < (NM > NF2 ? NM : NF2) || i < NF1; i++) {
        if (i < NM) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(m[i]));
        if (i < NF2) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(f2[i]) : "l"(c2));
        if (i < NF1) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f1[i]) : "f"(c1));
      }
      if (NF2 > NM) {
#pragma unroll
        for (int k = 1; k < (NM > 0 ? NF2 / NM : 1) && NM > 0; k++) { }
      }
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < NM; i++) s += m[i];
  for (int i = 0; i < NF2; i++) s += (float)(f2[i] & 0xff);
  for (int i = 0; i < NF1; i++) s += f1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int NM, int NF2, int NF1>
int bench(const char* name, int warps_per_sm) {
  float* out; long long* cyc; long long h;
  int threads = 256, blocks_per_sm = warps_per_sm * 32 / threads, grid = 148 * blocks_per_sm;
  CK(cudaMalloc(&out, sizeof(float) * grid * threads)); CK(cudaMalloc(&cyc, 8));
  int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  mix<NM, NF2, NF1><<<grid, threads>>>(out, 10, cyc);
  cudaEventRecord(e0);
  mix<NM, NF2, NF1><<<grid, threads>>>(out, iters, cyc);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long hc;
  CK(cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost));
  h = (long long)(ms * 1e-3 * 1.965e9);   // cycles at the nominal 1965 MHz from the event time
  printf("[clock64 %lld vs event-derived %lld] ", hc, h);
  double per_sm_threads = warps_per_sm * 32.0;
  double n = (double)iters * 4;
  printf("%-34s warps/SM %2d: cycles %9lld | MUFU %6.2f /clk/SM | FFMA2 %6.2f lane-inst/clk/SM (= %6.2f FMA lanes) | FFMA %6.2f /clk/SM | issue %5.2f warp-inst/clk/SM\n",
         name, warps_per_sm, h, NM * n * per_sm_threads / h, NF2 * n * per_sm_threads / h, 2 * NF2 * n * per_sm_threads / h,
         NF1 * n * per_sm_threads / h, (NM + NF2 + NF1) * n * warps_per_sm / h);
  cudaFree(out); cudaFree(cyc);
  return 0;
}

int main() {
  for (int w : {16, 32}) {
    bench<8, 0, 0>("8 MUFU", w);
    bench<0, 28, 0>("28 FFMA2", w);
    bench<0, 0, 28>("28 FFMA", w);
    bench<8, 28, 0>("8 MUFU + 28 FFMA2 (kernel mix)", w);
    bench<8, 24, 0>("8 MUFU + 24 FFMA2", w);
    bench<8, 20, 0>("8 MUFU + 20 FFMA2", w);
    bench<8, 16, 0>("8 MUFU + 16 FFMA2", w);
    bench<8, 0, 28>("8 MUFU + 28 FFMA", w);
    bench<8, 0, 56>("8 MUFU + 56 FFMA", w);
    bench<8, 14, 28>("8 MUFU + 14 FFMA2 + 28 FFMA", w);
  }
  return 0;
}

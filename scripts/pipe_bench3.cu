// Developer microbenchmark 3: does FFMA2 with three distinct 64-bit sources (or scalar-broadcast operands) keep rt = 2?
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1;} } while (0)
typedef unsigned long long u64;
template <int MODE, int NM>
__global__ void k(float* out, int iters) {
  u64 f[16], g[16], h[16]; float m[8]; float sc = 1.0000001f + threadIdx.x * 1e-9f, sc2 = 0.9999f;
  for (int i = 0; i < 16; i++) { f[i] = 0x3f8000003f800000ull + i; g[i] = 0x3f8000013f800001ull + i; h[i] = 0x3f7fff003f7fff00ull + i; }
  for (int i = 0; i < 8; i++) m[i] = 1.0f + i + threadIdx.x * 1e-3f;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 2; r++) {
#pragma unroll
      for (int i = 0; i < 16; i++) {
        if (MODE == 0) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(f[i]) : "l"(g[i]));              // 2 distinct
        if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(f[i]) : "l"(g[i]), "l"(h[i]));   // 3 distinct
        if (MODE == 2) { u64 b; asm volatile("mov.b64 %0, {%1, %1};" : "=l"(b) : "f"(sc)); asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(f[i]) : "l"(b), "l"(g[i])); }  // scalar + 2 pairs
        if (MODE == 3) { u64 b; asm volatile("mov.b64 %0, {%1, %1};" : "=l"(b) : "f"(sc)); asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(f[i]) : "l"(b)); }   // FADD2 pair + scalar
        if (i < NM && (i & 1) == 0) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(m[i >> 1]));
      }
    }
  }
  float s = 0; for (int i = 0; i < 16; i++) s += (float)((f[i] ^ g[i] ^ h[i]) & 0xff); for (int i = 0; i < 8; i++) s += m[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + sc2;
}
template <int MODE, int NM>
int bench(const char* name) {
  float* out; int threads = 256, grid = 148 * 2; CK(cudaMalloc(&out, 4 * grid * threads));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters = 20000;
  k<MODE, NM><<<grid, threads>>>(out, 10);
  cudaEventRecord(e0); k<MODE, NM><<<grid, threads>>>(out, iters); cudaEventRecord(e1);
  CK(cudaDeviceSynchronize()); float ms; cudaEventElapsedTime(&ms, e0, e1);
  double cyc = ms * 1e-3 * 1.965e9, n = iters * 2.0, thr = 16 * 32.0;
  printf("%-40s packed %6.2f lane-inst/clk/SM (peak 64)   MUFU %6.2f/clk/SM\n", name, 16 * n * thr / cyc, (NM / 2) * n * thr / cyc);
  cudaFree(out); return 0;
}
int main() {
  bench<0, 0>("FFMA2 2 distinct pairs");
  bench<1, 0>("FFMA2 3 distinct pairs");
  bench<2, 0>("FFMA2 scalar + 2 pairs");
  bench<3, 0>("FADD2 pair + scalar");
  bench<1, 8>("FFMA2 3 distinct + MUFU (16:4)");
  bench<2, 8>("FFMA2 scalar+2 pairs + MUFU (16:4)");
  return 0;
}

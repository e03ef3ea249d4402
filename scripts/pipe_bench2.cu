// Developer microbenchmark 2: FADD2 / FMUL2 / FFMA2 rates and the exact kernel mix, with and without LDS.128 broadcasts.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1;} } while (0)
typedef unsigned long long u64;

// per "round": NA add2, NMU mul2, NF fma2, NM mufu, NL lds.128 (broadcast)
template <int NA, int NMU, int NF, int NM, int NL>
__global__ void mix(float* out, int iters) {
  __shared__ ulonglong2 sm[64];
  if (threadIdx.x < 64) sm[threadIdx.x] = make_ulonglong2(0x3f8000003f800000ull, 0x3f8000003f800000ull);
  __syncthreads();
  u64 a[NA > 0 ? NA : 1], mu[NMU > 0 ? NMU : 1], f[NF > 0 ? NF : 1];
  float m[NM > 0 ? NM : 1];
  for (int i = 0; i < NA; i++) a[i] = 0x3f8000003f800000ull + i;
  for (int i = 0; i < NMU; i++) mu[i] = 0x3f8000003f800000ull + i;
  for (int i = 0; i < NF; i++) f[i] = 0x3f8000003f800000ull + i;
  for (int i = 0; i < NM; i++) m[i] = 1.0f + threadIdx.x * 1e-3f + i;
  u64 c2 = 0x3f8000013f800001ull;
  u64 ldacc = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int i = 0; i < NL; i++) {
        ulonglong2 v;
        asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"((unsigned)__cvta_generic_to_shared(&sm[(it + i + r) & 63])));
        ldacc ^= v.x ^ v.y;
      }
#pragma unroll
      for (int i = 0; i < 16; i++) {
        if (i < NA) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(c2));
        if (i < NF) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(f[i]) : "l"(c2));
        if (i < NM) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(m[i]));
        if (i < NMU) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(mu[i]) : "l"(c2));
      }
    }
  }
  float s = (float)(ldacc & 0xff);
  for (int i = 0; i < NA; i++) s += (float)(a[i] & 0xff);
  for (int i = 0; i < NMU; i++) s += (float)(mu[i] & 0xff);
  for (int i = 0; i < NF; i++) s += (float)(f[i] & 0xff);
  for (int i = 0; i < NM; i++) s += m[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NA, int NMU, int NF, int NM, int NL>
int bench(const char* name, int warps_per_sm) {
  float* out;
  int threads = 256, grid = 148 * warps_per_sm * 32 / threads;
  CK(cudaMalloc(&out, sizeof(float) * grid * threads));
  int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  mix<NA, NMU, NF, NM, NL><<<grid, threads>>>(out, 10);
  cudaEventRecord(e0);
  mix<NA, NMU, NF, NM, NL><<<grid, threads>>>(out, iters);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double cyc = ms * 1e-3 * 1.965e9, n = (double)iters * 4, thr = warps_per_sm * 32.0;
  printf("%-44s w/SM %2d | MUFU %6.2f/clk/SM | packed FP32 %6.2f lane-inst/clk/SM (peak 64) | LDS.128 %5.2f warp-inst/clk/SM | issue %5.2f\n", name, warps_per_sm,
         NM * n * thr / cyc, (NA + NMU + NF) * n * thr / cyc, NL * n * warps_per_sm / cyc, (NA + NMU + NF + NM + NL) * n * warps_per_sm / cyc);
  cudaFree(out);
  return 0;
}

int main() {
  for (int w : {16}) {
    bench<16, 0, 0, 0, 0>("16 FADD2", w);
    bench<0, 16, 0, 0, 0>("16 FMUL2", w);
    bench<0, 0, 16, 0, 0>("16 FFMA2", w);
    bench<12, 4, 12, 0, 0>("12 FADD2 + 4 FMUL2 + 12 FFMA2", w);
    bench<12, 4, 12, 8, 0>("kernel mix: 12 A2 + 4 M2 + 12 F2 + 8 MUFU", w);
    bench<12, 4, 12, 8, 2>("kernel mix + 2 LDS.128", w);
    bench<0, 0, 0, 8, 2>("8 MUFU + 2 LDS.128", w);
    bench<0, 0, 0, 8, 8>("8 MUFU + 8 LDS.128", w);
    bench<0, 0, 0, 0, 8>("8 LDS.128", w);
  }
  return 0;
}

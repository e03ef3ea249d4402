// Developer microbenchmark 5: the bare pair loop with NP packed pairs (2*NP atoms) per thread and different unroll
// factors: how much does one broadcast LDS.128 per site atom cost when it is shared by 8, 12 or 16 pairs?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 scripts/pipe_bench5.cu -o scripts/pipe_bench5
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1;} } while (0)
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float rsq(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ u64 r2of(u64 x, u64 y, u64 z, float4 sa) {
  u64 dx = fadd2(x, pack2(sa.x, sa.x)), dy = fadd2(y, pack2(sa.y, sa.y)), dz = fadd2(z, pack2(sa.z, sa.z));
  return ffma2(dz, dz, ffma2(dy, dy, fmul2(dx, dx)));
}
__device__ __forceinline__ u64 rsq2(u64 r) { float a, b; unpack2(r, a, b); return pack2(rsq(a), rsq(b)); }

constexpr int P = 44, S = 24;

// kMode 0: site data by LDS.128; 1: no loads at all (site entry derived from registers: the loop's own ceiling)
template <int NP, int UN, int kThreads, int kMinB, int kMode>
__global__ void __launch_bounds__(kThreads, kMinB) k(const float4* __restrict__ g_sites, const u64* __restrict__ xyz, float* out, int reps) {
  __shared__ float4 sm[S * P];
  for (int i = threadIdx.x; i < S * P; i += kThreads) sm[i] = g_sites[i];
  __syncthreads();
  u64 X[NP], Y[NP], Z[NP];
#pragma unroll
  for (int j = 0; j < NP; j++) { X[j] = xyz[(threadIdx.x % 256) * 12 + j % 4] + 3 * j; Y[j] = xyz[(threadIdx.x % 256) * 12 + 4 + j % 4] + 5 * j; Z[j] = xyz[(threadIdx.x % 256) * 12 + 8 + j % 4] + 7 * j; }   // distinct values: no common subexpressions between pairs
  float accum = 0.f;
  float4 fake = g_sites[threadIdx.x & 7];
  for (int rep = 0; rep < reps; rep++) {
    for (int s = 0; s < S; s++) {
      const float4* sp = sm + s * P;
      u64 V[NP], R[NP], RI[NP];
      float4 s0 = sp[0], s1 = sp[1];
#pragma unroll
      for (int j = 0; j < NP; j++) { V[j] = 0; RI[j] = rsq2(r2of(X[j], Y[j], Z[j], s0)); }
#pragma unroll
      for (int j = 0; j < NP; j++) R[j] = r2of(X[j], Y[j], Z[j], s1);
      float dq2 = s0.w, dq1 = s1.w;
#pragma unroll UN
      for (int c = 2; c < P; c++) {
        float4 sa;
        if (kMode == 0) sa = sp[c];
        else { fake.x += 1.0f; sa = fake; }
#pragma unroll
        for (int j = 0; j < NP; j++) {
          V[j] = ffma2(pack2(dq2, dq2), RI[j], V[j]);
          RI[j] = rsq2(R[j]);
          R[j] = r2of(X[j], Y[j], Z[j], sa);
        }
        dq2 = dq1; dq1 = sa.w;
      }
#pragma unroll
      for (int j = 0; j < NP; j++) { V[j] = ffma2(pack2(dq2, dq2), RI[j], V[j]); V[j] = ffma2(pack2(dq1, dq1), rsq2(R[j]), V[j]); }
#pragma unroll
      for (int j = 0; j < NP; j++) { float a, b; unpack2(V[j], a, b); accum += a + b; }
    }
  }
  out[blockIdx.x * kThreads + threadIdx.x] = accum;
}

template <int NP, int UN, int kThreads, int kMinB, int kMode>
int bench(const char* name, const float4* d_sites, const u64* d_xyz, float* d_out) {
  int reps = 20, grid = 148 * kMinB;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto kern = k<NP, UN, kThreads, kMinB, kMode>;
  kern<<<grid, kThreads>>>(d_sites, d_xyz, d_out, 1);
  cudaEventRecord(e0); kern<<<grid, kThreads>>>(d_sites, d_xyz, d_out, reps); cudaEventRecord(e1);
  CK(cudaDeviceSynchronize()); float ms; cudaEventElapsedTime(&ms, e0, e1);
  double pairs = (double)grid * kThreads * 2 * NP * S * P * reps;
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
  printf("%-40s pairs/thr %2d unroll %2d thr %3d x %d CTA/SM regs %3d  %.3f ms  %.1f Gpairs/s  frac %.3f\n", name, 2 * NP, UN, kThreads, kMinB, fa.numRegs, ms,
         pairs / ms / 1e6, pairs / ms / 1e6 / 4653.1);
  return 0;
}

int main() {
  float4 h[S * P]; for (int i = 0; i < S * P; i++) h[i] = make_float4(-1.f - 0.01f * i, -2.f - 0.02f * i, -3.f + 0.01f * i, 0.01f * ((i % 7) - 3));
  float4* d_sites; CK(cudaMalloc(&d_sites, sizeof(h))); CK(cudaMemcpy(d_sites, h, sizeof(h), cudaMemcpyHostToDevice));
  u64* d_xyz; CK(cudaMalloc(&d_xyz, 256 * 12 * 8));
  float hx[256 * 24]; for (int i = 0; i < 256 * 24; i++) hx[i] = 0.37f * (i % 97);
  CK(cudaMemcpy(d_xyz, hx, sizeof(hx), cudaMemcpyHostToDevice));
  float* d_out; CK(cudaMalloc(&d_out, 148 * 4 * 512 * 4));
  for (int rep = 0; rep < 2; rep++) {
    bench<4, 21, 256, 2, 0>("LDS", d_sites, d_xyz, d_out);
    bench<4, 14, 256, 2, 0>("LDS", d_sites, d_xyz, d_out);
    bench<4, 42, 256, 2, 1>("no loads (loop ceiling)", d_sites, d_xyz, d_out);
    bench<4, 21, 256, 2, 1>("no loads (loop ceiling)", d_sites, d_xyz, d_out);
    bench<6, 21, 256, 2, 0>("LDS", d_sites, d_xyz, d_out);
    bench<6, 14, 256, 2, 0>("LDS", d_sites, d_xyz, d_out);
    bench<6, 42, 256, 2, 1>("no loads (loop ceiling)", d_sites, d_xyz, d_out);
    bench<8, 21, 256, 2, 0>("LDS", d_sites, d_xyz, d_out);
    bench<8, 14, 256, 2, 0>("LDS", d_sites, d_xyz, d_out);
    bench<8, 42, 256, 2, 1>("no loads (loop ceiling)", d_sites, d_xyz, d_out);
  }
  return 0;
}

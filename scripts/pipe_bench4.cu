// Developer microbenchmark 4: the real pair loop (dependency structure included) with the site data coming from
// (a) shared memory (vector registers, as in cdc_kernel) or (b) constant memory (uniform registers / constant operands).
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
__constant__ float4 c_sites[S * P];

template <bool kConst>
__global__ void __launch_bounds__(256, 2) k(const float4* __restrict__ g_sites, const u64* __restrict__ xyz, float* out, int reps) {
  __shared__ float4 sm[S * P];
  for (int i = threadIdx.x; i < S * P; i += 256) sm[i] = g_sites[i];
  __syncthreads();
  u64 X[4], Y[4], Z[4];
  for (int j = 0; j < 4; j++) { X[j] = xyz[threadIdx.x * 12 + j]; Y[j] = xyz[threadIdx.x * 12 + 4 + j]; Z[j] = xyz[threadIdx.x * 12 + 8 + j]; }
  float accum = 0.f;
  for (int rep = 0; rep < reps; rep++) {
    for (int s = 0; s < S; s++) {
      const float4* sp = (kConst ? c_sites : sm) + s * P;
      u64 V[4] = {0, 0, 0, 0}, R[4], RI[4];
      float4 s0 = sp[0], s1 = sp[1];
      for (int j = 0; j < 4; j++) RI[j] = rsq2(r2of(X[j], Y[j], Z[j], s0));
      for (int j = 0; j < 4; j++) R[j] = r2of(X[j], Y[j], Z[j], s1);
      float dq2 = s0.w, dq1 = s1.w;
#pragma unroll 21
      for (int c = 2; c < P; c++) {
        const float4 sa = sp[c];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          V[j] = ffma2(pack2(dq2, dq2), RI[j], V[j]);
          RI[j] = rsq2(R[j]);
          R[j] = r2of(X[j], Y[j], Z[j], sa);
        }
        dq2 = dq1; dq1 = sa.w;
      }
      for (int j = 0; j < 4; j++) { V[j] = ffma2(pack2(dq2, dq2), RI[j], V[j]); V[j] = ffma2(pack2(dq1, dq1), rsq2(R[j]), V[j]); }
      for (int j = 0; j < 4; j++) { float a, b; unpack2(V[j], a, b); accum += a + b; }
    }
  }
  out[blockIdx.x * 256 + threadIdx.x] = accum;
}

template <bool kConst>
int bench(const char* name, const float4* d_sites, const u64* d_xyz, float* d_out) {
  int reps = 20, grid = 296;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<kConst><<<grid, 256>>>(d_sites, d_xyz, d_out, 1);
  cudaEventRecord(e0); k<kConst><<<grid, 256>>>(d_sites, d_xyz, d_out, reps); cudaEventRecord(e1);
  CK(cudaDeviceSynchronize()); float ms; cudaEventElapsedTime(&ms, e0, e1);
  double pairs = (double)grid * 256 * 8 * S * P * reps;
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k<kConst>);
  printf("%-34s regs %3d  %.3f ms  %.1f Gpairs/s  frac of 16/clk/SM@1.965GHz %.3f\n", name, fa.numRegs, ms, pairs / ms / 1e6, pairs / ms / 1e6 / 4653.1);
  return 0;
}

int main() {
  float4 h[S * P]; for (int i = 0; i < S * P; i++) h[i] = make_float4(-1.f - 0.01f * i, -2.f - 0.02f * i, -3.f + 0.01f * i, 0.01f * ((i % 7) - 3));
  float4* d_sites; CK(cudaMalloc(&d_sites, sizeof(h))); CK(cudaMemcpy(d_sites, h, sizeof(h), cudaMemcpyHostToDevice));
  CK(cudaMemcpyToSymbol(c_sites, h, sizeof(h)));
  u64* d_xyz; CK(cudaMalloc(&d_xyz, 256 * 12 * 8));
  float hx[256 * 24]; for (int i = 0; i < 256 * 24; i++) hx[i] = 0.37f * (i % 97);
  CK(cudaMemcpy(d_xyz, hx, sizeof(hx), cudaMemcpyHostToDevice));
  float* d_out; CK(cudaMalloc(&d_out, 296 * 256 * 4));
  bench<false>("site data in shared memory", d_sites, d_xyz, d_out);
  bench<true>("site data in constant memory", d_sites, d_xyz, d_out);
  return 0;
}

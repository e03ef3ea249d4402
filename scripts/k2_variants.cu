// Developer microbenchmark: times template variants of the CDC pair kernel on C2-shaped random data.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I cgromcorrfmo_b200/csrc scripts/k2_variants.cu -o scripts/k2_variants
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#include <cmath>
#include <string>
#include "cdc_kernel.cuh"

using namespace cgc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

struct Data {
  int N = 100000, n_pad, S = 24, P = 44, G, F;
  float* xyz8; float4* sites; double* kq; int* col; double* acc;
};

template <int MB, int UN, int FL = 0>
double run(const Data& d, const char* name, int sm_count, double* checksum, int chunk_sites = 8) {
  auto kern = cdc_kernel<MB, UN, FL>;
  size_t smem = cdc_smem_layout(chunk_sites, d.P, nullptr);
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTileThreads, smem));
  cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
  int tiles = d.n_pad / kTileAtoms, cpi = (d.S + chunk_sites - 1) / chunk_sites;
  long long units = (long long)tiles * d.F * cpi;
  int grid = (int)std::min<long long>(units, sm_count * per_sm);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto launch = [&]() { kern<<<grid, kTileThreads, smem>>>(d.xyz8, d.sites, d.kq, d.col, d.acc, d.n_pad, tiles, units, d.S, d.P, d.G, chunk_sites, cpi); };
  CK(cudaMemset(d.acc, 0, sizeof(double) * (size_t)d.F * d.S * d.G));
  for (int i = 0; i < 2; i++) launch();
  CK(cudaDeviceSynchronize());
  CK(cudaMemset(d.acc, 0, sizeof(double) * (size_t)d.F * d.S * d.G));
  const int reps = 5;
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; i++) launch();
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
  std::vector<double> h((size_t)d.F * d.S * d.G);
  CK(cudaMemcpy(h.data(), d.acc, sizeof(double) * h.size(), cudaMemcpyDeviceToHost));
  double cs = 0; for (double v : h) cs += fabs(v);
  *checksum = cs / reps;
  double pairs = (double)d.F * d.S * d.P * (d.N - 140);
  double gp = pairs / ms / 1e6;
  printf("chunk %2d %-25s regs %3d occ %d grid %4d units/CTA %7.2f smem %6zu  %8.3f ms  %8.1f Gpairs/s  frac(16/clk@1.965) %.3f  checksum %.9e\n", chunk_sites, name, fa.numRegs, per_sm, grid, (double)units / grid, smem, ms, gp, gp / 4653.1, *checksum);
  return ms;
}

int main(int argc, char** argv) {
  Data d; d.F = argc > 1 ? atoi(argv[1]) : 64;
  d.n_pad = (d.N + kTileAtoms - 1) / kTileAtoms * kTileAtoms;
  std::mt19937 rng(1);
  std::uniform_real_distribution<float> U(0.f, 10.f);
  std::normal_distribution<float> Nq(0.f, 0.4f);
  // columns: 16.5k protein atoms in runs of 7..24, then 24 x 140 chromophore atoms, then one solvent group
  std::vector<int> col(d.n_pad, -1); std::vector<double> kq(d.n_pad, 0.0);
  int a = 0, g = 24;
  while (a < 16500) { int k = 7 + rng() % 18; for (int i = 0; i < k && a < 16500; i++) col[a++] = g; g++; }
  for (int s = 0; s < 24; s++) for (int i = 0; i < 140; i++) col[a++] = s;
  const bool permol = argc > 2 && std::string(argv[2]) == "permol";   // every 3 solvent atoms their own column (C4-like)
  { int a_s = a; for (; a < d.N; a++) col[a] = permol ? g + (a - a_s) / 3 : g; g = col[d.N - 1]; }
  d.G = g + 1;
  for (int i = 0; i < d.N; i++) kq[i] = 33.2 * Nq(rng);
  std::vector<float> xyz((size_t)d.F * d.n_pad * 3, 0.f);
  std::vector<float4> sites((size_t)d.F * d.S * d.P);
  for (int f = 0; f < d.F; f++) {
    for (int i = 0; i < d.N; i++) for (int k = 0; k < 3; k++) xyz[((size_t)f * d.n_pad + (i & ~7)) * 3 + k * 8 + (i & 7)] = U(rng);
    for (int s = 0; s < d.S; s++) for (int c = 0; c < d.P; c++) {
      // a dQ atom of chromophore s: reuse the coordinates of one of its atoms
      int atom = 16500 + s * 140 + c * 3;
      const float* p = &xyz[((size_t)f * d.n_pad + (atom & ~7)) * 3 + (atom & 7)];
      float dq = 0.03f * Nq(rng);
      sites[(size_t)f * d.S * d.P + s * d.P + c] = make_float4(-p[0], -p[8], -p[16], dq);
    }
  }
  CK(cudaMalloc(&d.xyz8, xyz.size() * 4)); CK(cudaMemcpy(d.xyz8, xyz.data(), xyz.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d.sites, sites.size() * 16)); CK(cudaMemcpy(d.sites, sites.data(), sites.size() * 16, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d.kq, kq.size() * 8)); CK(cudaMemcpy(d.kq, kq.data(), kq.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d.col, col.size() * 4)); CK(cudaMemcpy(d.col, col.data(), col.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d.acc, sizeof(double) * (size_t)d.F * d.S * d.G));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("%s, %d SMs, F=%d frames, G=%d\n", prop.name, prop.multiProcessorCount, d.F, d.G);
  double cs_;
  int sm = prop.multiProcessorCount;
  // reference checksum of the round-1 kernel on the same data (F=128): 2.584917105e+06
  printf("kTileThreads %d, tile %d atoms, n_pad %d\n", kTileThreads, kTileAtoms, d.n_pad);
  for (int rep = 0; rep < 2; rep++) {
    run<2, 42, 0>(d, "unroll42", sm, &cs_, 8);
    run<2, 42, 1>(d, "unroll42 NO-EPI", sm, &cs_, 8);
  }
  return 0;
}

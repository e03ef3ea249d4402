// Developer measurement: what does it cost to get N MB of page-locked host memory, by which route?
//   nvcc -O2 -std=c++17 -o scripts/pinned_alloc_bench scripts/pinned_alloc_bench.cu -lpthread
#include <cuda_runtime.h>
#include <sys/mman.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
  const size_t mb = argc > 1 ? (size_t)atoll(argv[1]) : 128, bytes = mb << 20;
  cudaFree(0);
  char* d = nullptr;
  cudaMalloc(&d, bytes);
  cudaStream_t st;
  cudaStreamCreate(&st);
  auto h2d = [&](const char* src) {
    cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);
    double t0 = now();
    for (int i = 0; i < 4; i++) cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);
    return 4.0 * bytes / (now() - t0) / 1e9;
  };
  for (int rep = 0; rep < 3; rep++) {
    double t0 = now();
    char* a = nullptr;
    cudaHostAlloc(&a, bytes, cudaHostAllocDefault);
    double t1 = now();
    printf("cudaHostAlloc %zu MB: %.1f ms, H2D %.1f GB/s", mb, (t1 - t0) * 1e3, h2d(a));
    t0 = now();
    cudaFreeHost(a);
    printf(", free %.1f ms\n", (now() - t0) * 1e3);
  }
  for (int mode = 0; mode < 3; mode++) {
    for (int rep = 0; rep < 2; rep++) {
      double t0 = now();
      void* p = mmap(nullptr, bytes + (2 << 20), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
      char* a = (char*)(((uintptr_t)p + (2 << 20) - 1) & ~(uintptr_t)((2 << 20) - 1));
      if (mode >= 1) madvise(a, bytes, MADV_HUGEPAGE);
      const int nt = mode == 2 ? 8 : 1;
      std::vector<std::thread> th;
      for (int t = 0; t < nt; t++) th.emplace_back([=] { memset(a + bytes / nt * t, 0, bytes / nt); });
      for (auto& t : th) t.join();
      double t1 = now();
      cudaError_t e = cudaHostRegister(a, bytes, cudaHostRegisterDefault);
      double t2 = now();
      printf("mmap%s + touch x%d %.1f ms + cudaHostRegister %.1f ms (%s)", mode >= 1 ? " + MADV_HUGEPAGE" : "", nt, (t1 - t0) * 1e3,
             (t2 - t1) * 1e3, cudaGetErrorString(e));
      if (e == cudaSuccess) {
        printf(", H2D %.1f GB/s", h2d(a));
        t0 = now();
        cudaHostUnregister(a);
        printf(", unregister %.1f ms", (now() - t0) * 1e3);
      }
      munmap(p, bytes + (2 << 20));
      printf("\n");
    }
  }
  FILE* f = fopen("/sys/kernel/mm/transparent_hugepage/enabled", "r");
  char buf[128] = {};
  if (f) { if (fgets(buf, sizeof buf, f)) printf("THP enabled: %s", buf); fclose(f); }
  return 0;
}

// Developer experiment: how fast can page-cache pages of a trajectory file be made DMA-able?
//   (a) pread into a pinned staging buffer with N threads (what the CLI does today)
//   (b) cudaHostRegister windows of a read-only mmap of the file, then cudaMemcpyAsync straight from the mapping
// nvcc -O2 -std=c++17 scripts/dev_hostregister.cu -o scripts/dev_hostregister -lpthread
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
  const char* path = argv[1];
  const size_t win = (size_t)(argc > 2 ? atoll(argv[2]) : 512) << 20;
  int fd = open(path, O_RDONLY);
  struct stat sb; fstat(fd, &sb);
  size_t n = sb.st_size; n = n / win * win; if (n > ((size_t)8 << 30)) n = (size_t)8 << 30;
  printf("file %.2f GB used, window %zu MB\n", n / 1e9, win >> 20);
  char* d; CK(cudaMalloc(&d, win));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  // (a) threaded pread into pinned staging
  for (int nt : {8, 16}) {
    char* h; CK(cudaHostAlloc(&h, win, cudaHostAllocDefault));
    double t0 = now();
    for (size_t off = 0; off < n; off += win) {
      std::vector<std::thread> th;
      size_t chunk = win / nt;
      for (int i = 0; i < nt; i++) th.emplace_back([=] { size_t b = i * chunk, e = b + chunk; while (b < e) { ssize_t g = pread(fd, h + b, e - b, off + b); if (g <= 0) break; b += g; } });
      for (auto& t : th) t.join();
      CK(cudaMemcpyAsync(d, h, win, cudaMemcpyHostToDevice, st));
      CK(cudaStreamSynchronize(st));
    }
    double dt = now() - t0;
    printf("(a) pread x%d threads -> pinned -> H2D (serialised): %.2f GB/s\n", nt, n / dt / 1e9);
    t0 = now();
    for (size_t off = 0; off < n; off += win) {
      std::vector<std::thread> th;
      size_t chunk = win / nt;
      for (int i = 0; i < nt; i++) th.emplace_back([=] { size_t b = i * chunk, e = b + chunk; while (b < e) { ssize_t g = pread(fd, h + b, e - b, off + b); if (g <= 0) break; b += g; } });
      for (auto& t : th) t.join();
    }
    dt = now() - t0;
    printf("(a) pread x%d threads -> pinned only: %.2f GB/s\n", nt, n / dt / 1e9);
    cudaFreeHost(h);
  }
  // (b) register windows of the mapping
  for (unsigned flags : {(unsigned)cudaHostRegisterReadOnly, (unsigned)cudaHostRegisterDefault}) {
    char* m = (char*)mmap(nullptr, n, PROT_READ, MAP_SHARED, fd, 0);
    if (m == MAP_FAILED) { printf("mmap failed\n"); return 1; }
    double t_reg = 0, t_copy = 0, t_unreg = 0;
    bool ok = true;
    for (size_t off = 0; off < n && ok; off += win) {
      double t0 = now();
      cudaError_t e = cudaHostRegister(m + off, win, flags);
      if (e != cudaSuccess) { printf("(b) cudaHostRegister(flags=%u) failed: %s\n", flags, cudaGetErrorString(e)); cudaGetLastError(); ok = false; break; }
      double t1 = now();
      CK(cudaMemcpyAsync(d, m + off, win, cudaMemcpyHostToDevice, st));
      CK(cudaStreamSynchronize(st));
      double t2 = now();
      CK(cudaHostUnregister(m + off));
      double t3 = now();
      t_reg += t1 - t0; t_copy += t2 - t1; t_unreg += t3 - t2;
    }
    if (ok) printf("(b) flags=%u: register %.2f GB/s, H2D from the mapping %.2f GB/s, unregister %.2f GB/s; serialised total %.2f GB/s\n", flags,
                   n / t_reg / 1e9, n / t_copy / 1e9, n / t_unreg / 1e9, n / (t_reg + t_copy + t_unreg) / 1e9);
    munmap(m, n);
  }
  // (d) per window: every thread populates its slice of a fresh mapping (MADV_POPULATE_READ) and copies it
  for (int nt : {8, 16}) {
    char* m = (char*)mmap(nullptr, n, PROT_READ, MAP_SHARED, fd, 0);
    char* h; CK(cudaHostAlloc(&h, win, cudaHostAllocDefault));
    double t0 = now();
    int fails = 0;
    for (size_t off = 0; off < n; off += win) {
      std::vector<std::thread> th;
      size_t chunk = win / nt;
      for (int i = 0; i < nt; i++) th.emplace_back([=, &fails] {
#ifdef MADV_POPULATE_READ
        if (madvise(m + off + i * chunk, chunk, MADV_POPULATE_READ) != 0) fails++;
#endif
        memcpy(h + i * chunk, m + off + i * chunk, chunk);
      });
      for (auto& t : th) t.join();
    }
    double dt = now() - t0;
    printf("(d) MADV_POPULATE_READ + memcpy x%d threads, fresh mapping -> pinned: %.2f GB/s (madvise failures %d)\n", nt, n / dt / 1e9, fails);
    munmap(m, n);
    cudaFreeHost(h);
  }
  // (e) plain memcpy from a fresh (unpopulated) mapping: page faults with fault-around
  {
    int nt = 16;
    char* m = (char*)mmap(nullptr, n, PROT_READ, MAP_SHARED, fd, 0);
    char* h; CK(cudaHostAlloc(&h, win, cudaHostAllocDefault));
    double t0 = now();
    for (size_t off = 0; off < n; off += win) {
      std::vector<std::thread> th;
      size_t chunk = win / nt;
      for (int i = 0; i < nt; i++) th.emplace_back([=] { memcpy(h + i * chunk, m + off + i * chunk, chunk); });
      for (auto& t : th) t.join();
    }
    double dt = now() - t0;
    printf("(e) memcpy x%d threads from a fresh mapping (page faults) -> pinned: %.2f GB/s\n", nt, n / dt / 1e9);
    munmap(m, n);
    cudaFreeHost(h);
  }
  // (c) memcpy from a populated mapping with threads
  {
    char* m = (char*)mmap(nullptr, n, PROT_READ, MAP_SHARED | MAP_POPULATE, fd, 0);
    char* h; CK(cudaHostAlloc(&h, win, cudaHostAllocDefault));
    int nt = 16;
    double t0 = now();
    for (size_t off = 0; off < n; off += win) {
      std::vector<std::thread> th;
      size_t chunk = win / nt;
      for (int i = 0; i < nt; i++) th.emplace_back([=] { memcpy(h + i * chunk, m + off + i * chunk, chunk); });
      for (auto& t : th) t.join();
    }
    double dt = now() - t0;
    printf("(c) memcpy x%d threads from a MAP_POPULATE mapping -> pinned: %.2f GB/s\n", nt, n / dt / 1e9);
    munmap(m, n);
  }
  return 0;
}

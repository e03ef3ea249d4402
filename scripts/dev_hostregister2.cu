// Developer experiment 2: can the pages of a trajectory file be DMA'd in place?  mmap(MAP_SHARED, PROT_READ|PROT_WRITE) of a
// file opened O_RDWR, cudaHostRegister of 128 MB windows, cudaMemcpyAsync straight from the mapping, unregister.
//   nvcc -O2 -std=c++17 scripts/dev_hostregister2.cu -o scripts/dev_hostregister2 -lpthread ; scripts/dev_hostregister2 /dev/shm/x.bin
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
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
  const char* path = argv[1];
  const size_t win = (size_t)(argc > 2 ? atoll(argv[2]) : 128) << 20;
  const int nthreads = argc > 3 ? atoi(argv[3]) : 4;
  cudaFree(0);
  int ro_supported = 0;
  cudaDeviceGetAttribute(&ro_supported, cudaDevAttrHostRegisterReadOnlySupported, 0);
  printf("cudaDevAttrHostRegisterReadOnlySupported = %d\n", ro_supported);
  for (int mode = 0; mode < 2; mode++) {   // 0: O_RDWR + PROT_READ|PROT_WRITE, default flags; 1: O_RDONLY + PROT_READ, ReadOnly flag
    int fd = open(path, mode == 0 ? O_RDWR : O_RDONLY);
    if (fd < 0) { printf("open failed\n"); continue; }
    struct stat sb; fstat(fd, &sb);
    size_t n = (size_t)sb.st_size / win * win;
    char* m = (char*)mmap(nullptr, n, mode == 0 ? PROT_READ | PROT_WRITE : PROT_READ, MAP_SHARED, fd, 0);
    if (m == MAP_FAILED) { printf("mmap failed\n"); close(fd); continue; }
    char* d; cudaMalloc(&d, win);
    cudaStream_t st; cudaStreamCreate(&st);
    const unsigned flags = mode == 0 ? cudaHostRegisterDefault : cudaHostRegisterReadOnly;
    double t_reg = 0, t_copy = 0, t_unreg = 0;
    bool ok = true;
    size_t done = 0;
    for (size_t off = 0; off < n && ok; off += win) {
      double t0 = now();
      cudaError_t e = cudaHostRegister(m + off, win, flags);
      if (e != cudaSuccess) { printf("mode %d: cudaHostRegister failed: %s\n", mode, cudaGetErrorString(e)); cudaGetLastError(); ok = false; break; }
      double t1 = now();
      cudaMemcpyAsync(d, m + off, win, cudaMemcpyHostToDevice, st);
      cudaStreamSynchronize(st);
      double t2 = now();
      cudaHostUnregister(m + off);
      double t3 = now();
      t_reg += t1 - t0; t_copy += t2 - t1; t_unreg += t3 - t2; done += win;
    }
    if (ok && done) printf("mode %d (%s): %zu windows of %zu MB: register %.1f ms, H2D %.1f ms (%.1f GB/s), unregister %.1f ms per window; serial rate %.1f GB/s\n", mode,
                           mode == 0 ? "RDWR mapping" : "RDONLY mapping + ReadOnly flag", done / win, win >> 20, t_reg / (done / win) * 1e3, t_copy / (done / win) * 1e3,
                           done / t_copy / 1e9, t_unreg / (done / win) * 1e3, done / (t_reg + t_copy + t_unreg) / 1e9);
    // threaded registration: nthreads threads register different windows at the same time
    if (ok) {
      double t0 = now();
      std::vector<std::thread> th;
      for (int t = 0; t < nthreads; t++) th.emplace_back([=] {
        for (size_t off = (size_t)t * win; off < n; off += (size_t)nthreads * win) { cudaHostRegister(m + off, win, flags); }
      });
      for (auto& t : th) t.join();
      double t1 = now();
      th.clear();
      for (int t = 0; t < nthreads; t++) th.emplace_back([=] {
        for (size_t off = (size_t)t * win; off < n; off += (size_t)nthreads * win) { cudaHostUnregister(m + off); }
      });
      for (auto& t : th) t.join();
      double t2 = now();
      printf("mode %d: %d threads registering all windows concurrently: %.1f GB/s; unregistering: %.1f GB/s\n", mode, nthreads, n / (t1 - t0) / 1e9, n / (t2 - t1) / 1e9);
    }
    cudaFree(d);
    munmap(m, n);
    close(fd);
  }
  return 0;
}

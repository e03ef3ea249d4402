// Developer measurement: which GPUs of a box share a host->device bottleneck?  One process, one thread per GPU, pinned
// source buffers, barrier-synchronised back-to-back cudaMemcpyAsync H2D copies on the chosen subset of GPUs.
//   nvcc -O2 -std=c++17 -o scripts/h2d_matrix scripts/h2d_matrix.cu -lpthread
//   scripts/h2d_matrix [MB per copy = 512] [copies = 4]
#include <cuda_runtime.h>
#include <sys/mman.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Gpu {
  char* h = nullptr;      // cudaHostAlloc
  char* h_thp = nullptr;  // mmap + MADV_HUGEPAGE + cudaHostRegister
  char* d = nullptr;
  cudaStream_t st = nullptr;
};

int main(int argc, char** argv) {
  const size_t mb = argc > 1 ? (size_t)atoll(argv[1]) : 512, bytes = mb << 20;
  const int reps = argc > 2 ? atoi(argv[2]) : 4;
  int n = 0;
  cudaGetDeviceCount(&n);
  printf("%d GPUs, %zu MB per copy, %d copies per test\n", n, mb, reps);
  std::vector<Gpu> g((size_t)n);
  for (int i = 0; i < n; i++) {
    cudaSetDevice(i);
    cudaHostAlloc(&g[i].h, bytes, cudaHostAllocPortable);
    memset(g[i].h, 1, bytes);
    void* p = mmap(nullptr, bytes + (2 << 20), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p != MAP_FAILED) {
      char* a = (char*)(((uintptr_t)p + (2 << 20) - 1) & ~(uintptr_t)((2 << 20) - 1));
      madvise(a, bytes, MADV_HUGEPAGE);
      memset(a, 2, bytes);
      if (cudaHostRegister(a, bytes, cudaHostRegisterPortable) == cudaSuccess) g[i].h_thp = a;
      else cudaGetLastError();
    }
    cudaMalloc(&g[i].d, bytes);
    cudaStreamCreateWithFlags(&g[i].st, cudaStreamNonBlocking);
  }
  // mode 0: H2D from the GPU's own cudaHostAlloc buffer; 1: H2D from the THP-backed registered buffer; 2: H2D, every GPU
  // reads GPU 0's buffer; 3: D2H into the own buffer
  auto run = [&](const std::vector<int>& set, int mode, const char* label) {
    std::atomic<int> ready{0};
    std::atomic<bool> go{false};
    std::vector<double> secs(set.size());
    std::vector<std::thread> th;
    for (size_t k = 0; k < set.size(); k++) {
      th.emplace_back([&, k] {
        const int i = set[k];
        cudaSetDevice(i);
        const char* src = mode == 1 ? (g[i].h_thp ? g[i].h_thp : g[i].h) : mode == 2 ? g[0].h : g[i].h;
        cudaMemcpyAsync(g[i].d, src, 1 << 20, cudaMemcpyHostToDevice, g[i].st);   // warm
        cudaStreamSynchronize(g[i].st);
        ready++;
        while (!go.load()) {}
        const double t0 = now();
        for (int r = 0; r < reps; r++) {
          if (mode == 3) cudaMemcpyAsync(g[i].h, g[i].d, bytes, cudaMemcpyDeviceToHost, g[i].st);
          else cudaMemcpyAsync(g[i].d, src, bytes, cudaMemcpyHostToDevice, g[i].st);
        }
        cudaStreamSynchronize(g[i].st);
        secs[k] = now() - t0;
      });
    }
    while (ready.load() < (int)set.size()) {}
    go.store(true);
    for (auto& t : th) t.join();
    double sum = 0;
    std::string per;
    for (size_t k = 0; k < set.size(); k++) {
      const double gbs = (double)reps * bytes / secs[k] / 1e9;
      sum += gbs;
      char buf[64];
      snprintf(buf, sizeof buf, " g%d=%.1f", set[k], gbs);
      per += buf;
    }
    printf("%-28s total %6.1f GB/s |%s\n", label, sum, per.c_str());
    fflush(stdout);
  };
  auto name = [](const char* what, const std::vector<int>& s) {
    std::string r = what;
    r += " {";
    for (size_t i = 0; i < s.size(); i++) r += (i ? "," : "") + std::to_string(s[i]);
    return r + "}";
  };
  std::vector<std::vector<int>> sets;
  for (int i = 0; i < n; i++) sets.push_back({i});
  for (int j = 1; j < n; j++) sets.push_back({0, j});
  if (n >= 4) { sets.push_back({1, 2}); sets.push_back({2, 3}); sets.push_back({0, 1, 2, 3}); }
  if (n >= 8) {
    sets.push_back({4, 5}); sets.push_back({6, 7}); sets.push_back({4, 5, 6, 7}); sets.push_back({0, 1, 4, 5});
    sets.push_back({0, 2, 4, 6}); sets.push_back({1, 3, 5, 7}); sets.push_back({2, 3, 6, 7}); sets.push_back({0, 3, 4, 7});
    sets.push_back({0, 1, 2, 3, 4, 5, 6, 7});
  }
  for (auto& s : sets) run(s, 0, name("H2D", s).c_str());
  std::vector<int> all;
  for (int i = 0; i < n; i++) all.push_back(i);
  run(all, 0, "H2D all (again)");
  run(all, 1, "H2D all, THP-backed source");
  run(all, 2, "H2D all, one shared source");
  run(all, 3, "D2H all");
  run({0}, 1, "H2D {0}, THP-backed source");
  run({0}, 3, "D2H {0}");
  return 0;
}

// C-ABI of the B200-native traj2dEcsv hot path (include/cgromcorr.h): context, streaming pipeline, writers.
#include "../../include/cgromcorr.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <fcntl.h>
#include <nccl.h>     // types and enums only: the functions are bound with dlopen (cgc_cov_allreduce)
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <map>
#include <mutex>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <nvtx3/nvToolsExt.h>   // header-only; ranges cost nothing unless a profiler is attached

#include "fmt_g.h"
#include "host_parse.h"
#include "kernels.cuh"

using namespace cgc;

namespace {

thread_local std::string g_create_error;

// Guard bands (CGC_GUARD=1 in the environment; developer aid): every device buffer a context owns is allocated with 4 KiB
// of a known byte pattern in front of it and behind it.  cgc_debug_guard_check reads the bands back: a kernel that wrote
// before or past one of its buffers has changed them.  (compute-sanitizer is closed on the B200 pool this was developed
// on; this covers the out-of-bounds WRITES of the hot-path kernels on all their buffers, at full speed.)
constexpr size_t kGuardBytes = 4096;
constexpr unsigned char kGuardByte = 0xA5;
struct GuardInfo { char* base; size_t bytes, padded; int device; };   // bytes: what was asked for; padded: rounded up to 256
std::mutex g_guard_mu;
std::map<void*, GuardInfo> g_guard_map;
bool guard_on() {
  static const bool on = [] { const char* e = getenv("CGC_GUARD"); return e && *e && *e != '0'; }();
  return on;
}
cudaError_t guard_malloc(void** p, size_t bytes) {
  if (!guard_on()) return cudaMalloc(p, bytes);
  const size_t padded = (bytes + 255) & ~(size_t)255;   // the band behind the buffer starts right after its (256-byte rounded) end
  char* base = nullptr;
  cudaError_t e = cudaMalloc(&base, padded + 2 * kGuardBytes);
  if (e != cudaSuccess) return e;
  e = cudaMemset(base, kGuardByte, padded + 2 * kGuardBytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();   // the fill runs on the legacy stream; the owners use non-blocking streams
  if (e != cudaSuccess) return e;
  int dev = 0;
  cudaGetDevice(&dev);
  *p = base + kGuardBytes;
  std::lock_guard<std::mutex> g(g_guard_mu);
  g_guard_map[*p] = {base, bytes, padded, dev};
  return cudaSuccess;
}
cudaError_t guard_free(void* p) {
  if (!guard_on() || !p) return cudaFree(p);
  char* base = nullptr;
  {
    std::lock_guard<std::mutex> g(g_guard_mu);
    auto it = g_guard_map.find(p);
    if (it == g_guard_map.end()) return cudaFree(p);
    base = it->second.base;
    g_guard_map.erase(it);
  }
  return cudaFree(base);
}

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define CUDA_OK(expr)                                                                                         \
  do {                                                                                                        \
    cudaError_t e__ = (expr);                                                                                 \
    if (e__ != cudaSuccess) {                                                                                 \
      throw CudaError(std::string(#expr) + ": " + cudaGetErrorString(e__) + " (" __FILE__ ":" +               \
                      std::to_string(__LINE__) + ")");                                                        \
    }                                                                                                         \
  } while (0)

// Copies byte ranges on a few background threads (page cache / pageable memory -> pinned staging, pinned rows -> caller).
// A job is cut into chunks that the workers pull from one FIFO queue, so two jobs submitted one after the other finish in
// that order and every worker stays busy across the boundary between them.
class CopyPool {
 public:
  struct Job {
    std::atomic<int64_t> left{0};   // chunks not yet copied
  };
  explicit CopyPool(int n_threads) {
    // non-temporal stores for the staging copies: 6144 C2 frames from tmpfs in 586-611 ms against 629-832 ms with plain
    // memcpy (profiles/r02_cli_staging_nt_stores.log); CGC_STAGE_NT=0 switches back
    if (const char* e = getenv("CGC_STAGE_NT")) streaming_ = *e && *e != '0';
    for (int i = 0; i < std::max(1, n_threads); i++) th_.emplace_back([this] { loop(); });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> g(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  int threads() const { return (int)th_.size(); }
  // drop_src_pages: src is a file mapping that is read once — give its page-table entries back as soon as a chunk is
  // copied (MADV_DONTNEED on a shared file mapping keeps the page cache, it only unmaps), so that the cost of unmapping
  // gigabytes does not pile up at the end of the run.
  void submit(Job* job, char* dst, const char* src, int64_t n, bool drop_src_pages) {
    const int64_t kChunk = 2 << 20;
    const int64_t chunks = std::max<int64_t>(1, (n + kChunk - 1) / kChunk);
    job->left.store(chunks);
    {
      std::lock_guard<std::mutex> g(mu_);
      for (int64_t i = 0; i < chunks; i++) {
        const int64_t b = i * kChunk, e = std::min(n, b + kChunk);
        q_.push_back({job, dst + b, src + b, e - b, drop_src_pages});
      }
    }
    cv_.notify_all();
  }
  void wait(Job* job) {
    std::unique_lock<std::mutex> g(done_mu_);
    done_cv_.wait(g, [job] { return job->left.load() <= 0; });
  }
  void copy(char* dst, const char* src, int64_t n) {   // synchronous, for the caller's convenience
    if (n < (8 << 20)) {
      memcpy(dst, src, (size_t)n);
      return;
    }
    Job j;
    submit(&j, dst, src, n, false);
    wait(&j);
  }

 private:
  struct Item { Job* job; char* dst; const char* src; int64_t n; bool drop; };
  void loop() {
    for (;;) {
      Item it;
      {
        std::unique_lock<std::mutex> g(mu_);
        cv_.wait(g, [this] { return stop_ || !q_.empty(); });
        if (q_.empty()) return;
        it = q_.front();
        q_.pop_front();
      }
      cgc::stream_copy(it.dst, it.src, (size_t)it.n, streaming_);
      if (it.drop) {
        const uintptr_t a = ((uintptr_t)it.src + 4095) & ~(uintptr_t)4095, b = ((uintptr_t)it.src + (uintptr_t)it.n) & ~(uintptr_t)4095;
        if (b > a) madvise((void*)a, (size_t)(b - a), MADV_DONTNEED);
      }
      if (it.job->left.fetch_sub(1) == 1) {
        std::lock_guard<std::mutex> g(done_mu_);
        done_cv_.notify_all();
      }
    }
  }
  std::mutex mu_, done_mu_;
  std::condition_variable cv_, done_cv_;
  std::deque<Item> q_;
  bool stop_ = false;
  bool streaming_ = true;
  std::vector<std::thread> th_;
};

// Page-locked host memory for the large buffers (staging ring, rows, sink): an anonymous mapping advised to use transparent
// huge pages, touched by several threads, then registered with CUDA.  On the B200 box 128 MB cost 5 ms this way against
// 50 ms for cudaHostAlloc (which pins 4 KiB page by page and holds a driver-wide lock while it does: a helper thread that
// allocated the ring in the background stalled the streaming thread's launches for as long), at the same H2D rate
// (profiles/r02_pinned_alloc.txt).  Falls back to cudaHostAlloc when the mapping cannot be registered.
struct PinnedBlock {
  char* p = nullptr;
  void* map = nullptr;      // the mapping p lies in; null: cudaHostAlloc
  size_t map_bytes = 0;
};
PinnedBlock pinned_alloc(size_t bytes) {
  PinnedBlock b;
  const size_t kHuge = (size_t)2 << 20;
  if (bytes >= ((size_t)4 << 20)) {
    const size_t len = ((bytes + kHuge - 1) / kHuge) * kHuge;
    void* m = mmap(nullptr, len + kHuge, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (m != MAP_FAILED) {
      char* a = (char*)(((uintptr_t)m + kHuge - 1) & ~(uintptr_t)(kHuge - 1));
      madvise(a, len, MADV_HUGEPAGE);
      const int nt = (int)std::max<size_t>(1, std::min<size_t>(8, len / ((size_t)16 << 20)));
      std::vector<std::thread> th;
      for (int t = 0; t < nt; t++) {
        const size_t lo = len / kHuge * t / nt * kHuge, hi = len / kHuge * (t + 1) / nt * kHuge;
        th.emplace_back([a, lo, hi] { memset(a + lo, 0, hi - lo); });
      }
      for (auto& t : th) t.join();
      if (cudaHostRegister(a, len, cudaHostRegisterPortable) == cudaSuccess) {
        b.p = a;
        b.map = m;
        b.map_bytes = len + kHuge;
        return b;
      }
      cudaGetLastError();
      munmap(m, len + kHuge);
    }
  }
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
    throw CudaError(std::string("cudaHostAlloc of ") + std::to_string(bytes) + " bytes failed: " + cudaGetErrorString(cudaGetLastError()));
  }
  b.p = (char*)p;
  return b;
}
void pinned_free(PinnedBlock& b) {
  if (!b.p) return;
  if (b.map) {
    cudaHostUnregister(b.p);
    munmap(b.map, b.map_bytes);
  } else {
    cudaFreeHost(b.p);
  }
  b = PinnedBlock();
}

// One slot of the streaming ring (cgc_context::kSlots of them): text in, dE out.
struct Slot {
  char* h_text = nullptr;   // pinned staging (file path only)
  PinnedBlock h_text_block, h_dE_block;
  char* d_text = nullptr;
  int64_t* d_off = nullptr;
  int64_t* h_off = nullptr;  // pinned
  float* d_xyz8 = nullptr;
  void* d_xtc = nullptr;     // .xtc only: checkpoint table of K1x's index pass
  float4* d_sites = nullptr;
  double* d_acc = nullptr;   // K2's bins
  float* d_dE = nullptr;
  float* h_dE = nullptr;     // pinned
  // `.time` text of the batch, formatted on the device (cgc_run_write_time); allocated on first use
  char* d_ttext = nullptr;
  int64_t* d_rowoff = nullptr;  // rows + 1
  int64_t* h_tinfo = nullptr;   // pinned [2]: total bytes, format flags
  int* d_fmt_flag = nullptr;    // K5: this batch held a value outside the device formatter's range
  int64_t ttext_cap = 0;
  cudaEvent_t ev_copied = nullptr, ev_done = nullptr;
  cudaEvent_t ev_text = nullptr;       // .xtc: the H2D copy alone (ev_copied then follows the index pass)
  cudaEvent_t ev_text_out = nullptr;   // the D2H copy of d_ttext has finished (the next batch may overwrite d_ttext)
  bool text_out_pending = false;
  CopyPool::Job stage_job;   // staging copy of the batch planned into this slot
  bool staged_here = false;  // the H2D copy reads h_text (else the caller's pinned text directly)
  const char* src = nullptr; // where the H2D copy reads from
  int64_t nb = 0;            // bytes of text of the batch
  int frames = 0;            // frames of the batch planned / in flight
  int64_t out_frame = 0;     // where its rows go in the caller's dE block
  bool busy = false;         // enqueued on the device, not yet drained
};

}  // namespace

struct cgc_context {
  ChargeTable table;
  int device = -1;
  int sm_count = 148;
  mutable std::string err;

  // trajectory text
  const char* text = nullptr;
  int64_t nbytes = 0;
  bool mapped = false;
  int fd = -1;              // the trajectory file behind the mapping
  FrameLayout L;
  std::vector<FrameSpan> frames;  // frames located so far
  bool index_complete = false;
  bool opened = false;
  int trr_real = 0;               // 0: .gro text; 4 / 8: .trr with single / double precision coordinates
  bool xtc = false;               // GROMACS .xtc (compressed coordinates, frames of varying size)
  int64_t frame_span_max = 0;     // .trr / .xtc: largest distance between the coordinate data of consecutive frames

  // static per-atom data (host copies kept for the getters)
  std::vector<float> q_ground;    // table charge of the atom's own key
  std::vector<float> q_exc;       // dQ for chromophore atoms (their ",EXC" entry), 0 otherwise
  std::vector<int32_t> col;
  DeviceSystem sys;
  bool dev_static = false;

  // options
  int n_sites = 0;
  int cov_opt = -1;  // -1 auto
  int batch_frames = 0;

  // pipeline
  // batches in the ring: up to three staged ahead by the background copiers, up to three on the device (H2D copy, queued,
  // computing) — the host thread only enqueues and drains
  static constexpr int kSlots = 6;
  int ring_slots = kSlots;      // option "ring_slots" (2..kSlots): half of them staged ahead, half on the device
  Slot slot[kSlots];
  bool dev_pipeline = false;
  int alloc_batch = 0;
  int64_t alloc_text = 0;
  int alloc_slots = 0;          // slots that have their buffers (1 for the device-resident entry points, kSlots for the ring)
  cudaStream_t st_copy = nullptr, st_comp = nullptr, st_out = nullptr;
  cudaStream_t st_index = nullptr;   // .xtc: the index pass of K1x, beside the previous batch's kernels (high priority)
  int* d_err = nullptr;
  int* d_fmt_flags = nullptr;   // K5 flag word of the one-shot formatter calls (cgc_write_mtx); the ring has one per slot
  int64_t launches = 0;
  int stage_threads = 0;        // option "stage_threads"; 0 = default
  CopyPool* pool = nullptr;     // created on first use
  std::atomic<bool> stop_requested{false};   // cgc_request_stop
  // pinned buffers the `.time` sink writes from (the device text of a batch is copied into one of them; the writer thread
  // hands it back once the bytes are in the file), kept across calls
  static constexpr int kSinkBufs = 6;
  char* sink_buf[kSinkBufs] = {};
  PinnedBlock sink_block[kSinkBufs];
  cudaEvent_t sink_ev[kSinkBufs] = {};
  int64_t sink_cap = 0;

  // optional per-kernel timing (option "profile"): event pairs around each launch of the hot path
  bool profile = false;
  struct Timed { cudaEvent_t a, b; int kind; };
  std::vector<Timed> timed;            // recorded, not yet read
  std::vector<cudaEvent_t> ev_pool;    // recycled events
  double kernel_ms[5] = {0, 0, 0, 0, 0};  // K1 decode, K2 cdc, finish, K3 cov, [4] the H2D copies of the text
  int64_t kernel_n[5] = {0, 0, 0, 0, 0};

  // Correlate
  double* d_cov = nullptr;
  bool cov_bound = false;
  bool shift_set = false;
  float* d_compat = nullptr;  // float32 [2][S*G]
  bool compat_on = false;     // option "mtx_compat": keep the reference's literal float32 sums while running
  double* d_centred = nullptr;  // scratch: centred FP64 rows of the batch being accumulated
  int64_t centred_frames = 0;
  // rows waiting for K3: the streaming loop gathers the float32 rows of small batches here and runs the Correlate kernels
  // once per kCovGather frames (their write-back touches all S*G*G sums, whatever the batch size)
  float* d_cov_rows = nullptr;
  int64_t cov_rows_cap = 0, cov_rows_pending = 0;
  double host_cov_gather_s = 0, host_cov_flush_s = 0;   // host time spent enqueueing K3 work (CGC_TRACE)

  ~cgc_context();
};

namespace {

void release_text(cgc_context* c) {
  if (c->mapped && c->text) munmap((void*)c->text, (size_t)c->nbytes);
  if (c->fd >= 0) close(c->fd);
  c->fd = -1;
  c->text = nullptr;
  c->nbytes = 0;
  c->mapped = false;
}

void free_pipeline(cgc_context* c) {
  if (c->device < 0) return;
  cudaSetDevice(c->device);
  for (auto& s : c->slot) {
    pinned_free(s.h_text_block);
    if (s.d_text) guard_free(s.d_text);
    if (s.d_off) guard_free(s.d_off);
    if (s.h_off) cudaFreeHost(s.h_off);
    if (s.d_xyz8) guard_free(s.d_xyz8);
    if (s.d_xtc) guard_free(s.d_xtc);
    if (s.d_sites) guard_free(s.d_sites);
    if (s.d_acc) guard_free(s.d_acc);
    if (s.d_dE) guard_free(s.d_dE);
    pinned_free(s.h_dE_block);
    if (s.d_ttext) guard_free(s.d_ttext);
    if (s.d_rowoff) guard_free(s.d_rowoff);
    if (s.h_tinfo) cudaFreeHost(s.h_tinfo);
    if (s.d_fmt_flag) guard_free(s.d_fmt_flag);
    if (s.ev_copied) cudaEventDestroy(s.ev_copied);
    if (s.ev_text) cudaEventDestroy(s.ev_text);
    if (s.ev_done) cudaEventDestroy(s.ev_done);
    if (s.ev_text_out) cudaEventDestroy(s.ev_text_out);
    s.h_text = nullptr; s.d_text = nullptr; s.d_off = nullptr; s.h_off = nullptr; s.d_xyz8 = nullptr; s.d_xtc = nullptr; s.d_sites = nullptr;
    s.d_acc = nullptr; s.d_dE = nullptr; s.h_dE = nullptr; s.d_ttext = nullptr; s.d_rowoff = nullptr; s.h_tinfo = nullptr;
    s.d_fmt_flag = nullptr; s.ttext_cap = 0; s.ev_copied = s.ev_done = s.ev_text_out = s.ev_text = nullptr;
    s.text_out_pending = false; s.staged_here = false; s.src = nullptr; s.nb = 0; s.frames = 0; s.out_frame = 0; s.busy = false;
  }
  for (int i = 0; i < cgc_context::kSinkBufs; i++) {
    pinned_free(c->sink_block[i]);
    if (c->sink_ev[i]) cudaEventDestroy(c->sink_ev[i]);
    c->sink_buf[i] = nullptr;
    c->sink_ev[i] = nullptr;
  }
  c->sink_cap = 0;
  c->dev_pipeline = false;
  c->alloc_batch = 0;
  c->alloc_text = 0;
  c->alloc_slots = 0;
}

void free_static(cgc_context* c) {
  if (c->device < 0) return;
  cudaSetDevice(c->device);
  if (c->sys.kq) guard_free(c->sys.kq);
  if (c->sys.col) guard_free(c->sys.col);
  if (c->sys.slot) guard_free(c->sys.slot);
  if (c->sys.slot_dq) guard_free(c->sys.slot_dq);
  c->sys.kq = nullptr; c->sys.col = nullptr; c->sys.slot = nullptr; c->sys.slot_dq = nullptr;
  if (c->d_cov && !c->cov_bound) guard_free(c->d_cov);
  c->d_cov = nullptr;
  c->cov_bound = false;
  if (c->d_compat) guard_free(c->d_compat);
  c->d_compat = nullptr;
  if (c->d_centred) guard_free(c->d_centred);
  c->d_centred = nullptr;
  c->centred_frames = 0;
  if (c->d_cov_rows) guard_free(c->d_cov_rows);
  c->d_cov_rows = nullptr;
  c->cov_rows_cap = c->cov_rows_pending = 0;
  c->shift_set = false;
  c->dev_static = false;
}

}  // namespace

cgc_context::~cgc_context() {
  delete pool;
  pool = nullptr;
  free_pipeline(this);
  free_static(this);
  if (device >= 0) {
    for (auto& t : timed) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
    for (auto& e : ev_pool) cudaEventDestroy(e);
    if (st_copy) cudaStreamDestroy(st_copy);
    if (st_index) cudaStreamDestroy(st_index);
    if (st_comp) cudaStreamDestroy(st_comp);
    if (st_out) cudaStreamDestroy(st_out);
    if (d_fmt_flags) guard_free(d_fmt_flags);
    if (d_err) guard_free(d_err);
  }
  release_text(this);
}

namespace {

template <typename F>
int guarded(cgc_context* c, F&& fn) {
  try {
    return fn();
  } catch (const InputFileMalformed& e) {
    if (c) c->err = e.what(); else g_create_error = e.what();
    return CGC_ERR_INPUT_MALFORMED;
  } catch (const InputFileMinorMalformed& e) {
    if (c) c->err = e.what(); else g_create_error = e.what();
    return CGC_ERR_INPUT_MINOR;
  } catch (const ReadError& e) {
    if (c) c->err = e.what(); else g_create_error = e.what();
    return CGC_ERR_READ;
  } catch (const Unsupported& e) {
    if (c) c->err = e.what(); else g_create_error = e.what();
    return CGC_ERR_UNSUPPORTED;
  } catch (const CudaError& e) {
    if (c) c->err = e.what(); else g_create_error = e.what();
    return CGC_ERR_CUDA;
  } catch (const std::exception& e) {
    if (c) c->err = e.what(); else g_create_error = e.what();
    return CGC_ERR_ARG;
  }
}

// Set-up-time zeroing: cudaMemset of device memory returns before the fill has run, on the legacy stream, and the contexts'
// streams are non-blocking (not ordered after it): wait, so that nothing enqueued later on any stream can overtake the fill.
cudaError_t zero_now(void* p, size_t bytes) {
  cudaError_t e = cudaMemset(p, 0, bytes);
  return e == cudaSuccess ? cudaDeviceSynchronize() : e;
}

int fail(const cgc_context* c, int code, const std::string& msg) {
  if (c) c->err = msg; else g_create_error = msg;
  return code;
}

void need_device(cgc_context* c) {
  if (c->device < 0) throw CudaError("this context was created without a CUDA device (device < 0); there is no CPU path");
  CUDA_OK(cudaSetDevice(c->device));
}

// Everything AtomDataLookup_v2 (reference src/grom2atomFMO.cpp:549-610) resolves, done once: per-atom ground charge,
// per-atom dQ for chromophore atoms, dE column per atom.
void resolve_atoms(cgc_context* c) {
  const FrameLayout& L = c->L;
  const int n = L.n_atoms;
  c->q_ground.assign(n, 0.f);
  c->q_exc.assign(n, 0.f);
  c->col.assign(n, -1);
  for (int a = 0; a < n; a++) {
    const std::string& key = L.keys[a];
    int j = c->table.find(key);
    if (j < 0) throw InputFileMalformed("Atom name " + key + " from input file not found in lookup table.");
    c->q_ground[a] = c->table.charges[j];
    const int g = L.groups[a];
    if (g < 0) {
      // the reference looks the ",EXC" entry up whenever this atom's chromophore is the excited site (:579-596)
      int k = c->table.find(key + ",EXC");
      if (k < 0) throw InputFileMalformed("Atom name " + key + ",EXC from input file not found in lookup table.");
      c->q_exc[a] = c->table.charges[k];
      // the reference's outer loop runs over 1..nGroups, so a chromophore index beyond nGroups is never visited
      // (reference src/grom2atomFMO.cpp:669,674)
      c->col[a] = (-g <= L.n_groups) ? (-g - 1) : -1;
    } else {
      c->col[a] = L.n_chromo + g - 1;
    }
  }
}

void build_static(cgc_context* c) {
  need_device(c);
  free_pipeline(c);
  free_static(c);
  const FrameLayout& L = c->L;
  DeviceSystem& sys = c->sys;
  const int n = L.n_atoms;
  sys.n_atoms = n;
  sys.n_pad = (n + kTileAtoms - 1) / kTileAtoms * kTileAtoms;
  sys.n_chromo = L.n_chromo;
  sys.n_groups = L.n_groups;
  sys.num_tot = L.n_chromo + L.n_groups;
  sys.n_sites = c->n_sites;
  sys.line_bytes = L.line_bytes;
  sys.x_col = L.x_col;
  sys.field_w = L.field_w;
  sys.ndec = L.ndec;
  sys.traj_mode = c->xtc ? 3 : c->trr_real == 4 ? 1 : c->trr_real == 8 ? 2 : 0;
  // dQ != 0 atoms of every computed site, in file order (zero-dQ atoms add exactly +-0 to every bin)
  std::vector<int> count(sys.n_sites, 0);
  std::vector<int32_t> slot(sys.n_pad, -1);
  for (int a = 0; a < n; a++) {
    int g = L.groups[a];
    if (g < 0 && -g <= sys.n_sites && c->q_exc[a] != 0.f) count[-g - 1]++;
  }
  int P = 1;
  for (int s = 0; s < sys.n_sites; s++) P = std::max(P, count[s]);
  P = cdc_pad_site_cap(P);
  sys.site_cap = P;
  if (cdc_smem_bytes(sys) > 200 * 1024) {
    throw Unsupported("site block (8 sites x dQ atoms x 16 B, a ring of 4) exceeds shared memory: too many dQ != 0 atoms per chromophore");
  }
  std::vector<float> slot_dq((size_t)sys.n_sites * P, 0.f);
  std::fill(count.begin(), count.end(), 0);
  for (int a = 0; a < n; a++) {
    int g = L.groups[a];
    if (g < 0 && -g <= sys.n_sites && c->q_exc[a] != 0.f) {
      int s = -g - 1;
      slot[a] = s * P + count[s];
      slot_dq[(size_t)s * P + count[s]] = c->q_exc[a];
      count[s]++;
    }
  }
  std::vector<double> kq(sys.n_pad, 0.0);
  std::vector<int32_t> col(sys.n_pad, -1);
  for (int a = 0; a < n; a++) {
    kq[a] = (double)33.2f * (double)c->q_ground[a];  // reference src/grom2atomFMO.cpp:666,687
    col[a] = c->col[a];
  }
  CUDA_OK(guard_malloc((void**)&sys.kq, sizeof(double) * sys.n_pad));
  CUDA_OK(guard_malloc((void**)&sys.col, sizeof(int32_t) * sys.n_pad));
  CUDA_OK(guard_malloc((void**)&sys.slot, sizeof(int32_t) * sys.n_pad));
  CUDA_OK(guard_malloc((void**)&sys.slot_dq, sizeof(float) * slot_dq.size()));
  CUDA_OK(cudaMemcpy(sys.kq, kq.data(), sizeof(double) * sys.n_pad, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(sys.col, col.data(), sizeof(int32_t) * sys.n_pad, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(sys.slot, slot.data(), sizeof(int32_t) * sys.n_pad, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(sys.slot_dq, slot_dq.data(), sizeof(float) * slot_dq.size(), cudaMemcpyHostToDevice));
  if (!c->st_copy) CUDA_OK(cudaStreamCreateWithFlags(&c->st_copy, cudaStreamNonBlocking));
  if (!c->st_index) {
    int least = 0, greatest = 0;
    CUDA_OK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    CUDA_OK(cudaStreamCreateWithPriority(&c->st_index, cudaStreamNonBlocking, greatest));
  }
  if (!c->st_comp) CUDA_OK(cudaStreamCreateWithFlags(&c->st_comp, cudaStreamNonBlocking));
  if (!c->st_out) CUDA_OK(cudaStreamCreateWithFlags(&c->st_out, cudaStreamNonBlocking));
  if (!c->d_fmt_flags) {
    CUDA_OK(guard_malloc((void**)&c->d_fmt_flags, sizeof(int)));
    CUDA_OK(zero_now(c->d_fmt_flags, sizeof(int)));
  }
  if (!c->d_err) {
    CUDA_OK(guard_malloc((void**)&c->d_err, sizeof(int)));
    CUDA_OK(zero_now(c->d_err, sizeof(int)));
  }
  c->dev_static = true;
}

int64_t cov_doubles(const cgc_context* c) {
  const int64_t SG = (int64_t)c->sys.n_sites * c->sys.num_tot;
  return 1 + 2 * SG + SG * c->sys.num_tot;
}

bool cov_enabled(const cgc_context* c) {
  if (c->cov_opt == 0) return false;
  if (c->cov_opt == 1) return true;
  return cov_doubles(c) * 8 <= ((int64_t)8 << 30);  // auto: on while the partial sums stay under 8 GiB
}

void ensure_cov(cgc_context* c) {
  if (!cov_enabled(c)) return;
  const int64_t SG = (int64_t)c->sys.n_sites * c->sys.num_tot;
  if (!c->d_cov) {
    CUDA_OK(guard_malloc((void**)&c->d_cov, sizeof(double) * cov_doubles(c)));
    CUDA_OK(zero_now(c->d_cov, sizeof(double) * cov_doubles(c)));
    c->cov_bound = false;
    c->shift_set = false;
  }
  if (!c->d_compat) {
    CUDA_OK(guard_malloc((void**)&c->d_compat, sizeof(float) * 2 * SG));
    CUDA_OK(zero_now(c->d_compat, sizeof(float) * 2 * SG));
  }
}

int default_batch(const cgc_context* c) {
  // about 128 MB of trajectory per batch (C2: 29 frames = 4263 K2 work units over 296 CTAs, a 2.4 ms H2D copy).  Measured
  // on the B200 box (2048 C2 frames from a file in the page cache): 192 MB x 4 slots 234 ms, 64 MB x 6 slots 324 ms — each
  // batch costs the host thread a fixed ~0.3 ms of enqueueing and two synchronisations
  const int64_t frame_text = (int64_t)c->L.n_atoms * c->L.line_bytes + 256;
  int64_t b = ((int64_t)128 << 20) / frame_text;
  // the FP64 bins of a batch should stay small next to the text
  const int64_t per_frame_out = (int64_t)c->sys.n_sites * c->sys.num_tot * 12;
  b = std::min<int64_t>(b, ((int64_t)2 << 30) / std::max<int64_t>(1, per_frame_out));
  // .xtc: the index pass of K1x takes the same time for 32 frames as for 512 (one warp per frame, all at once), so its
  // batches are as large as the buffers sensibly allow
  return (int)std::max<int64_t>(1, std::min<int64_t>(c->xtc ? 512 : 256, b));
}

// n_slots: kSlots for the streaming ring (run_impl), 1 for the entry points that work on text already on the device (they
// use slot 0's scratch only — no pinned buffers, no text buffer of their own)
void ensure_pipeline(cgc_context* c, int batch, int64_t text_bytes, bool need_host_text, int n_slots = cgc_context::kSlots) {
  need_device(c);
  if (!c->dev_static) throw std::runtime_error("internal: static data missing");
  const DeviceSystem& sys = c->sys;
  if (c->dev_pipeline && c->alloc_batch >= batch && c->alloc_text >= text_bytes && c->alloc_slots >= n_slots &&
      (!need_host_text || c->slot[0].h_text)) {
    return;
  }
  free_pipeline(c);
  const int64_t sg = (int64_t)sys.n_sites * sys.num_tot;
  const bool ring = n_slots > 1;
  for (int si = 0; si < n_slots; si++) {
    Slot& s = c->slot[si];
    if (ring) CUDA_OK(guard_malloc((void**)&s.d_text, (size_t)text_bytes + 64));
    if (need_host_text) {
      s.h_text_block = pinned_alloc((size_t)text_bytes + 64);
      s.h_text = s.h_text_block.p;
    }
    CUDA_OK(guard_malloc((void**)&s.d_off, sizeof(int64_t) * batch));
    CUDA_OK(cudaHostAlloc(&s.h_off, sizeof(int64_t) * batch, cudaHostAllocDefault));
    CUDA_OK(guard_malloc((void**)&s.d_xyz8, sizeof(float) * 3 * (size_t)sys.n_pad * batch));
    if (sys.traj_mode == 3) CUDA_OK(guard_malloc(&s.d_xtc, xtc_scratch_bytes(sys, batch)));
    CUDA_OK(guard_malloc((void**)&s.d_sites, (size_t)16 * sys.n_sites * sys.site_cap * batch));
    CUDA_OK(guard_malloc((void**)&s.d_acc, sizeof(double) * sg * batch));
    CUDA_OK(cudaMemsetAsync(s.d_acc, 0, sizeof(double) * sg * batch, c->st_comp));
    if (ring) {
      CUDA_OK(guard_malloc((void**)&s.d_dE, sizeof(float) * sg * batch));
      s.h_dE_block = pinned_alloc(sizeof(float) * (size_t)(sg * batch));
      s.h_dE = (float*)s.h_dE_block.p;
    }
    CUDA_OK(cudaEventCreateWithFlags(&s.ev_copied, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&s.ev_text, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&s.ev_text_out, cudaEventDisableTiming));
    launch_init_sites(sys, s.d_sites, batch, c->st_comp);
    c->launches++;
  }
  CUDA_OK(cudaStreamSynchronize(c->st_comp));
  c->dev_pipeline = true;
  c->alloc_batch = batch;
  c->alloc_text = text_bytes;
  c->alloc_slots = n_slots;
}

CopyPool* copy_pool(cgc_context* c) {
  if (!c->pool) {
    // 16 copier threads moved 68-83 GB/s from a fresh file mapping into pinned memory on the B200 box, 8 threads about half
    // (profiles/r01_host_staging_paths.log); with several contexts in one process (one per GPU) the caller divides the
    // cores with the "stage_threads" option
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const int n = c->stage_threads > 0 ? c->stage_threads : (int)std::max(1u, std::min(16u, hw > 4 ? hw - 2 : hw));   // two cores stay free for this thread and the writer
    c->pool = new CopyPool(n);
  }
  return c->pool;
}

// Appends byte ranges to a file from a background thread.  Every range gets its file offset when it is queued (the sizes
// are known then) and is written with pwrite().  One thread: page-cache writes ran at 3.4-5.6 GB/s from one thread on the
// B200 box and no faster from four (they serialise on the inode), while the `.time` text of the C2 workload needs 240 kB
// per frame — 2.6 GB/s at 11k frames/s.  The bytes of a batch come either from one of the context's pinned sink buffers
// (filled by a D2H copy the writer waits for through `ready`; handed back to the free list once written) or from a
// heap string (the host-formatted fallback).
class TextSink {
 public:
  TextSink(int fd, cgc_context* c) : fd_(fd), c_(c) {
    const off_t end = lseek(fd, 0, SEEK_END);
    next_off_ = end < 0 ? 0 : (int64_t)end;
    for (int i = 0; i < cgc_context::kSinkBufs; i++) free_.push_back(i);
    th_ = std::thread([this] { loop(); });
  }
  ~TextSink() { finish(); }
  // a free pinned buffer (index into cgc_context::sink_buf); waits for the writer when all of them are queued
  int acquire() {
    std::unique_lock<std::mutex> g(mu_);
    free_cv_.wait(g, [this] { return !free_.empty(); });
    const int i = free_.back();
    free_.pop_back();
    return i;
  }
  void push_buffer(int buf, int64_t n) {
    {
      std::lock_guard<std::mutex> g(mu_);
      q_.push_back({buf, std::string(), n, next_off_});
      next_off_ += n;
    }
    cv_.notify_one();
  }
  void push_string(std::string&& text) {
    {
      std::lock_guard<std::mutex> g(mu_);
      const int64_t n = (int64_t)text.size();
      q_.push_back({-1, std::move(text), n, next_off_});
      next_off_ += n;
    }
    cv_.notify_one();
  }
  bool finish() {   // drains the queue, joins the thread; returns false when a write failed
    {
      std::lock_guard<std::mutex> g(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    if (th_.joinable()) th_.join();
    return ok_;
  }

 private:
  struct Item { int buf; std::string text; int64_t n; int64_t off; };
  void loop() {
    if (c_->device >= 0) cudaSetDevice(c_->device);
    for (;;) {
      Item it;
      {
        std::unique_lock<std::mutex> g(mu_);
        cv_.wait(g, [this] { return stop_ || !q_.empty(); });
        if (q_.empty()) return;
        it = std::move(q_.front());
        q_.pop_front();
      }
      const char* p = it.text.data();
      if (it.buf >= 0) {
        if (cudaEventSynchronize(c_->sink_ev[it.buf]) != cudaSuccess) ok_ = false;   // the D2H copy into the buffer
        p = c_->sink_buf[it.buf];
      }
      int64_t off = 0;
      while (ok_ && off < it.n) {
        ssize_t w = pwrite(fd_, p + off, (size_t)(it.n - off), (off_t)(it.off + off));
        if (w <= 0) ok_ = false; else off += w;
      }
      if (it.buf >= 0) {
        {
          std::lock_guard<std::mutex> g(mu_);
          free_.push_back(it.buf);
        }
        free_cv_.notify_one();
      }
    }
  }
  int fd_;
  cgc_context* c_;
  int64_t next_off_ = 0;
  std::mutex mu_;
  std::condition_variable cv_, free_cv_;
  std::deque<Item> q_;
  std::vector<int> free_;
  bool stop_ = false;
  std::atomic<bool> ok_{true};
  std::thread th_;
};

void ensure_sink_bufs(cgc_context* c, int64_t cap) {
  if (c->sink_cap >= cap) return;
  for (int i = 0; i < cgc_context::kSinkBufs; i++) {
    pinned_free(c->sink_block[i]);
    c->sink_buf[i] = nullptr;
    c->sink_block[i] = pinned_alloc((size_t)cap);
    c->sink_buf[i] = c->sink_block[i].p;
    if (!c->sink_ev[i]) CUDA_OK(cudaEventCreateWithFlags(&c->sink_ev[i], cudaEventDisableTiming));
  }
  c->sink_cap = cap;
}

void ensure_time_text(cgc_context* c, int batch) {
  const int64_t rows = (int64_t)batch * c->sys.n_sites;
  const int64_t cap = format_time_capacity(rows, c->sys.num_tot) + 64;
  for (auto& s : c->slot) {
    if (s.ttext_cap >= cap) continue;
    if (s.d_ttext) CUDA_OK(guard_free(s.d_ttext));
    if (s.d_rowoff) CUDA_OK(guard_free(s.d_rowoff));
    s.d_ttext = nullptr; s.d_rowoff = nullptr; s.ttext_cap = 0;
    CUDA_OK(guard_malloc((void**)&s.d_ttext, (size_t)cap));
    CUDA_OK(guard_malloc((void**)&s.d_rowoff, sizeof(int64_t) * (size_t)(rows + 1)));
    if (!s.h_tinfo) CUDA_OK(cudaHostAlloc(&s.h_tinfo, sizeof(int64_t) * 2, cudaHostAllocDefault));
    if (!s.d_fmt_flag) CUDA_OK(guard_malloc((void**)&s.d_fmt_flag, sizeof(int)));
    s.ttext_cap = cap;
  }
  ensure_sink_bufs(c, cap);
}

// "%g" of a float, as std::ostream << float prints it (precision 6)
inline int fmt_g(char* p, float v) { return fmt_g6(p, v); }

// format rows [r0, r1) of a float matrix into `out`
void format_rows(const float* m, int64_t r0, int64_t r1, int ncol, bool to_cm, std::string& out) {
  out.clear();
  out.reserve((size_t)(r1 - r0) * ncol * 10);
  char buf[40];
  for (int64_t r = r0; r < r1; r++) {
    const float* row = m + r * ncol;
    for (int j = 0; j < ncol; j++) {
      float v = row[j];
      if (to_cm) v = (float)((double)v * 349.7);  // reference src/FMOparse.cpp:151
      if (j) out.push_back(',');
      int n = fmt_g(buf, v);
      out.append(buf, (size_t)n);
    }
    out.push_back('\n');
  }
}

// make sure frames [0, upto] are located; returns the number of frames known
int64_t index_upto(cgc_context* c, int64_t upto) {
  while (!c->index_complete && (int64_t)c->frames.size() <= upto) {
    int64_t pos = c->frames.empty() ? 0 : c->frames.back().end;
    FrameSpan sp;
    bool trunc = false;
    if (c->xtc) {
      XtcFrame xf;
      if (!locate_xtc_frame(c->text, c->nbytes, pos, xf, &trunc)) {
        c->index_complete = true;
        break;
      }
      if (xf.n_atoms != c->L.n_atoms) {
        throw InputFileMalformed("Number of atoms changes between frames; the group layout of the structure file no longer applies.");
      }
      sp.begin = xf.begin;
      sp.atoms = xf.coords;
      sp.end = xf.end;
      if (!c->frames.empty()) c->frame_span_max = std::max(c->frame_span_max, sp.end - c->frames.back().atoms);
      c->frame_span_max = std::max(c->frame_span_max, sp.end - sp.begin);
    } else if (c->trr_real) {
      TrrFrame tf;
      if (!locate_trr_frame(c->text, c->nbytes, pos, tf, &trunc)) {
        c->index_complete = true;
        break;
      }
      if (tf.n_atoms != c->L.n_atoms) {
        throw InputFileMalformed("Number of atoms changes between frames; the group layout of the structure file no longer applies.");
      }
      if (tf.real_bytes != c->trr_real) throw Unsupported(".trr precision changes between frames");
      sp.begin = tf.begin;
      sp.atoms = tf.x_off;
      sp.end = tf.end;
      if (!c->frames.empty()) c->frame_span_max = std::max(c->frame_span_max, sp.atoms - c->frames.back().atoms);
      c->frame_span_max = std::max(c->frame_span_max, sp.end - sp.begin);
    } else if (!locate_frame(c->text, c->nbytes, pos, c->L.n_atoms, c->L.line_bytes, sp, nullptr, &trunc)) {
      c->index_complete = true;
      break;
    }
    c->frames.push_back(sp);
  }
  return (int64_t)c->frames.size();
}

cudaEvent_t take_event(cgc_context* c) {
  if (!c->ev_pool.empty()) {
    cudaEvent_t e = c->ev_pool.back();
    c->ev_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  CUDA_OK(cudaEventCreate(&e));
  return e;
}

struct ScopedKernelTimer {
  cgc_context* c;
  cudaStream_t st;
  cgc_context::Timed t{};
  bool on;
  ScopedKernelTimer(cgc_context* c_, int kind, cudaStream_t st_) : c(c_), st(st_), on(c_->profile) {
    // the reference's timer scopes FileRead / dE_compute / Correlate (src/FMOparse.cpp:33-36) as NVTX ranges
    static const char* const kNames[5] = {"cgc:FileRead(K1 decode)", "cgc:dE_compute(K2 CDC)", "cgc:dE_rows(finish)", "cgc:Correlate(K3)",
                                          "cgc:FileRead(H2D copy)"};
    nvtxRangePushA(kNames[kind < 5 ? kind : 0]);
    if (!on) return;
    t.a = take_event(c);
    t.b = take_event(c);
    t.kind = kind;
    cudaEventRecord(t.a, st);
  }
  ~ScopedKernelTimer() {
    nvtxRangePop();
    if (!on) return;
    cudaEventRecord(t.b, st);
    c->timed.push_back(t);
  }
};

void collect_kernel_times(cgc_context* c) {
  for (auto& t : c->timed) {
    float ms = 0.f;
    if (cudaEventSynchronize(t.b) == cudaSuccess && cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) {
      c->kernel_ms[t.kind] += ms;
      c->kernel_n[t.kind]++;
    }
    c->ev_pool.push_back(t.a);
    c->ev_pool.push_back(t.b);
  }
  c->timed.clear();
}

double* centred_scratch(cgc_context* c, int n_frames) {
  if (c->centred_frames < n_frames) {
    if (c->d_centred) CUDA_OK(guard_free(c->d_centred));  // synchronises: nothing in flight still reads it
    c->d_centred = nullptr;
    const int64_t want = std::max<int64_t>(n_frames, 64);
    CUDA_OK(guard_malloc((void**)&c->d_centred, sizeof(double) * (size_t)cov_centred_doubles((int)want, c->sys.n_sites, c->sys.num_tot)));
    c->centred_frames = want;
  }
  return c->d_centred;
}

constexpr int64_t kCovGather = 256;

// run K3 on the gathered rows
void flush_cov_rows(cgc_context* c, cudaStream_t st) {
  if (c->cov_rows_pending <= 0) return;
  const DeviceSystem& sys = c->sys;
  const int n = (int)c->cov_rows_pending;
  {
    ScopedKernelTimer t(c, 3, st);
    CUDA_OK(launch_cov_accumulate(c->d_cov_rows, n, c->d_cov, c->compat_on ? c->d_compat : nullptr, centred_scratch(c, n), sys.n_sites, sys.num_tot, st));
  }
  c->launches += 3;
  c->cov_rows_pending = 0;
}

// Room for the gather buffer and the centred scratch of K3, sized once for the largest launch the streaming loop can
// produce (kCovGather - 1 gathered frames + one more batch): growing either in the middle of a run would mean a cudaFree,
// i.e. a device-wide synchronisation with everything the loop has queued.
void ensure_cov_buffers(cgc_context* c, int batch) {
  if (!c->d_cov) return;
  const DeviceSystem& sys = c->sys;
  const int64_t sg = (int64_t)sys.n_sites * sys.num_tot;
  const int64_t cap = kCovGather + batch;
  if (c->cov_rows_cap < cap) {
    if (c->d_cov_rows) CUDA_OK(guard_free(c->d_cov_rows));
    c->d_cov_rows = nullptr;
    c->cov_rows_cap = 0;
    CUDA_OK(guard_malloc((void**)&c->d_cov_rows, sizeof(float) * (size_t)(cap * sg)));
    c->cov_rows_cap = cap;
  }
  centred_scratch(c, (int)cap);
}

// hand the rows of one batch to K3: directly when the batch is large, through the gather buffer otherwise
void accumulate_rows(cgc_context* c, const float* d_dE, int n_frames, cudaStream_t st) {
  const DeviceSystem& sys = c->sys;
  const int64_t sg = (int64_t)sys.n_sites * sys.num_tot;
  if (n_frames >= kCovGather && c->cov_rows_pending == 0) {
    ScopedKernelTimer t(c, 3, st);
    CUDA_OK(launch_cov_accumulate(d_dE, n_frames, c->d_cov, c->compat_on ? c->d_compat : nullptr, centred_scratch(c, n_frames), sys.n_sites, sys.num_tot, st));
    c->launches += 3;
    return;
  }
  if (c->cov_rows_cap < c->cov_rows_pending + n_frames) {   // not sized by ensure_cov_buffers (a caller outside the loop)
    flush_cov_rows(c, st);
    CUDA_OK(cudaStreamSynchronize(st));
    ensure_cov_buffers(c, n_frames);
  }
  auto wall = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t0 = wall();
  CUDA_OK(cudaMemcpyAsync(c->d_cov_rows + c->cov_rows_pending * sg, d_dE, sizeof(float) * (size_t)(n_frames * sg),
                          cudaMemcpyDeviceToDevice, st));
  c->cov_rows_pending += n_frames;
  const double t1 = wall();
  if (c->cov_rows_pending >= kCovGather) flush_cov_rows(c, st);
  c->host_cov_gather_s += t1 - t0;
  c->host_cov_flush_s += wall() - t1;
}

// K1 (or K1x) for one batch whose bytes are in d_text.  xtc_indexed: the streaming loop has run K1x's index pass already
void enqueue_decode(cgc_context* c, const char* d_text, int64_t text_bytes, const int64_t* d_off, int n_frames, const Slot& s,
                    cudaStream_t st, bool xtc_indexed = false) {
  ScopedKernelTimer t(c, 0, st);
  if (c->sys.traj_mode == 3) {
    if (xtc_indexed) {   // the streaming loop has run the index pass on its own stream already
      launch_xtc_unpack(c->sys, d_text, text_bytes, d_off, n_frames, s.d_xtc, s.d_xyz8, s.d_sites, c->d_err, st);
    } else {
      launch_xtc_decode(c->sys, d_text, text_bytes, d_off, n_frames, s.d_xtc, s.d_xyz8, s.d_sites, c->d_err, st);
      c->launches += 1;   // index pass + unpack pass: one launch more than the callers count for the decode step
    }
  } else {
    launch_decode(c->sys, d_text, text_bytes, d_off, n_frames, s.d_xyz8, s.d_sites, c->d_err, st);
  }
}

// K1 -> K2 -> finish (-> K3) for one batch whose text is already in d_text; everything on `st`.
// `s`: the slot whose scratch (coordinates, site blocks, bins, .xtc checkpoints) the batch uses
void enqueue_compute(cgc_context* c, const char* d_text, int64_t text_bytes, const int64_t* d_off, int n_frames,
                     const Slot& s, float* d_dE, bool accumulate, cudaStream_t st, bool xtc_indexed = false) {
  const DeviceSystem& sys = c->sys;
  float* d_xyz8 = s.d_xyz8;
  float4* d_sites = s.d_sites;
  double* d_acc = s.d_acc;
  enqueue_decode(c, d_text, text_bytes, d_off, n_frames, s, st, xtc_indexed);
  {
    ScopedKernelTimer t(c, 1, st);
    launch_cdc(sys, d_xyz8, d_sites, d_acc, n_frames, c->sm_count, st);
  }
  {
    ScopedKernelTimer t(c, 2, st);
    launch_finish(d_acc, d_dE, (int64_t)n_frames * sys.n_sites * sys.num_tot, st);
  }
  c->launches += 3;
  if (accumulate && c->d_cov) {
    if (!c->shift_set) {
      launch_cov_set_shift(d_dE, c->d_cov, sys.n_sites, sys.num_tot, st);
      c->shift_set = true;
      c->launches++;
    }
    accumulate_rows(c, d_dE, n_frames, st);
  }
  CUDA_OK(cudaGetLastError());
}

// A read-only mapping of a whole file, released on scope exit.
struct MappedFile {
  const char* p = nullptr;
  int64_t n = 0;
  explicit MappedFile(const char* path) {
    int fd = open(path, O_RDONLY);
    if (fd < 0) throw ReadError("Open failed on file passed in.");
    struct stat sb;
    if (fstat(fd, &sb) != 0 || sb.st_size == 0) {
      close(fd);
      throw ReadError("file is empty or cannot be stat'ed");
    }
    void* m = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (m == MAP_FAILED) throw ReadError("mmap failed");
    p = (const char*)m;
    n = sb.st_size;
  }
  ~MappedFile() { if (p) munmap((void*)p, (size_t)n); }
  MappedFile(const MappedFile&) = delete;
  MappedFile& operator=(const MappedFile&) = delete;
};

// structure_gro == nullptr: c->text is .gro text and frame 0 is its own structure.  Otherwise c->text is a .trr and
// the names, residue numbers and hence groups/keys come from the first frame of the .gro file `structure_gro`.
enum TrajKind { kTrajGro = 0, kTrajTrr = 1, kTrajXtc = 2 };
void open_common(cgc_context* c, const char* structure_gro = nullptr, TrajKind kind = kTrajGro) {
  c->frames.clear();
  c->index_complete = false;
  c->opened = false;
  c->trr_real = 0;
  c->xtc = false;
  c->frame_span_max = 0;
  FrameSpan sp;
  c->L = FrameLayout();
  if (structure_gro) {
    {
      MappedFile st(structure_gro);
      scan_first_frame(st.p, st.n, c->L, sp);
    }
    if (kind == kTrajXtc) {
      XtcFrame xf;
      bool cut = false;
      if (!locate_xtc_frame(c->text, c->nbytes, 0, xf, &cut)) {
        throw InputFileMalformed(cut ? "first frame of the .xtc trajectory is incomplete" : "empty .xtc trajectory");
      }
      if (xf.n_atoms != c->L.n_atoms) {
        throw InputFileMalformed("the structure file has " + std::to_string(c->L.n_atoms) + " atoms, the .xtc trajectory " +
                                 std::to_string(xf.n_atoms));
      }
      c->xtc = true;
      // the whole index up front (one header per frame): frames differ in size, the staging buffers take the largest
      index_upto(c, INT64_MAX - 1);
      // "bytes per atom" of the largest frame: what the batch sizing works with
      c->L.line_bytes = (int)std::max<int64_t>(1, (c->frame_span_max + c->L.n_atoms - 1) / c->L.n_atoms);
    } else {
      TrrFrame tf;
      bool trunc = false;
      if (!locate_trr_frame(c->text, c->nbytes, 0, tf, &trunc)) {
        throw InputFileMalformed(trunc ? "first frame of the .trr trajectory is incomplete" : "empty .trr trajectory");
      }
      if (tf.n_atoms != c->L.n_atoms) {
        throw InputFileMalformed("the structure file has " + std::to_string(c->L.n_atoms) + " atoms, the .trr trajectory " +
                                 std::to_string(tf.n_atoms));
      }
      c->trr_real = tf.real_bytes;
      c->L.line_bytes = 3 * tf.real_bytes;   // one "line" = the coordinates of one atom
      // the whole index up front (one header per frame): the staging buffers are sized from the largest frame
      index_upto(c, INT64_MAX - 1);
    }
  } else {
    scan_first_frame(c->text, c->nbytes, c->L, sp);
    c->frames.push_back(sp);
  }
  resolve_atoms(c);
  c->n_sites = c->L.n_chromo;
  c->compat_on = false;
  c->cov_opt = -1;
  c->batch_frames = 0;
  if (c->L.n_chromo < 1) throw InputFileMalformed("no BCL/BCX residue in the trajectory: nothing to excite");
  if (c->device >= 0) build_static(c);
  c->opened = true;
}

}  // namespace

// =====================================================================================================================
extern "C" {

const char* cgc_version(void) { return "cgromcorr-b200 0.1 (sm_100a)"; }

const char* cgc_last_error(const cgc_context* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int cgc_create(const char* topology_path, const char* excited_itp_path, int device, cgc_context** out) {
  if (!out || !topology_path || !excited_itp_path) return fail(nullptr, CGC_ERR_ARG, "null argument");
  *out = nullptr;
  cgc_context* c = new cgc_context();
  int rc = guarded(nullptr, [&]() {
    load_tables(topology_path, excited_itp_path, c->table);
    c->device = device;
    if (device >= 0) {
      int count = 0;
      CUDA_OK(cudaGetDeviceCount(&count));
      if (device >= count) throw CudaError("CUDA device " + std::to_string(device) + " not present");
      CUDA_OK(cudaSetDevice(device));
      cudaDeviceProp prop;
      CUDA_OK(cudaGetDeviceProperties(&prop, device));
      if (prop.major < 10) throw CudaError(std::string("device '") + prop.name + "' is not sm_100-class; this library is built for sm_100a only");
      c->sm_count = prop.multiProcessorCount;
    }
    return CGC_OK;
  });
  if (rc != CGC_OK) {
    c->device = -1;
    delete c;
    return rc;
  }
  *out = c;
  return CGC_OK;
}

void cgc_destroy(cgc_context* ctx) { delete ctx; }

int cgc_table_size(const cgc_context* ctx) { return ctx ? (int)ctx->table.names.size() : 0; }

int cgc_table_entry(const cgc_context* ctx, int i, char* name, int name_cap, float* mass, float* charge) {
  if (!ctx || i < 0 || i >= (int)ctx->table.names.size()) return fail(ctx, CGC_ERR_ARG, "table index out of range");
  if (name && name_cap > 0) {
    snprintf(name, (size_t)name_cap, "%s", ctx->table.names[i].c_str());
  }
  if (mass) *mass = ctx->table.masses[i];
  if (charge) *charge = ctx->table.charges[i];
  return CGC_OK;
}

int cgc_topology_warnings(const cgc_context* ctx) { return ctx ? ctx->table.warnings : 0; }

int cgc_open(cgc_context* ctx, const char* traj_path) {
  if (!ctx || !traj_path) return fail(ctx, CGC_ERR_ARG, "null argument");
  return guarded(ctx, [&]() {
    if (ctx->device >= 0) { free_pipeline(ctx); }
    release_text(ctx);
    int fd = open(traj_path, O_RDONLY);
    if (fd < 0) throw ReadError("Open failed on file passed in.");  // reference src/grom2atomFMO.cpp:495-497
    struct stat sb;
    if (fstat(fd, &sb) != 0 || sb.st_size == 0) {
      close(fd);
      throw ReadError("trajectory file is empty or cannot be stat'ed");
    }
    void* p = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_SHARED, fd, 0);
    if (p == MAP_FAILED) {
      close(fd);
      throw ReadError("mmap of the trajectory failed");
    }
    // the mapping serves the frame-0 pass, the frame index, and the bulk text: several threads memcpy it into the pinned
    // staging buffers (see CopyPool)
    ctx->text = (const char*)p;
    ctx->nbytes = sb.st_size;
    ctx->mapped = true;
    ctx->fd = fd;
    open_common(ctx);
    return CGC_OK;
  });
}

int cgc_open_memory(cgc_context* ctx, const char* text, int64_t nbytes) {
  if (!ctx || !text || nbytes <= 0) return fail(ctx, CGC_ERR_ARG, "null or empty text");
  return guarded(ctx, [&]() {
    if (ctx->device >= 0) { free_pipeline(ctx); }
    release_text(ctx);
    ctx->text = text;
    ctx->nbytes = nbytes;
    ctx->mapped = false;
    // A caller-pinned buffer (cudaHostAlloc / cudaHostRegister / torch pin_memory) is detected in cgc_run and then
    // copied H2D directly, without the pinned staging copy.
    open_common(ctx);
    return CGC_OK;
  });
}

int cgc_open_trr(cgc_context* ctx, const char* trr_path, const char* structure_gro_path) {
  if (!ctx || !trr_path || !structure_gro_path) return fail(ctx, CGC_ERR_ARG, "null argument");
  return guarded(ctx, [&]() {
    if (ctx->device >= 0) { free_pipeline(ctx); }
    release_text(ctx);
    int fd = open(trr_path, O_RDONLY);
    if (fd < 0) throw ReadError("Open failed on file passed in.");
    struct stat sb;
    if (fstat(fd, &sb) != 0 || sb.st_size == 0) {
      close(fd);
      throw ReadError("trajectory file is empty or cannot be stat'ed");
    }
    void* p = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_SHARED, fd, 0);
    if (p == MAP_FAILED) {
      close(fd);
      throw ReadError("mmap of the trajectory failed");
    }
    ctx->text = (const char*)p;
    ctx->nbytes = sb.st_size;
    ctx->mapped = true;
    ctx->fd = fd;
    open_common(ctx, structure_gro_path, kTrajTrr);
    return CGC_OK;
  });
}

int cgc_open_memory_trr(cgc_context* ctx, const char* data, int64_t nbytes, const char* structure_gro_path) {
  if (!ctx || !data || nbytes <= 0 || !structure_gro_path) return fail(ctx, CGC_ERR_ARG, "null or empty argument");
  return guarded(ctx, [&]() {
    if (ctx->device >= 0) { free_pipeline(ctx); }
    release_text(ctx);
    ctx->text = data;
    ctx->nbytes = nbytes;
    ctx->mapped = false;
    open_common(ctx, structure_gro_path, kTrajTrr);
    return CGC_OK;
  });
}

int cgc_open_xtc(cgc_context* ctx, const char* xtc_path, const char* structure_gro_path) {
  if (!ctx || !xtc_path || !structure_gro_path) return fail(ctx, CGC_ERR_ARG, "null argument");
  return guarded(ctx, [&]() {
    if (ctx->device >= 0) { free_pipeline(ctx); }
    release_text(ctx);
    int fd = open(xtc_path, O_RDONLY);
    if (fd < 0) throw ReadError("Open failed on file passed in.");
    struct stat sb;
    if (fstat(fd, &sb) != 0 || sb.st_size == 0) {
      close(fd);
      throw ReadError("trajectory file is empty or cannot be stat'ed");
    }
    void* p = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_SHARED, fd, 0);
    if (p == MAP_FAILED) {
      close(fd);
      throw ReadError("mmap of the trajectory failed");
    }
    ctx->text = (const char*)p;
    ctx->nbytes = sb.st_size;
    ctx->mapped = true;
    ctx->fd = fd;
    open_common(ctx, structure_gro_path, kTrajXtc);
    return CGC_OK;
  });
}

int cgc_open_memory_xtc(cgc_context* ctx, const char* data, int64_t nbytes, const char* structure_gro_path) {
  if (!ctx || !data || nbytes <= 0 || !structure_gro_path) return fail(ctx, CGC_ERR_ARG, "null or empty argument");
  return guarded(ctx, [&]() {
    if (ctx->device >= 0) { free_pipeline(ctx); }
    release_text(ctx);
    ctx->text = data;
    ctx->nbytes = nbytes;
    ctx->mapped = false;
    open_common(ctx, structure_gro_path, kTrajXtc);
    return CGC_OK;
  });
}

int64_t cgc_xtc_encode(const int32_t* coords, int n_atoms, int step, float time_ps, const float* box9, float precision,
                       char* out, int64_t cap) {
  if (!coords || n_atoms < 1 || !(precision > 0.f) || (cap > 0 && !out)) {
    fail(nullptr, CGC_ERR_ARG, "bad argument");
    return -1;
  }
  int64_t n = -1;
  guarded(nullptr, [&]() {
    static const float zero_box[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<char> buf;
    buf.reserve((size_t)n_atoms * 8 + 128);
    xtc_encode_frame(coords, n_atoms, step, time_ps, box9 ? box9 : zero_box, precision, buf);
    n = (int64_t)buf.size();
    if (n <= cap) memcpy(out, buf.data(), (size_t)n);
    return CGC_OK;
  });
  return n;
}

int cgc_dims(const cgc_context* ctx, int* n_atoms, int* n_chromo, int* n_groups, int* num_tot) {
  if (!ctx || !ctx->opened) return fail(ctx, CGC_ERR_STATE, "no trajectory opened");
  if (n_atoms) *n_atoms = ctx->L.n_atoms;
  if (n_chromo) *n_chromo = ctx->L.n_chromo;
  if (n_groups) *n_groups = ctx->L.n_groups;
  if (num_tot) *num_tot = ctx->L.n_chromo + ctx->L.n_groups;
  return CGC_OK;
}

int cgc_get_groups(const cgc_context* ctx, int32_t* out) {
  if (!ctx || !ctx->opened || !out) return fail(ctx, CGC_ERR_STATE, "no trajectory opened");
  memcpy(out, ctx->L.groups.data(), sizeof(int32_t) * ctx->L.groups.size());
  return CGC_OK;
}

int cgc_get_columns(const cgc_context* ctx, int32_t* out) {
  if (!ctx || !ctx->opened || !out) return fail(ctx, CGC_ERR_STATE, "no trajectory opened");
  memcpy(out, ctx->col.data(), sizeof(int32_t) * ctx->col.size());
  return CGC_OK;
}

int cgc_get_site_charges(const cgc_context* ctx, int site, float* out) {
  if (!ctx || !ctx->opened || !out) return fail(ctx, CGC_ERR_STATE, "no trajectory opened");
  if (site < 0 || site > ctx->L.n_chromo) return fail(ctx, CGC_ERR_ARG, "site out of range");
  for (int a = 0; a < ctx->L.n_atoms; a++) {
    out[a] = (ctx->L.groups[a] == -site && site > 0) ? ctx->q_exc[a] : ctx->q_ground[a];
  }
  return CGC_OK;
}

int cgc_get_key(const cgc_context* ctx, int atom, char* buf, int cap) {
  if (!ctx || !ctx->opened || !buf || cap <= 0) return fail(ctx, CGC_ERR_STATE, "no trajectory opened");
  if (atom < 0 || atom >= ctx->L.n_atoms) return fail(ctx, CGC_ERR_ARG, "atom out of range");
  snprintf(buf, (size_t)cap, "%s", ctx->L.keys[atom].c_str());
  return CGC_OK;
}

int cgc_get_layout(const cgc_context* ctx, int* line_bytes, int* x_col, int* field_width, int* decimals) {
  if (!ctx || !ctx->opened) return fail(ctx, CGC_ERR_STATE, "no trajectory opened");
  if (line_bytes) *line_bytes = ctx->L.line_bytes;
  if (x_col) *x_col = ctx->L.x_col;
  if (field_width) *field_width = ctx->L.field_w;
  if (decimals) *decimals = ctx->L.ndec;
  return CGC_OK;
}

int cgc_frame_count(cgc_context* ctx, int64_t* n_frames) {
  if (!ctx || !ctx->opened || !n_frames) return fail(ctx, CGC_ERR_STATE, "no trajectory opened");
  return guarded(ctx, [&]() {
    *n_frames = index_upto(ctx, INT64_MAX - 1);
    return CGC_OK;
  });
}

int64_t cgc_frame_atoms_offset(cgc_context* ctx, int64_t frame) {
  if (!ctx || !ctx->opened || frame < 0) return -1;
  int64_t r = -1;
  guarded(ctx, [&]() {
    if (index_upto(ctx, frame) > frame) r = ctx->frames[(size_t)frame].atoms;
    return CGC_OK;
  });
  return r;
}

int cgc_set_option(cgc_context* ctx, const char* key, double value) {
  if (!ctx || !key) return fail(ctx, CGC_ERR_ARG, "null argument");
  if (!ctx->opened) return fail(ctx, CGC_ERR_STATE, "options are per trajectory: call cgc_open first");
  return guarded(ctx, [&]() {
    std::string k(key);
    if (k == "sites") {
      int v = (int)value;
      if (v < 1 || v > ctx->L.n_chromo) {
        // the reference asserts that the site exists (src/grom2atomFMO.cpp:632)
        throw std::invalid_argument("sites must be within 1.." + std::to_string(ctx->L.n_chromo));
      }
      if (v != ctx->n_sites) {
        ctx->n_sites = v;
        if (ctx->device >= 0) build_static(ctx);
      }
    } else if (k == "covariance") {
      ctx->cov_opt = value != 0.0 ? 1 : 0;
    } else if (k == "profile") {
      ctx->profile = value != 0.0;
    } else if (k == "mtx_compat") {
      ctx->compat_on = value != 0.0;
    } else if (k == "batch_frames") {
      if (value < 1 || value > 65535) throw std::invalid_argument("batch_frames must be within 1..65535");
      ctx->batch_frames = (int)value;
    } else if (k == "ring_slots") {
      if (value < 2 || value > cgc_context::kSlots) throw std::invalid_argument("ring_slots must be within 2..6");
      ctx->ring_slots = (int)value;
    } else if (k == "stage_threads") {
      if (value < 1 || value > 256) throw std::invalid_argument("stage_threads must be within 1..256");
      if ((int)value != ctx->stage_threads) {
        delete ctx->pool;   // idle between calls
        ctx->pool = nullptr;
        ctx->stage_threads = (int)value;
      }
    } else {
      throw std::invalid_argument("unknown option '" + k + "'");
    }
    return CGC_OK;
  });
}

double cgc_get_option(const cgc_context* ctx, const char* key) {
  if (!ctx || !key || !ctx->opened) return NAN;
  std::string k(key);
  if (k == "sites") return ctx->n_sites;
  if (k == "mtx_compat") return ctx->compat_on ? 1 : 0;
  if (k == "covariance") return ctx->device >= 0 && ctx->dev_static ? (cov_enabled(ctx) ? 1 : 0) : ctx->cov_opt;
  if (k == "batch_frames") return ctx->batch_frames > 0 ? ctx->batch_frames : default_batch(ctx);
  if (k == "ring_slots") return ctx->ring_slots;
  if (k == "stage_threads") return ctx->pool ? ctx->pool->threads() : ctx->stage_threads;
  if (k == "site_cap") return ctx->sys.site_cap;
  if (k == "n_pad") return ctx->sys.n_pad;
  return NAN;
}

int64_t cgc_launch_count(const cgc_context* ctx) { return ctx ? ctx->launches : 0; }

int cgc_debug_guard_check(int* n_buffers, int* n_damaged) {
  if (n_buffers) *n_buffers = 0;
  if (n_damaged) *n_damaged = 0;
  if (!guard_on()) return CGC_ERR_STATE;
  std::lock_guard<std::mutex> g(g_guard_mu);
  std::vector<unsigned char> band(kGuardBytes + 256);
  int bad = 0, n = 0, dev0 = 0;
  cudaGetDevice(&dev0);
  for (const auto& kv : g_guard_map) {
    const GuardInfo& gi = kv.second;
    if (cudaSetDevice(gi.device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) return CGC_ERR_CUDA;
    bool damaged = false;
    for (int side = 0; side < 2; side++) {   // behind the buffer the band starts at its very last byte + 1
      const char* src = side == 0 ? gi.base : gi.base + kGuardBytes + gi.bytes;
      const size_t len = side == 0 ? kGuardBytes : kGuardBytes + (gi.padded - gi.bytes);
      if (cudaMemcpy(band.data(), src, len, cudaMemcpyDeviceToHost) != cudaSuccess) return CGC_ERR_CUDA;
      for (size_t k = 0; k < len; k++) damaged = damaged || band[k] != kGuardByte;
    }
    n++;
    bad += damaged ? 1 : 0;
  }
  cudaSetDevice(dev0);
  if (n_buffers) *n_buffers = n;
  if (n_damaged) *n_damaged = bad;
  return CGC_OK;
}

int cgc_kernel_times(cgc_context* ctx, double* ms, int64_t* launches, int reset) {
  if (!ctx || !ms || !launches) return fail(ctx, CGC_ERR_ARG, "null argument");
  return guarded(ctx, [&]() {
    need_device(ctx);
    collect_kernel_times(ctx);
    for (int i = 0; i < 4; i++) {
      ms[i] = ctx->kernel_ms[i];
      launches[i] = ctx->kernel_n[i];
      if (reset) {
        ctx->kernel_ms[i] = 0;
        ctx->kernel_n[i] = 0;
      }
    }
    return CGC_OK;
  });
}

int cgc_copy_times(cgc_context* ctx, double* ms, int64_t* copies, int reset) {
  if (!ctx || !ms || !copies) return fail(ctx, CGC_ERR_ARG, "null argument");
  return guarded(ctx, [&]() {
    need_device(ctx);
    collect_kernel_times(ctx);
    *ms = ctx->kernel_ms[4];
    *copies = ctx->kernel_n[4];
    if (reset) {
      ctx->kernel_ms[4] = 0;
      ctx->kernel_n[4] = 0;
    }
    return CGC_OK;
  });
}

int cgc_decode_errors(cgc_context* ctx, int* flags) {
  if (!ctx || !flags) return fail(ctx, CGC_ERR_ARG, "null argument");
  return guarded(ctx, [&]() {
    need_device(ctx);
    *flags = 0;
    if (ctx->d_err) CUDA_OK(cudaMemcpy(flags, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost));
    return CGC_OK;
  });
}

// ---------------------------------------------------------------------------------------------------------------------
namespace {

// The streaming loop behind cgc_run and cgc_run_write_time.  sink != nullptr: the `.time` text of every batch is formatted
// on the device (K5) and appended through the sink.
//
// Three parties work on a ring of kSlots batches:
//   * the copier threads (CopyPool) move the text of batches b+1, b+2 from the file mapping / pageable memory into the
//     slots' pinned staging buffers (skipped when the caller's text is already page-locked),
//   * the device has up to kInFlight batches enqueued: H2D copy on st_copy, K1 -> K2 -> rows (-> K3, K5) on st_comp,
//   * this thread plans batches, hands staged ones to the device and drains finished ones (rows to the caller, device text
//     to a sink buffer by a D2H copy on st_out that the writer thread waits for).
// It blocks only on the staging of the batch it wants to enqueue next and on the oldest batch on the device.
int run_impl(cgc_context* ctx, int64_t first_frame, int64_t n_frames, float* dE_kcal, int64_t* frames_done, TextSink* sink) {
  if (frames_done) *frames_done = 0;
  if (!ctx) return CGC_ERR_ARG;
  if (!ctx->opened) return fail(ctx, CGC_ERR_STATE, "no trajectory opened");
  if (first_frame < 0 || n_frames < 0) return fail(ctx, CGC_ERR_ARG, "negative frame range");
  return guarded(ctx, [&]() {
    need_device(ctx);
    const DeviceSystem& sys = ctx->sys;
    const int K = ctx->ring_slots, kInFlight = K / 2, kStageAhead = K - kInFlight;
    const int B = ctx->batch_frames > 0 ? ctx->batch_frames : default_batch(ctx);
    const int64_t frame_text_max = std::max<int64_t>((int64_t)sys.n_atoms * sys.line_bytes, ctx->frame_span_max) + 4096;
    // is the text already page-locked (caller-pinned memory given to cgc_open_memory)?
    bool direct = false;
    if (!ctx->mapped) {
      cudaPointerAttributes at;
      direct = cudaPointerGetAttributes(&at, ctx->text) == cudaSuccess && at.type == cudaMemoryTypeHost;
      cudaGetLastError();
    }
    const double t_setup0 = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    auto wall = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    ensure_pipeline(ctx, B, frame_text_max * B, !direct, K);
    const double t_setup1 = wall();
    ensure_cov(ctx);
    ensure_cov_buffers(ctx, B);
    const double t_setup2 = wall();
    if (sink) ensure_time_text(ctx, B);
    CopyPool* pool = copy_pool(ctx);
    const double t_setup = wall() - t_setup0;
    // a previous call that ended in an error may have left batches on the device: nothing of theirs is delivered
    for (auto& sl : ctx->slot) {
      if (sl.busy || sl.text_out_pending) {
        CUDA_OK(cudaStreamSynchronize(ctx->st_copy));
        if (ctx->st_index) CUDA_OK(cudaStreamSynchronize(ctx->st_index));
        CUDA_OK(cudaStreamSynchronize(ctx->st_comp));
        CUDA_OK(cudaStreamSynchronize(ctx->st_out));
      }
      sl.busy = false;
      sl.text_out_pending = false;
    }
    ctx->cov_rows_pending = 0;
    ctx->host_cov_gather_s = ctx->host_cov_flush_s = 0;
    const int64_t sg = (int64_t)sys.n_sites * sys.num_tot;
    const bool want_rows = dE_kcal != nullptr || sink != nullptr;   // the text path keeps the rows for its host fallback
    // .xtc: index pass on its own stream (CGC_XTC_OVERLAP=0: both passes of K1x in the batch's turn on the compute stream)
    bool xtc_overlap = ctx->xtc;
    if (const char* e = getenv("CGC_XTC_OVERLAP")) xtc_overlap = xtc_overlap && *e && *e != '0';

    // The shift c of the Correlate sums is the dE row of the trajectory's frame 0 on EVERY rank (so partial sums of
    // different frame ranges add up).  A run that does not start at frame 0 evaluates frame 0 first, unaccumulated.
    if (ctx->d_cov && !ctx->shift_set && first_frame != 0) {
      Slot& s = ctx->slot[0];
      const FrameSpan& f0 = ctx->frames[0];
      const int64_t b0 = f0.atoms & ~(int64_t)15;
      const int64_t nb = (ctx->xtc ? f0.end : f0.atoms + (int64_t)sys.n_atoms * sys.line_bytes) - b0;
      const int64_t nb_clamped = std::min(nb, ctx->nbytes - b0);
      CUDA_OK(cudaMemcpyAsync(s.d_text, ctx->text + b0, (size_t)nb_clamped, cudaMemcpyHostToDevice, ctx->st_comp));
      s.h_off[0] = f0.atoms - b0;
      CUDA_OK(cudaMemcpyAsync(s.d_off, s.h_off, sizeof(int64_t), cudaMemcpyHostToDevice, ctx->st_comp));
      enqueue_compute(ctx, s.d_text, nb_clamped, s.d_off, 1, s, s.d_dE, false, ctx->st_comp);
      launch_cov_set_shift(s.d_dE, ctx->d_cov, sys.n_sites, sys.num_tot, ctx->st_comp);
      ctx->launches++;
      ctx->shift_set = true;
      CUDA_OK(cudaStreamSynchronize(ctx->st_comp));
    }

    // host-side time of this call by phase, printed when CGC_TRACE is set (developer aid)
    const bool trace = getenv("CGC_TRACE") != nullptr;
    double t_wait_gpu = 0, t_rows = 0, t_text = 0, t_wait_stage = 0, t_enqueue = 0;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_call0 = now();
    int64_t delivered = 0;   // frames whose rows / text have been handed over

    // crude per-frame model: host->device copy at ~45 GB/s against the pair kernel at ~0.8 of its 16 pairs/clk/SM ceiling
    const double copy_s = (double)frame_text_max / 45e9;
    const double pair_s = (double)sys.n_sites * sys.site_cap * sys.n_atoms / (0.8 * 16.0 * ctx->sm_count * 1.9e9);
    const bool ramp_down = copy_s > pair_s;
    int64_t planned = 0;                       // frames planned so far
    int64_t n_staged = 0, n_enq = 0, n_drained = 0;   // batches planned (+ staging started) / on the device / delivered
    bool eof = false, stopped = false;
    std::exception_ptr plan_error;   // a malformed later frame: everything planned before it is still delivered

    // plan the next batch into its slot and start its staging copy; false when there is nothing left to plan
    auto plan_and_stage_unguarded = [&]() -> bool {
      if (eof || stopped || plan_error || planned >= n_frames) return false;
      if (ctx->stop_requested.load()) {
        stopped = true;
        return false;
      }
      const int64_t f_begin = first_frame + planned;
      // Batch size ramps up at the start of a call (the first H2D copy is not hidden behind any compute) and down at
      // its end (neither are the last batch's kernels and D2H copy): B/4, B/2, B, ..., B, then halving.
      int want = (int)std::min<int64_t>(B, n_frames - planned);
      {
        const int64_t min_b = std::max<int64_t>(1, B / 4), left = n_frames - planned;
        // (not for .xtc: its stream is a tenth of the text, and every launch of K1x pays the full latency of the index walk)
        if (!ctx->xtc) want = (int)std::min<int64_t>(want, planned + min_b);
        // the ramp down only pays when the copies are the bottleneck (text): a compute-bound stream (.trr) would just
        // finish on small, less efficient launches
        if (ramp_down && left > min_b) want = (int)std::min<int64_t>(want, std::max<int64_t>(min_b, (left + 1) / 2));
      }
      const int64_t known = index_upto(ctx, f_begin + want - 1);
      if (known < f_begin + want) {
        want = (int)std::max<int64_t>(0, known - f_begin);
        eof = true;
        if (want == 0) return false;
      }
      Slot& s = ctx->slot[n_staged % K];
      const FrameSpan& fa = ctx->frames[(size_t)f_begin];
      const FrameSpan& fb = ctx->frames[(size_t)(f_begin + want - 1)];
      const int64_t b0 = fa.atoms & ~(int64_t)15;
      int64_t b1 = ctx->xtc ? fb.end : fb.atoms + (int64_t)sys.n_atoms * sys.line_bytes;
      b1 = std::min(b1, ctx->nbytes);
      const int64_t nb = b1 - b0;
      if (nb > ctx->alloc_text) throw std::runtime_error("internal: batch text larger than the staging buffer");
      for (int i = 0; i < want; i++) s.h_off[i] = ctx->frames[(size_t)(f_begin + i)].atoms - b0;
      s.frames = want;
      s.out_frame = planned;
      s.nb = nb;
      s.staged_here = !direct;
      if (direct) {
        s.src = ctx->text + b0;
      } else {
        s.src = s.h_text;
        pool->submit(&s.stage_job, s.h_text, ctx->text + b0, nb, ctx->mapped);   // page cache / pageable memory -> pinned
      }
      planned += want;
      n_staged++;
      return true;
    };
    auto plan_and_stage = [&]() -> bool {
      try {
        return plan_and_stage_unguarded();
      } catch (const CudaError&) {
        throw;
      } catch (...) {
        plan_error = std::current_exception();
        return false;
      }
    };

    auto enqueue = [&](Slot& s) {
      const int want = s.frames;
      {
        ScopedKernelTimer t(ctx, 4, ctx->st_copy);
        CUDA_OK(cudaMemcpyAsync(s.d_text, s.src, (size_t)s.nb, cudaMemcpyHostToDevice, ctx->st_copy));
      }
      CUDA_OK(cudaMemcpyAsync(s.d_off, s.h_off, sizeof(int64_t) * want, cudaMemcpyHostToDevice, ctx->st_copy));
      if (xtc_overlap) {
        // K1x's index pass needs only the bytes of the frames: it goes on a high-priority stream of its own, so that it
        // neither holds up the next batch's copy nor waits for this batch's turn on the compute stream — the pair kernel
        // of the batch before fills every SM, and the walk (one warp per frame) slips in as that kernel drains and runs
        // on beside the next one
        CUDA_OK(cudaEventRecord(s.ev_text, ctx->st_copy));
        CUDA_OK(cudaStreamWaitEvent(ctx->st_index, s.ev_text, 0));
        launch_xtc_index(sys, s.d_text, s.nb, s.d_off, want, s.d_xtc, ctx->d_err, ctx->st_index);
        ctx->launches++;
        CUDA_OK(cudaEventRecord(s.ev_copied, ctx->st_index));
      } else {
        CUDA_OK(cudaEventRecord(s.ev_copied, ctx->st_copy));
      }
      CUDA_OK(cudaStreamWaitEvent(ctx->st_comp, s.ev_copied, 0));
      enqueue_compute(ctx, s.d_text, s.nb, s.d_off, want, s, s.d_dE, true, ctx->st_comp, xtc_overlap);
      if (want_rows) {
        CUDA_OK(cudaMemcpyAsync(s.h_dE, s.d_dE, sizeof(float) * (size_t)(sg * want), cudaMemcpyDeviceToHost, ctx->st_comp));
      }
      if (sink) {
        const int64_t rows = (int64_t)want * sys.n_sites;
        if (s.text_out_pending) {   // the slot's previous text must have left d_ttext
          CUDA_OK(cudaStreamWaitEvent(ctx->st_comp, s.ev_text_out, 0));
          s.text_out_pending = false;
        }
        CUDA_OK(cudaMemsetAsync(s.d_fmt_flag, 0, sizeof(int), ctx->st_comp));
        launch_format_time(s.d_dE, rows, sys.num_tot, true, s.d_rowoff, s.d_ttext, s.d_fmt_flag, ctx->st_comp);
        ctx->launches += 3;
        CUDA_OK(cudaMemcpyAsync(&s.h_tinfo[0], s.d_rowoff + rows, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->st_comp));
        s.h_tinfo[1] = 0;
        CUDA_OK(cudaMemcpyAsync(&s.h_tinfo[1], s.d_fmt_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->st_comp));
      }
      CUDA_OK(cudaEventRecord(s.ev_done, ctx->st_comp));
      s.busy = true;
    };

    auto drain = [&](Slot& s) {
      if (!s.busy) return;
      double t0 = now();
      CUDA_OK(cudaEventSynchronize(s.ev_done));
      t_wait_gpu += now() - t0;
      t0 = now();
      if (dE_kcal) {
        pool->copy((char*)(dE_kcal + s.out_frame * sg), (const char*)s.h_dE, (int64_t)sizeof(float) * sg * s.frames);
      }
      t_rows += now() - t0;
      t0 = now();
      if (sink) {
        const int64_t total = s.h_tinfo[0];
        if (s.h_tinfo[1] != 0 || total < 0 || total > s.ttext_cap) {
          // a value the device formatter does not cover (non-finite, or outside [1e-16, 1e16)): this batch on the host
          std::string text;
          format_rows(s.h_dE, 0, (int64_t)s.frames * sys.n_sites, sys.num_tot, true, text);
          sink->push_string(std::move(text));
        } else {
          // the kernels of this batch are done (ev_done); later batches keep st_comp busy, so the copy goes on its own
          // stream, and the writer thread — not this one — waits for it
          const int b = sink->acquire();
          CUDA_OK(cudaMemcpyAsync(ctx->sink_buf[b], s.d_ttext, (size_t)total, cudaMemcpyDeviceToHost, ctx->st_out));
          CUDA_OK(cudaEventRecord(ctx->sink_ev[b], ctx->st_out));
          CUDA_OK(cudaEventRecord(s.ev_text_out, ctx->st_out));
          s.text_out_pending = true;
          sink->push_buffer(b, total);
        }
      }
      t_text += now() - t0;
      delivered += s.frames;
      if (frames_done) *frames_done = delivered;
      s.busy = false;
    };

    try {
      for (;;) {
        while (n_staged < n_drained + K && n_staged < n_enq + kStageAhead + 1) {
          if (!plan_and_stage()) break;
        }
        if (n_enq == n_staged) break;   // everything planned is on the device, and there is nothing left to plan
        Slot& s = ctx->slot[n_enq % K];
        double t0 = now();
        if (s.staged_here) pool->wait(&s.stage_job);
        t_wait_stage += now() - t0;
        t0 = now();
        enqueue(s);
        t_enqueue += now() - t0;
        n_enq++;
        while (n_enq - n_drained > kInFlight) {
          drain(ctx->slot[n_drained % K]);
          n_drained++;
        }
      }
      flush_cov_rows(ctx, ctx->st_comp);   // rows still gathered for K3
      while (n_drained < n_enq) {
        drain(ctx->slot[n_drained % K]);
        n_drained++;
      }
      CUDA_OK(cudaStreamSynchronize(ctx->st_comp));
      // A malformed frame met while planning: every batch before it has been delivered by now, so *frames_done, the
      // caller's rows, the `.time` text and the Correlate sums all cover the same frames.
      if (plan_error) std::rethrow_exception(plan_error);
    } catch (...) {
      // Deliver what is already on the device (its rows are in the Correlate sums).  Whatever cannot be delivered (the
      // device itself failed) is dropped by the next call's entry check.
      for (int64_t b = n_enq; b < n_staged; b++) {
        Slot& s = ctx->slot[b % K];
        if (s.staged_here) pool->wait(&s.stage_job);   // no copier may still write into a slot the next call reuses
      }
      try {
        flush_cov_rows(ctx, ctx->st_comp);
        while (n_drained < n_enq) {
          drain(ctx->slot[n_drained % K]);
          n_drained++;
        }
        cudaStreamSynchronize(ctx->st_comp);
      } catch (...) {
      }
      throw;
    }
    if (trace) {
      fprintf(stderr, "cgc trace: %lld frames in %.1f ms (+ %.1f ms buffer set-up: ring %.1f, Correlate %.1f, text %.1f) | host thread: wait for staging %.1f, enqueue %.1f, "
                      "(of which K3 gather copies %.1f, K3 launches %.1f), wait for GPU %.1f, rows to caller %.1f, text D2H + queue %.1f ms | %d frames per batch, %d copier threads\n",
              (long long)delivered, (now() - t_call0) * 1e3, t_setup * 1e3, (t_setup1 - t_setup0) * 1e3, (t_setup2 - t_setup1) * 1e3,
              (t_setup - (t_setup2 - t_setup0)) * 1e3, t_wait_stage * 1e3, t_enqueue * 1e3, ctx->host_cov_gather_s * 1e3,
              ctx->host_cov_flush_s * 1e3, t_wait_gpu * 1e3,
              t_rows * 1e3, t_text * 1e3, B, pool->threads());
    }
    int flags = 0;
    CUDA_OK(cudaMemcpy(&flags, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost));
    if (flags) {
      CUDA_OK(zero_now(ctx->d_err, sizeof(int)));
      std::string m = "device decode found malformed coordinate text:";
      if (flags & kDecBadChar) m += " unexpected character in a coordinate field;";
      if (flags & kDecBadPoint) m += " decimal point not at the column found in frame 0;";
      if (flags & kDecBadNewline) m += " atom line of a different length than in frame 0;";
      if (flags & kDecBadXtc) m = "device decode found an impossible .xtc frame (atom count, header fields, or a bit stream whose runs do not add up)";
      throw InputFileMalformed(m);
    }
    if (stopped) {
      ctx->stop_requested.store(false);
      return CGC_STOPPED;
    }
    return delivered < n_frames ? CGC_EOF : CGC_OK;
  });
}

}  // namespace

int cgc_prepare(cgc_context* ctx, int with_time_text) {
  if (!ctx) return CGC_ERR_ARG;
  if (!ctx->opened) return fail(ctx, CGC_ERR_STATE, "no trajectory opened");
  return guarded(ctx, [&]() {
    need_device(ctx);
    const DeviceSystem& sys = ctx->sys;
    const int B = ctx->batch_frames > 0 ? ctx->batch_frames : default_batch(ctx);
    const int64_t frame_text_max = std::max<int64_t>((int64_t)sys.n_atoms * sys.line_bytes, ctx->frame_span_max) + 4096;
    bool direct = false;
    if (!ctx->mapped) {
      cudaPointerAttributes at;
      direct = cudaPointerGetAttributes(&at, ctx->text) == cudaSuccess && at.type == cudaMemoryTypeHost;
      cudaGetLastError();
    }
    ensure_pipeline(ctx, B, frame_text_max * B, !direct, ctx->ring_slots);
    ensure_cov(ctx);
    ensure_cov_buffers(ctx, B);
    if (with_time_text) ensure_time_text(ctx, B);
    copy_pool(ctx);
    return CGC_OK;
  });
}

void cgc_request_stop(cgc_context* ctx) {
  if (ctx) ctx->stop_requested.store(true);   // one atomic store: safe to call from a signal handler
}

int cgc_run(cgc_context* ctx, int64_t first_frame, int64_t n_frames, float* dE_kcal, int64_t* frames_done) {
  return run_impl(ctx, first_frame, n_frames, dE_kcal, frames_done, nullptr);
}

int cgc_run_write_time(cgc_context* ctx, int64_t first_frame, int64_t n_frames, const char* time_path, int truncate,
                       float* dE_kcal, int64_t* frames_done) {
  if (frames_done) *frames_done = 0;
  if (!ctx || !time_path) return fail(ctx, CGC_ERR_ARG, "null argument");
  int fd = open(time_path, O_WRONLY | O_CREAT | (truncate ? O_TRUNC : 0), 0644);   // the sink appends by offset (pwrite)
  if (fd < 0) return fail(ctx, CGC_ERR_READ, std::string("cannot open ") + time_path + " for writing");
  int rc;
  bool wrote;
  {
    TextSink sink(fd, ctx);
    rc = run_impl(ctx, first_frame, n_frames, dE_kcal, frames_done, &sink);
    wrote = sink.finish();
  }
  if (close(fd) != 0) wrote = false;
  if (!wrote && (rc == CGC_OK || rc == CGC_EOF || rc == CGC_STOPPED)) return fail(ctx, CGC_ERR_READ, std::string("write to ") + time_path + " failed");
  return rc;
}

int cgc_format_text_device(cgc_context* ctx, const float* values, int64_t rows, int num_col, int to_cm, char* out,
                           int64_t cap, int64_t* nbytes, int* needs_host) {
  if (nbytes) *nbytes = 0;
  if (needs_host) *needs_host = 0;
  if (!ctx || !values || !out || !nbytes || rows < 0 || num_col < 1) return fail(ctx, CGC_ERR_ARG, "bad argument");
  return guarded(ctx, [&]() {
    need_device(ctx);
    if (rows == 0) return CGC_OK;
    if (!ctx->st_comp) CUDA_OK(cudaStreamCreateWithFlags(&ctx->st_comp, cudaStreamNonBlocking));
    float* d_x = nullptr;
    char* d_o = nullptr;
    int64_t* d_off = nullptr;
    int* d_fl = nullptr;
    cudaError_t err = cudaSuccess;
    auto ok = [&](cudaError_t e) { if (err == cudaSuccess) err = e; return e == cudaSuccess; };
    const int64_t n = rows * num_col, capd = format_time_capacity(rows, num_col);
    int64_t total = 0;
    int fl = 0;
    if (ok(cudaMalloc(&d_x, sizeof(float) * n)) && ok(cudaMalloc(&d_o, (size_t)capd)) &&
        ok(cudaMalloc(&d_off, sizeof(int64_t) * (rows + 1))) && ok(cudaMalloc(&d_fl, sizeof(int))) &&
        ok(cudaMemsetAsync(d_fl, 0, sizeof(int), ctx->st_comp)) &&
        ok(cudaMemcpyAsync(d_x, values, sizeof(float) * n, cudaMemcpyHostToDevice, ctx->st_comp))) {
      launch_format_time(d_x, rows, num_col, to_cm != 0, d_off, d_o, d_fl, ctx->st_comp);
      ctx->launches += 3;
      ok(cudaGetLastError());
      ok(cudaMemcpyAsync(&total, d_off + rows, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->st_comp));
      ok(cudaMemcpyAsync(&fl, d_fl, sizeof(int), cudaMemcpyDeviceToHost, ctx->st_comp));
      ok(cudaStreamSynchronize(ctx->st_comp));
      if (err == cudaSuccess && total <= cap) ok(cudaMemcpy(out, d_o, (size_t)total, cudaMemcpyDeviceToHost));
    }
    cudaFree(d_x); cudaFree(d_o); cudaFree(d_off); cudaFree(d_fl);
    CUDA_OK(err);
    *nbytes = total;
    if (needs_host) *needs_host = fl;
    if (total > cap) throw std::invalid_argument("output buffer too small for the formatted text");
    return CGC_OK;
  });
}

int cgc_run_device(cgc_context* ctx, const void* d_text, int64_t text_bytes, const int64_t* atoms_offset, int n_frames,
                   float* d_dE_kcal, void* stream) {
  if (!ctx) return CGC_ERR_ARG;
  if (!ctx->opened) return fail(ctx, CGC_ERR_STATE, "no trajectory opened");
  if (!d_text || !atoms_offset || !d_dE_kcal || n_frames < 1) return fail(ctx, CGC_ERR_ARG, "bad argument");
  if (((uintptr_t)d_text & 15) != 0) return fail(ctx, CGC_ERR_ARG, "d_text must be 16-byte aligned");
  return guarded(ctx, [&]() {
    need_device(ctx);
    const int B = std::max(n_frames, ctx->batch_frames > 0 ? ctx->batch_frames : default_batch(ctx));
    ensure_pipeline(ctx, B, std::max<int64_t>(ctx->alloc_text, 64), false, std::max(1, ctx->alloc_slots));
    Slot& s = ctx->slot[0];
    cudaStream_t st = (cudaStream_t)stream;
    // pageable source: the runtime stages it before returning, so the caller's array may be reused at once
    CUDA_OK(cudaMemcpyAsync(s.d_off, atoms_offset, sizeof(int64_t) * n_frames, cudaMemcpyHostToDevice, st));
    enqueue_compute(ctx, (const char*)d_text, text_bytes, s.d_off, n_frames, s, d_dE_kcal, false, st);
    return CGC_OK;
  });
}

int cgc_decode_device(cgc_context* ctx, const void* d_text, int64_t text_bytes, const int64_t* atoms_offset,
                      int n_frames, float* d_xyz, void* stream) {
  if (!ctx) return CGC_ERR_ARG;
  if (!ctx->opened) return fail(ctx, CGC_ERR_STATE, "no trajectory opened");
  if (!d_text || !atoms_offset || !d_xyz || n_frames < 1) return fail(ctx, CGC_ERR_ARG, "bad argument");
  if (((uintptr_t)d_text & 15) != 0) return fail(ctx, CGC_ERR_ARG, "d_text must be 16-byte aligned");
  return guarded(ctx, [&]() {
    need_device(ctx);
    const int B = std::max(n_frames, ctx->batch_frames > 0 ? ctx->batch_frames : default_batch(ctx));
    ensure_pipeline(ctx, B, std::max<int64_t>(ctx->alloc_text, 64), false, std::max(1, ctx->alloc_slots));
    Slot& s = ctx->slot[0];
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_OK(cudaMemcpyAsync(s.d_off, atoms_offset, sizeof(int64_t) * n_frames, cudaMemcpyHostToDevice, st));
    enqueue_decode(ctx, (const char*)d_text, text_bytes, s.d_off, n_frames, s, st);
    launch_unblock(ctx->sys, s.d_xyz8, d_xyz, n_frames, st);
    ctx->launches += 2;
    CUDA_OK(cudaGetLastError());
    return CGC_OK;
  });
}

// ---------------------------------------------------------------------------------------------------------------------
int64_t cgc_cov_size(const cgc_context* ctx) {
  if (!ctx || !ctx->opened || ctx->device < 0 || !cov_enabled(ctx)) return 0;
  return cov_doubles(ctx);
}

int cgc_cov_bind(cgc_context* ctx, double* d_buffer) {
  if (!ctx || !d_buffer) return fail(ctx, CGC_ERR_ARG, "null argument");
  if (!ctx->opened) return fail(ctx, CGC_ERR_STATE, "no trajectory opened");
  return guarded(ctx, [&]() {
    need_device(ctx);
    if (!cov_enabled(ctx)) throw std::invalid_argument("covariance is switched off for this trajectory");
    if (ctx->d_cov && !ctx->cov_bound) CUDA_OK(guard_free(ctx->d_cov));
    ctx->d_cov = d_buffer;
    ctx->cov_bound = true;
    ctx->shift_set = false;
    ctx->cov_rows_pending = 0;
    CUDA_OK(zero_now(ctx->d_cov, sizeof(double) * cov_doubles(ctx)));
    ensure_cov(ctx);
    const int64_t SG = (int64_t)ctx->sys.n_sites * ctx->sys.num_tot;
    CUDA_OK(zero_now(ctx->d_compat, sizeof(float) * 2 * SG));
    return CGC_OK;
  });
}

int cgc_cov_reset(cgc_context* ctx) {
  if (!ctx || !ctx->opened) return fail(ctx, CGC_ERR_STATE, "no trajectory opened");
  return guarded(ctx, [&]() {
    need_device(ctx);
    ensure_cov(ctx);
    if (!ctx->d_cov) throw std::invalid_argument("covariance is switched off for this trajectory");
    const int64_t SG = (int64_t)ctx->sys.n_sites * ctx->sys.num_tot;
    CUDA_OK(zero_now(ctx->d_cov, sizeof(double) * cov_doubles(ctx)));
    CUDA_OK(zero_now(ctx->d_compat, sizeof(float) * 2 * SG));
    ctx->shift_set = false;
    ctx->cov_rows_pending = 0;
    return CGC_OK;
  });
}

int cgc_cov_set_shift_device(cgc_context* ctx, const float* d_row, void* stream) {
  if (!ctx || !d_row) return fail(ctx, CGC_ERR_ARG, "null argument");
  if (!ctx->opened) return fail(ctx, CGC_ERR_STATE, "no trajectory opened");
  return guarded(ctx, [&]() {
    need_device(ctx);
    ensure_cov(ctx);
    if (!ctx->d_cov) throw std::invalid_argument("covariance is switched off for this trajectory");
    launch_cov_set_shift(d_row, ctx->d_cov, ctx->sys.n_sites, ctx->sys.num_tot, (cudaStream_t)stream);
    ctx->launches++;
    ctx->shift_set = true;
    return CGC_OK;
  });
}

int cgc_cov_accumulate_device(cgc_context* ctx, const float* d_dE_kcal, int n_frames, void* stream) {
  if (!ctx || !d_dE_kcal || n_frames < 1) return fail(ctx, CGC_ERR_ARG, "bad argument");
  if (!ctx->opened) return fail(ctx, CGC_ERR_STATE, "no trajectory opened");
  return guarded(ctx, [&]() {
    need_device(ctx);
    ensure_cov(ctx);
    if (!ctx->d_cov) throw std::invalid_argument("covariance is switched off for this trajectory");
    cudaStream_t st = (cudaStream_t)stream;
    if (!ctx->shift_set) {
      launch_cov_set_shift(d_dE_kcal, ctx->d_cov, ctx->sys.n_sites, ctx->sys.num_tot, st);
      ctx->shift_set = true;
      ctx->launches++;
    }
    {
      ScopedKernelTimer t(ctx, 3, st);
      CUDA_OK(launch_cov_accumulate(d_dE_kcal, n_frames, ctx->d_cov, ctx->compat_on ? ctx->d_compat : nullptr, centred_scratch(ctx, n_frames),
                                    ctx->sys.n_sites, ctx->sys.num_tot, st));
    }
    ctx->launches += 3;
    CUDA_OK(cudaGetLastError());
    return CGC_OK;
  });
}

int cgc_cov_partials(cgc_context* ctx, double* host_out) {
  if (!ctx || !host_out) return fail(ctx, CGC_ERR_ARG, "null argument");
  return guarded(ctx, [&]() {
    need_device(ctx);
    if (!ctx->d_cov) throw std::invalid_argument("no covariance sums on this context");
    CUDA_OK(cudaDeviceSynchronize());
    CUDA_OK(cudaMemcpy(host_out, ctx->d_cov, sizeof(double) * cov_doubles(ctx), cudaMemcpyDeviceToHost));
    return CGC_OK;
  });
}

// ---------------------------------------------------------------------------------------------------------------------
// The one collective of the hot path, for several contexts living in ONE process (one host thread and one context per GPU,
// the layout SURVEY.md section 5 names): ncclCommInitAll + ncclAllReduce(sum, double) over the partial sums.  NCCL is
// bound at run time (dlopen), so the library carries no link-time dependency on it and a process that already holds a
// libnccl.so.2 (torch) shares that one.
namespace {
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string problem;
};
NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (!api.lib) {
      api.problem = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "?");
      return;
    }
    auto sym = [&](const char* n) {
      void* p = dlsym(api.lib, n);
      if (!p && api.problem.empty()) api.problem = std::string("libnccl has no symbol ") + n;
      return p;
    };
    api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  });
  return api;
}
}  // namespace

struct cgc_comm {
  std::vector<int> devices;
  std::vector<ncclComm_t> comms;
};

int cgc_comm_create(const int* devices, int n, cgc_comm** out) {
  if (!devices || n < 1 || !out) return fail(nullptr, CGC_ERR_ARG, "null or empty device list");
  *out = nullptr;
  return guarded(nullptr, [&]() {
    NcclApi& api = nccl_api();
    if (!api.problem.empty()) throw Unsupported(api.problem);
    cgc_comm* c = new cgc_comm();
    c->devices.assign(devices, devices + n);
    c->comms.assign((size_t)n, nullptr);
    const ncclResult_t r = api.CommInitAll(c->comms.data(), n, c->devices.data());
    if (r != ncclSuccess) {
      delete c;
      throw CudaError(std::string("ncclCommInitAll: ") + (api.GetErrorString ? api.GetErrorString(r) : "NCCL error"));
    }
    *out = c;
    return CGC_OK;
  });
}

void cgc_comm_destroy(cgc_comm* comm) {
  if (!comm) return;
  NcclApi& api = nccl_api();
  for (auto cm : comm->comms)
    if (cm && api.CommDestroy) api.CommDestroy(cm);
  delete comm;
}

int cgc_cov_allreduce(cgc_context** ctxs, int n, cgc_comm* comm) {
  if (!ctxs || n < 1) return CGC_ERR_ARG;
  cgc_context* c0 = ctxs[0];
  if (!c0) return CGC_ERR_ARG;
  return guarded(c0, [&]() {
    for (int i = 0; i < n; i++) {
      cgc_context* c = ctxs[i];
      if (!c || !c->opened || c->device < 0) throw std::invalid_argument("allreduce: every context needs an open trajectory and a device");
      if (!c->d_cov) throw std::invalid_argument("allreduce: covariance is switched off on one of the contexts");
      if (c->sys.n_sites != c0->sys.n_sites || c->sys.num_tot != c0->sys.num_tot) {
        throw std::invalid_argument("allreduce: the contexts do not describe the same system (sites / columns differ)");
      }
      for (int j = 0; j < i; j++)
        if (ctxs[j]->device == c->device) throw std::invalid_argument("allreduce: two contexts share a device");
      if (comm && ((int)comm->devices.size() != n || comm->devices[(size_t)i] != c->device)) {
        throw std::invalid_argument("allreduce: the communicator was created for another device list");
      }
    }
    if (n == 1) return CGC_OK;
    NcclApi& api = nccl_api();
    if (!api.problem.empty()) throw Unsupported(api.problem);
    auto NCCL_OK = [&](ncclResult_t r, const char* what) {
      if (r != ncclSuccess) throw CudaError(std::string(what) + ": " + (api.GetErrorString ? api.GetErrorString(r) : "NCCL error"));
    };
    const int64_t SG = (int64_t)c0->sys.n_sites * c0->sys.num_tot;
    std::vector<int> devs((size_t)n);
    for (int i = 0; i < n; i++) {
      devs[(size_t)i] = ctxs[i]->device;
      CUDA_OK(cudaSetDevice(ctxs[i]->device));
      flush_cov_rows(ctxs[i], ctxs[i]->st_comp);
      CUDA_OK(cudaStreamSynchronize(ctxs[i]->st_comp));
    }
    cgc_comm* own = nullptr;
    if (!comm) {
      own = new cgc_comm();
      own->devices = devs;
      own->comms.assign((size_t)n, nullptr);
      const ncclResult_t r = api.CommInitAll(own->comms.data(), n, devs.data());
      if (r != ncclSuccess) {
        delete own;
        NCCL_OK(r, "ncclCommInitAll");
      }
      comm = own;
    }
    bool failed = false;
    std::string msg;
    try {
      // [ n | shift c | sum(x-c) | sum (x-c)(x-c)^T ]: the shift block is identical on every rank and stays as it is;
      // n, and the two sum blocks (contiguous), are added — as one grouped operation
      NCCL_OK(api.GroupStart(), "ncclGroupStart");
      for (int i = 0; i < n; i++) {
        double* b = ctxs[i]->d_cov;
        NCCL_OK(api.AllReduce(b, b, 1, ncclDouble, ncclSum, comm->comms[(size_t)i], ctxs[i]->st_comp), "ncclAllReduce");
        NCCL_OK(api.AllReduce(b + 1 + SG, b + 1 + SG, (size_t)(SG + SG * c0->sys.num_tot), ncclDouble, ncclSum, comm->comms[(size_t)i],
                              ctxs[i]->st_comp), "ncclAllReduce");
      }
      NCCL_OK(api.GroupEnd(), "ncclGroupEnd");
      for (int i = 0; i < n; i++) {
        CUDA_OK(cudaSetDevice(ctxs[i]->device));
        CUDA_OK(cudaStreamSynchronize(ctxs[i]->st_comp));
      }
    } catch (const std::exception& e) {
      failed = true;
      msg = e.what();
    }
    if (own) cgc_comm_destroy(own);
    if (failed) throw CudaError(msg);
    return CGC_OK;
  });
}

// Checkpoint of the Correlate state: header, the flat partial-sum buffer, the two float32 compat sums.
namespace {
struct CovStateHeader {
  char magic[8];        // "CGCCOV1"
  int32_t sites, num_tot;
  int64_t doubles;      // length of the flat buffer
  int32_t shift_set, reserved;
};
}  // namespace

int64_t cgc_cov_state_bytes(const cgc_context* ctx) {
  if (!ctx || !ctx->opened || ctx->device < 0 || !cov_enabled(ctx)) return 0;
  const int64_t SG = (int64_t)ctx->sys.n_sites * ctx->sys.num_tot;
  return (int64_t)sizeof(CovStateHeader) + (int64_t)sizeof(double) * cov_doubles(ctx) + (int64_t)sizeof(float) * 2 * SG;
}

int cgc_cov_save_state(cgc_context* ctx, void* host_out, int64_t cap) {
  if (!ctx || !host_out) return fail(ctx, CGC_ERR_ARG, "null argument");
  if (!ctx->opened) return fail(ctx, CGC_ERR_STATE, "no trajectory opened");
  return guarded(ctx, [&]() {
    need_device(ctx);
    ensure_cov(ctx);
    if (!ctx->d_cov) throw std::invalid_argument("covariance is switched off for this trajectory");
    if (cap < cgc_cov_state_bytes(ctx)) throw std::invalid_argument("buffer smaller than cgc_cov_state_bytes()");
    const int64_t SG = (int64_t)ctx->sys.n_sites * ctx->sys.num_tot, nd = cov_doubles(ctx);
    CovStateHeader h{};
    memcpy(h.magic, "CGCCOV1", 8);
    h.sites = ctx->sys.n_sites;
    h.num_tot = ctx->sys.num_tot;
    h.doubles = nd;
    h.shift_set = ctx->shift_set ? 1 : 0;
    char* p = (char*)host_out;
    memcpy(p, &h, sizeof h);
    CUDA_OK(cudaDeviceSynchronize());
    CUDA_OK(cudaMemcpy(p + sizeof h, ctx->d_cov, sizeof(double) * nd, cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemcpy(p + sizeof h + sizeof(double) * nd, ctx->d_compat, sizeof(float) * 2 * SG, cudaMemcpyDeviceToHost));
    return CGC_OK;
  });
}

int cgc_cov_load_state(cgc_context* ctx, const void* host_in, int64_t nbytes) {
  if (!ctx || !host_in) return fail(ctx, CGC_ERR_ARG, "null argument");
  if (!ctx->opened) return fail(ctx, CGC_ERR_STATE, "no trajectory opened");
  return guarded(ctx, [&]() {
    need_device(ctx);
    ensure_cov(ctx);
    if (!ctx->d_cov) throw std::invalid_argument("covariance is switched off for this trajectory");
    const int64_t SG = (int64_t)ctx->sys.n_sites * ctx->sys.num_tot, nd = cov_doubles(ctx);
    CovStateHeader h;
    if (nbytes < (int64_t)sizeof h) throw InputFileMalformed("Correlate checkpoint is truncated");
    memcpy(&h, host_in, sizeof h);
    if (memcmp(h.magic, "CGCCOV1", 8) != 0) throw InputFileMalformed("not a Correlate checkpoint");
    if (h.sites != ctx->sys.n_sites || h.num_tot != ctx->sys.num_tot || h.doubles != nd) {
      throw InputFileMalformed("Correlate checkpoint belongs to a different system (sites / columns differ)");
    }
    if (nbytes < cgc_cov_state_bytes(ctx)) throw InputFileMalformed("Correlate checkpoint is truncated");
    const char* p = (const char*)host_in + sizeof h;
    CUDA_OK(cudaDeviceSynchronize());
    CUDA_OK(cudaMemcpy(ctx->d_cov, p, sizeof(double) * nd, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(ctx->d_compat, p + sizeof(double) * nd, sizeof(float) * 2 * SG, cudaMemcpyHostToDevice));
    ctx->shift_set = h.shift_set != 0;
    ctx->cov_rows_pending = 0;
    return CGC_OK;
  });
}

int cgc_cov_finalize(cgc_context* ctx, int ddof, double* mean, double* cov) {
  if (!ctx || !mean || !cov) return fail(ctx, CGC_ERR_ARG, "null argument");
  return guarded(ctx, [&]() {
    need_device(ctx);
    if (!ctx->d_cov) throw std::invalid_argument("no covariance sums on this context");
    const int S = ctx->sys.n_sites, G = ctx->sys.num_tot;
    double *d_mean = nullptr, *d_covm = nullptr;
    CUDA_OK(cudaDeviceSynchronize());
    CUDA_OK(cudaMalloc(&d_mean, sizeof(double) * S * G));
    CUDA_OK(cudaMalloc(&d_covm, sizeof(double) * (size_t)S * G * G));
    launch_cov_finalize(ctx->d_cov, S, G, ddof, d_mean, d_covm, ctx->st_comp);
    ctx->launches++;
    cudaError_t e1 = cudaMemcpyAsync(mean, d_mean, sizeof(double) * S * G, cudaMemcpyDeviceToHost, ctx->st_comp);
    cudaError_t e2 = cudaMemcpyAsync(cov, d_covm, sizeof(double) * (size_t)S * G * G, cudaMemcpyDeviceToHost, ctx->st_comp);
    cudaError_t e3 = cudaStreamSynchronize(ctx->st_comp);
    cudaFree(d_mean);
    cudaFree(d_covm);
    CUDA_OK(e1);
    CUDA_OK(e2);
    CUDA_OK(e3);
    return CGC_OK;
  });
}

int cgc_eigh(cgc_context* ctx, const double* matrices, int n_mat, int n, double* evals, double* evecs, int* sweeps) {
  if (sweeps) *sweeps = 0;
  if (!ctx || !matrices || !evals || !evecs) return fail(ctx, CGC_ERR_ARG, "null argument");
  if (n_mat < 1 || n < 1) return fail(ctx, CGC_ERR_ARG, "bad matrix count or size");
  return guarded(ctx, [&]() {
    need_device(ctx);
    if (!ctx->st_comp) CUDA_OK(cudaStreamCreateWithFlags(&ctx->st_comp, cudaStreamNonBlocking));
    const size_t nn = (size_t)n_mat * n * n;
    double *d_a = nullptr, *d_w = nullptr, *d_v = nullptr, *d_e = nullptr;
    unsigned int* d_rot = nullptr;
    cudaError_t err = cudaSuccess;
    auto ok = [&](cudaError_t e) { if (err == cudaSuccess) err = e; return e == cudaSuccess; };
    int sw = -2;
    std::vector<double> ev((size_t)n_mat * n), vt;
    if (ok(cudaMalloc(&d_a, sizeof(double) * nn)) && ok(cudaMalloc(&d_w, sizeof(double) * nn)) &&
        ok(cudaMalloc(&d_v, sizeof(double) * nn)) && ok(cudaMalloc(&d_e, sizeof(double) * n_mat * n)) &&
        ok(cudaMalloc(&d_rot, sizeof(unsigned int))) &&
        ok(cudaMemcpyAsync(d_a, matrices, sizeof(double) * nn, cudaMemcpyHostToDevice, ctx->st_comp))) {
      sw = eigh_jacobi(d_a, n_mat, n, d_w, d_v, d_e, d_rot, 40, ctx->st_comp, &ctx->launches);
      ok(cudaGetLastError());
      vt.resize(nn);
      ok(cudaMemcpyAsync(ev.data(), d_e, sizeof(double) * n_mat * n, cudaMemcpyDeviceToHost, ctx->st_comp));
      ok(cudaMemcpyAsync(vt.data(), d_v, sizeof(double) * nn, cudaMemcpyDeviceToHost, ctx->st_comp));
      ok(cudaStreamSynchronize(ctx->st_comp));
    }
    cudaFree(d_a); cudaFree(d_w); cudaFree(d_v); cudaFree(d_e); cudaFree(d_rot);
    CUDA_OK(err);
    if (sw == -2) throw CudaError("eigh: the device reported an error during the Jacobi sweeps");
    if (sw < 0) throw std::runtime_error("eigh: Jacobi did not converge in 40 sweeps");
    if (sweeps) *sweeps = sw;
    // ascending eigenvalues, eigenvectors as rows in the same order (numpy.linalg.eigh's order; its columns are our rows)
    std::vector<int> order((size_t)n);
    for (int k = 0; k < n_mat; k++) {
      const double* e = ev.data() + (size_t)k * n;
      for (int i = 0; i < n; i++) order[(size_t)i] = i;
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return e[a] < e[b]; });
      for (int i = 0; i < n; i++) {
        evals[(size_t)k * n + i] = e[order[(size_t)i]];
        memcpy(evecs + ((size_t)k * n + i) * n, vt.data() + ((size_t)k * n + order[(size_t)i]) * n, sizeof(double) * n);
      }
    }
    return CGC_OK;
  });
}

int cgc_project(cgc_context* ctx, const float* dE_rows, int64_t n_frames, const double* mean, const double* modes,
                int n_modes, double* out) {
  if (!ctx || !dE_rows || !mean || !modes || !out) return fail(ctx, CGC_ERR_ARG, "null argument");
  if (!ctx->opened) return fail(ctx, CGC_ERR_STATE, "no trajectory opened");
  if (n_frames < 0 || n_modes < 1) return fail(ctx, CGC_ERR_ARG, "bad frame or mode count");
  return guarded(ctx, [&]() {
    need_device(ctx);
    const int S = ctx->sys.n_sites, G = ctx->sys.num_tot, M = n_modes;
    if (n_frames == 0) return CGC_OK;
    const int64_t sg = (int64_t)S * G, sm = (int64_t)S * M;
    // frames per pass: rows (4 B) + centred rows (8 B) + projections (8 B) within about 1 GiB
    const int64_t chunk = std::max<int64_t>(64, std::min<int64_t>(n_frames, ((int64_t)1 << 30) / (sg * 12 + sm * 8)));
    float* d_x = nullptr;
    double *d_c = nullptr, *d_o = nullptr, *d_mean = nullptr, *d_modes = nullptr;
    cudaError_t err = cudaSuccess;
    auto ok = [&](cudaError_t e) { if (err == cudaSuccess) err = e; return e == cudaSuccess; };
    CUDA_OK(cudaDeviceSynchronize());
    if (ok(cudaMalloc(&d_x, sizeof(float) * chunk * sg)) && ok(cudaMalloc(&d_c, sizeof(double) * chunk * sg)) &&
        ok(cudaMalloc(&d_o, sizeof(double) * chunk * sm)) && ok(cudaMalloc(&d_mean, sizeof(double) * sg)) &&
        ok(cudaMalloc(&d_modes, sizeof(double) * sm * G)) &&
        ok(cudaMemcpy(d_mean, mean, sizeof(double) * sg, cudaMemcpyHostToDevice)) &&
        ok(cudaMemcpy(d_modes, modes, sizeof(double) * sm * G, cudaMemcpyHostToDevice))) {
      for (int64_t t = 0; t < n_frames && err == cudaSuccess; t += chunk) {
        const int n = (int)std::min<int64_t>(chunk, n_frames - t);
        if (!ok(cudaMemcpyAsync(d_x, dE_rows + t * sg, sizeof(float) * n * sg, cudaMemcpyHostToDevice, ctx->st_comp))) break;
        launch_project(d_x, n, d_mean, d_modes, M, d_c, d_o, S, G, ctx->st_comp);
        ctx->launches += 2;
        ok(cudaGetLastError());
        ok(cudaMemcpyAsync(out + t * sm, d_o, sizeof(double) * n * sm, cudaMemcpyDeviceToHost, ctx->st_comp));
        ok(cudaStreamSynchronize(ctx->st_comp));
      }
    }
    cudaFree(d_x); cudaFree(d_c); cudaFree(d_o); cudaFree(d_mean); cudaFree(d_modes);
    CUDA_OK(err);
    return CGC_OK;
  });
}

int cgc_time_corr(cgc_context* ctx, const double* series, int n_series, int64_t n, const int32_t* pair_first,
                  const int32_t* pair_second, int n_pairs, int64_t corr_len, int circular, double* out, double* kernel_ms) {
  if (kernel_ms) *kernel_ms = 0.0;
  if (!ctx || !series || !pair_first || !pair_second || !out) return fail(ctx, CGC_ERR_ARG, "null argument");
  if (n_series < 1 || n < 1 || n_pairs < 0) return fail(ctx, CGC_ERR_ARG, "bad series or pair count");
  if (corr_len < 1 || corr_len > n) return fail(ctx, CGC_ERR_ARG, "corr_len must be in [1, n]: lag n has no samples");
  for (int p = 0; p < n_pairs; p++)
    if (pair_first[p] < 0 || pair_first[p] >= n_series || pair_second[p] < 0 || pair_second[p] >= n_series)
      return fail(ctx, CGC_ERR_ARG, "pair index outside the series");
  if (n_pairs == 0) return CGC_OK;
  return guarded(ctx, [&]() {
    need_device(ctx);
    if (!ctx->st_comp) CUDA_OK(cudaStreamCreateWithFlags(&ctx->st_comp, cudaStreamNonBlocking));
    const int64_t stride = time_corr_series_stride(n);
    // pairs per pass: the grid's z extent, and about 1 GiB of partial sums (CGC_TIMECORR_SCRATCH_DOUBLES: test hook)
    int64_t scratch_cap = (int64_t)1 << 27;
    if (const char* e = getenv("CGC_TIMECORR_SCRATCH_DOUBLES")) scratch_cap = std::max<int64_t>(1, atoll(e));
    int batch = std::min(n_pairs, 65535);
    int splits = 0, wsplits = 0;
    int64_t scratch = 0;
    for (;;) {
      time_corr_plan(n, corr_len, batch, circular != 0, ctx->sm_count, &splits, &wsplits, &scratch);
      if (batch == 1 || scratch <= scratch_cap) break;
      batch = std::max(1, batch / 2);
    }
    double *d_s = nullptr, *d_part = nullptr, *d_o = nullptr;
    int32_t* d_pairs = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t err = cudaSuccess;
    auto ok = [&](cudaError_t e) { if (err == cudaSuccess) err = e; return e == cudaSuccess; };
    float ms_total = 0.f;
    if (ok(cudaMalloc(&d_s, sizeof(double) * n_series * stride)) && ok(cudaMalloc(&d_part, sizeof(double) * scratch)) &&
        ok(cudaMalloc(&d_o, sizeof(double) * batch * corr_len)) && ok(cudaMalloc(&d_pairs, sizeof(int32_t) * 2 * batch)) &&
        ok(cudaEventCreate(&e0)) && ok(cudaEventCreate(&e1)) &&
        ok(cudaMemsetAsync(d_s, 0, sizeof(double) * n_series * stride, ctx->st_comp)) &&
        ok(cudaMemcpy2DAsync(d_s, sizeof(double) * stride, series, sizeof(double) * n, sizeof(double) * n, (size_t)n_series,
                             cudaMemcpyHostToDevice, ctx->st_comp))) {
      for (int p0 = 0; p0 < n_pairs && err == cudaSuccess; p0 += batch) {
        const int np = std::min(batch, n_pairs - p0);
        if (!ok(cudaMemcpyAsync(d_pairs, pair_first + p0, sizeof(int32_t) * np, cudaMemcpyHostToDevice, ctx->st_comp)) ||
            !ok(cudaMemcpyAsync(d_pairs + batch, pair_second + p0, sizeof(int32_t) * np, cudaMemcpyHostToDevice, ctx->st_comp)))
          break;
        ok(cudaEventRecord(e0, ctx->st_comp));
        ctx->launches += launch_time_corr(d_s, n, d_pairs, d_pairs + batch, np, corr_len, circular != 0, ctx->sm_count, d_part,
                                          d_o, ctx->st_comp);
        ok(cudaGetLastError());
        ok(cudaEventRecord(e1, ctx->st_comp));
        ok(cudaMemcpyAsync(out + (int64_t)p0 * corr_len, d_o, sizeof(double) * np * corr_len, cudaMemcpyDeviceToHost,
                           ctx->st_comp));
        ok(cudaStreamSynchronize(ctx->st_comp));
        float ms = 0.f;
        if (err == cudaSuccess && ok(cudaEventElapsedTime(&ms, e0, e1))) ms_total += ms;
      }
    }
    cudaFree(d_s); cudaFree(d_part); cudaFree(d_o); cudaFree(d_pairs);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    CUDA_OK(err);
    if (kernel_ms) *kernel_ms = ms_total;
    return CGC_OK;
  });
}

// ---------------------------------------------------------------------------------------------------------------------
// K8: `.time` / `.mtx` text -> float64
// ---------------------------------------------------------------------------------------------------------------------
namespace {

// strtod of every token of lines [text + b, text + e) (each line ends with '\n' or at e): the path for chunks that hold
// tokens the device does not take (nan, inf, long mantissas, huge exponents) — and the one that names a bad token.
// Returns the number of rows, or throws InputFileMalformed.
int64_t parse_rows_host(const char* text, int64_t b, int64_t e, int num_col, double* out, int64_t cap_rows, int64_t row0) {
  std::vector<int64_t> starts;
  for (int64_t p = b; p < e;) {
    starts.push_back(p);
    const void* nl = memchr(text + p, '\n', (size_t)(e - p));
    p = nl ? (const char*)nl - text + 1 : e;
  }
  const int64_t rows = (int64_t)starts.size();
  if (rows > cap_rows) throw std::runtime_error("parse_text: output buffer too small");
  starts.push_back(e);
  const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(16u, std::max(1u, std::thread::hardware_concurrency())), rows / 64));
  std::vector<std::string> errs((size_t)nt);
  std::vector<int64_t> err_row((size_t)nt, INT64_MAX);
  auto work = [&](int t) {
    const int64_t r0 = rows * t / nt, r1 = rows * (t + 1) / nt;
    std::string tok;
    for (int64_t r = r0; r < r1; r++) {
      int64_t p = starts[(size_t)r], end = starts[(size_t)r + 1];
      if (end > p && text[end - 1] == '\n') end--;
      int col = 0;
      while (p <= end) {
        const void* cm = p < end ? memchr(text + p, ',', (size_t)(end - p)) : nullptr;
        const int64_t q = cm ? (const char*)cm - text : end;
        tok.assign(text + p, (size_t)(q - p));
        char* stop = nullptr;
        const double v = strtod(tok.c_str(), &stop);
        while (stop && (*stop == ' ' || *stop == '\t' || *stop == '\r')) stop++;
        // Python's float() takes neither hexadecimal floats nor nan(payload), both of which strtod would accept
        const bool foreign = tok.find_first_of("xXpP(") != std::string::npos;
        if (stop == tok.c_str() || !stop || *stop != 0 || foreign) {
          err_row[(size_t)t] = r;
          errs[(size_t)t] = "could not convert string to float: '" + tok + "' (row " + std::to_string(row0 + r) + ")";
          return;
        }
        if (col < num_col) out[r * num_col + col] = v;
        col++;
        p = q + 1;
      }
      if (col != num_col) {
        err_row[(size_t)t] = r;
        errs[(size_t)t] = "row " + std::to_string(row0 + r) + " has " + std::to_string(col) + " values, expected " + std::to_string(num_col);
        return;
      }
    }
  };
  if (nt == 1) work(0);
  else {
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back(work, t);
    for (auto& t : th) t.join();
  }
  int best = -1;
  for (int t = 0; t < nt; t++)
    if (err_row[(size_t)t] != INT64_MAX && (best < 0 || err_row[(size_t)t] < err_row[(size_t)best])) best = t;
  if (best >= 0) throw InputFileMalformed(errs[(size_t)best]);
  return rows;
}

}  // namespace

int cgc_parse_text(cgc_context* ctx, const char* text, int64_t nbytes, int* num_col, double* values, int64_t cap_values,
                   int64_t* rows_out, double* kernel_ms) {
  if (rows_out) *rows_out = 0;
  if (kernel_ms) *kernel_ms = 0.0;
  if (!ctx || !text || !num_col || !rows_out) return fail(ctx, CGC_ERR_ARG, "null argument");
  if (nbytes < 0 || *num_col < 0 || cap_values < 0) return fail(ctx, CGC_ERR_ARG, "negative size");
  return guarded(ctx, [&]() {
    // what f.strip() does in the reference's reader (scripts/2014-06-csv2hdf5.py:57): surrounding whitespace goes
    int64_t b = 0, e = nbytes;
    auto is_space = [](char c) { return c == ' ' || c == '\n' || c == '\r' || c == '\t' || c == '\f' || c == '\v'; };
    while (b < e && is_space(text[b])) b++;
    while (e > b && is_space(text[e - 1])) e--;
    if (b == e) return CGC_OK;   // no rows
    int ncol = *num_col;
    if (ncol == 0) {   // columns of the first line
      const void* nl = memchr(text + b, '\n', (size_t)(e - b));
      const int64_t le = nl ? (const char*)nl - text : e;
      ncol = 1;
      for (int64_t p = b; p < le; p++) ncol += text[p] == ',';
      *num_col = ncol;
    }
    if (!values) {   // count only: the rows are the newlines (plus a last line without one), counted on the host
      const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(8, (e - b) >> 24));
      std::vector<int64_t> part((size_t)nt, 0);
      auto count = [&](int t) {
        const int64_t p0 = b + (e - b) * t / nt, p1 = b + (e - b) * (t + 1) / nt;
        part[(size_t)t] = std::count(text + p0, text + p1, '\n');
      };
      if (nt == 1) count(0);
      else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++) th.emplace_back(count, t);
        for (auto& t : th) t.join();
      }
      int64_t rows = 1;   // the stripped text does not end with a newline
      for (int64_t v : part) rows += v;
      *rows_out = rows;
      return CGC_OK;
    }
    need_device(ctx);
    if (!ctx->st_comp) CUDA_OK(cudaStreamCreateWithFlags(&ctx->st_comp, cudaStreamNonBlocking));
    cudaStream_t st = ctx->st_comp;
    // chunks of at most ~1 GiB of text, cut behind a newline
    int64_t chunk_cap = (int64_t)1 << 30;
    if (const char* env = getenv("CGC_PARSE_CHUNK_BYTES")) chunk_cap = std::max<int64_t>(64, atoll(env));   // test hook
    const int64_t biggest = std::min<int64_t>(e - b, chunk_cap) + 1;
    char* d_text = nullptr;
    int* d_counts = nullptr;
    int64_t* d_offsets = nullptr;
    unsigned long long* d_tot = nullptr;   // [0] delimiters [1] newlines [2] first bad row
    int* d_flags = nullptr;
    double* d_out = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
    cudaError_t err = cudaSuccess;
    auto ok = [&](cudaError_t x) { if (err == cudaSuccess) err = x; return x == cudaSuccess; };
    std::string problem;
    int problem_code = CGC_OK;
    int64_t rows_done = 0;
    int64_t out_cap = 0;
    float ms_total = 0.f;
    const int64_t tiles_max = parse_text_tiles(biggest);
    if (ok(cudaMalloc(&d_text, (size_t)parse_text_padded_bytes(biggest))) && ok(cudaMalloc(&d_counts, sizeof(int) * (size_t)tiles_max)) &&
        ok(cudaMalloc(&d_offsets, (size_t)parse_text_scan_bytes(biggest))) && ok(cudaMalloc(&d_tot, sizeof(unsigned long long) * 3)) &&
        ok(cudaMalloc(&d_flags, sizeof(int))) && ok(cudaEventCreate(&e0)) && ok(cudaEventCreate(&e1)) && ok(cudaEventCreate(&e2))) {
      for (int64_t p = b; p < e && err == cudaSuccess && problem_code == CGC_OK;) {
        // this chunk: [p, q) of the text plus a final newline if the text does not supply one
        int64_t q = std::min(e, p + chunk_cap);
        if (q < e) {
          const void* last = memrchr(text + p, '\n', (size_t)(q - p));
          if (last) q = (const char*)last - text + 1;
          else {   // a line longer than the chunk: take it whole
            const void* nl = memchr(text + q, '\n', (size_t)(e - q));
            q = nl ? (const char*)nl - text + 1 : e;
          }
        }
        const bool add_nl = text[q - 1] != '\n';
        const int64_t n = q - p + (add_nl ? 1 : 0);
        if (parse_text_padded_bytes(n) > parse_text_padded_bytes(biggest)) {   // only after the long-line case above
          problem = "parse_text: a single line exceeds the chunk size";
          problem_code = CGC_ERR_UNSUPPORTED;
          break;
        }
        const char nlc = '\n';
        if (!ok(cudaMemcpyAsync(d_text, text + p, (size_t)(q - p), cudaMemcpyHostToDevice, st))) break;
        if (add_nl && !ok(cudaMemcpyAsync(d_text + (q - p), &nlc, 1, cudaMemcpyHostToDevice, st))) break;
        if (!ok(cudaMemsetAsync(d_text + n, 0, (size_t)(parse_text_padded_bytes(n) - n), st))) break;
        ok(cudaEventRecord(e0, st));
        launch_parse_count(d_text, n, d_counts, d_tot, ctx->sm_count, st);
        ctx->launches += 1;
        ok(cudaEventRecord(e1, st));
        unsigned long long tot[2] = {0, 0};
        ok(cudaMemcpyAsync(tot, d_tot, sizeof(tot), cudaMemcpyDeviceToHost, st));
        if (!ok(cudaStreamSynchronize(st)) || !ok(cudaGetLastError())) break;
        const int64_t rows = (int64_t)tot[1];
        {
          if ((rows_done + rows) * ncol > cap_values) {
            problem = "parse_text: output buffer too small";
            problem_code = CGC_ERR_ARG;
            break;
          }
          const int64_t nvals = rows * ncol;
          if (nvals > out_cap) {
            if (d_out) cudaFree(d_out);
            d_out = nullptr;
            if (!ok(cudaMalloc(&d_out, sizeof(double) * (size_t)nvals))) break;
            out_cap = nvals;
          }
          float ms_count = 0.f;
          ok(cudaEventElapsedTime(&ms_count, e0, e1));
          ok(cudaEventRecord(e0, st));
          launch_parse(d_text, n, d_counts, d_offsets, ncol, nvals, d_out, d_flags, d_tot + 2, ctx->sm_count, st);
          ctx->launches += 3;
          ok(cudaEventRecord(e2, st));
          int flags = 0;
          unsigned long long bad_row = 0;
          ok(cudaMemcpyAsync(&flags, d_flags, sizeof(int), cudaMemcpyDeviceToHost, st));
          ok(cudaMemcpyAsync(&bad_row, d_tot + 2, sizeof(bad_row), cudaMemcpyDeviceToHost, st));
          ok(cudaMemcpyAsync(values + rows_done * ncol, d_out, sizeof(double) * (size_t)nvals, cudaMemcpyDeviceToHost, st));
          if (!ok(cudaStreamSynchronize(st)) || !ok(cudaGetLastError())) break;
          float ms = 0.f;
          if (ok(cudaEventElapsedTime(&ms, e0, e2))) ms_total += ms + ms_count;
          if (flags != 0 || (int64_t)tot[0] != nvals) {
            // a token for strtod, or a ragged row: the host parses this chunk and either fills it in or names the culprit
            try {
              parse_rows_host(text, p, q, ncol, values + rows_done * ncol, (cap_values / ncol) - rows_done, rows_done);
            } catch (const std::exception& ex) {
              problem = ex.what();
              problem_code = CGC_ERR_INPUT_MALFORMED;
              break;
            }
          }
        }
        rows_done += rows;
        p = q;
      }
    }
    cudaFree(d_text); cudaFree(d_counts); cudaFree(d_offsets); cudaFree(d_tot); cudaFree(d_flags); cudaFree(d_out);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (e2) cudaEventDestroy(e2);
    CUDA_OK(err);
    if (problem_code != CGC_OK) {
      ctx->err = problem;
      return problem_code;
    }
    *rows_out = rows_done;
    if (kernel_ms) *kernel_ms = ms_total;
    return CGC_OK;
  });
}

// ---------------------------------------------------------------------------------------------------------------------
// writers
// ---------------------------------------------------------------------------------------------------------------------
namespace {

int write_matrix_text(const char* path, bool truncate, const float* m, int64_t rows, int ncol, bool to_cm) {
  int fd = open(path, O_WRONLY | O_CREAT | (truncate ? O_TRUNC : O_APPEND), 0644);
  if (fd < 0) return -1;
  const int nt = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  const int64_t per = std::max<int64_t>(1, (((int64_t)2 << 20) / std::max(1, ncol * 9)));  // ~2 MB of text per task
  // groups of nt tasks; group g+1 is formatted by the pool while this thread writes group g
  auto format_group = [&](int64_t base, std::vector<std::string>& parts) {
    std::vector<std::thread> th;
    for (int i = 0; i < nt; i++) {
      int64_t r0 = base + i * per, r1 = std::min(rows, r0 + per);
      if (r0 >= r1) break;
      th.emplace_back([&, i, r0, r1] { format_rows(m, r0, r1, ncol, to_cm, parts[(size_t)i]); });
    }
    for (auto& t : th) t.join();
  };
  std::vector<std::string> cur((size_t)nt), nxt((size_t)nt);
  bool ok = true;
  if (rows > 0) format_group(0, cur);
  for (int64_t base = 0; base < rows && ok; base += per * nt) {
    const int64_t next_base = base + per * nt;
    std::thread ahead;
    if (next_base < rows) ahead = std::thread([&] { format_group(next_base, nxt); });
    for (auto& p : cur) {
      size_t off = 0;
      while (off < p.size()) {
        ssize_t w = write(fd, p.data() + off, p.size() - off);
        if (w <= 0) { ok = false; break; }
        off += (size_t)w;
      }
      p.clear();
      if (!ok) break;
    }
    if (ahead.joinable()) ahead.join();
    std::swap(cur, nxt);
  }
  if (close(fd) != 0) ok = false;
  return ok ? 0 : -1;
}

}  // namespace

int cgc_write_time(const char* path, int truncate, const float* dE_kcal, int64_t rows, int num_tot) {
  if (!path || (rows > 0 && !dE_kcal) || num_tot < 1) return CGC_ERR_ARG;
  if (write_matrix_text(path, truncate != 0, dE_kcal, rows, num_tot, true) != 0) return CGC_ERR_READ;
  return CGC_OK;
}

int cgc_write_mtx(cgc_context* ctx, const char* path, int compat) {
  if (!ctx || !path) return fail(ctx, CGC_ERR_ARG, "null argument");
  return guarded(ctx, [&]() {
    need_device(ctx);
    if (!ctx->d_cov) throw std::invalid_argument("no covariance sums on this context");
    const int S = ctx->sys.n_sites, G = ctx->sys.num_tot;
    std::vector<float> m;
    if (!compat) {
      // covariance -> float32 -> "%g" text, all on the device (K5); the text leaves through the pinned sink buffers, a
      // piece per D2H copy, while the writer thread puts the previous pieces into the file.  A value the device formatter
      // does not cover (|v| < 1e-16 and not 0, or non-finite) sends the whole matrix to the host formatter.
      const int64_t n = (int64_t)S * G * G, rows = (int64_t)S * G, cap = format_time_capacity(rows, G);
      double *d_mean = nullptr, *d_covm = nullptr;
      float* d_m32 = nullptr;
      char* d_txt = nullptr;
      int64_t* d_off = nullptr;
      int64_t total = -1;
      int flag = 0;
      bool wrote = true;
      cudaError_t err = cudaSuccess;
      auto ok = [&](cudaError_t e) { if (err == cudaSuccess) err = e; return e == cudaSuccess; };
      CUDA_OK(cudaDeviceSynchronize());
      if (ok(cudaMalloc(&d_mean, sizeof(double) * S * G)) && ok(cudaMalloc(&d_covm, sizeof(double) * n)) &&
          ok(cudaMalloc(&d_m32, sizeof(float) * n))) {
        launch_cov_finalize(ctx->d_cov, S, G, 0, d_mean, d_covm, ctx->st_comp);
        launch_narrow(d_covm, d_m32, n, ctx->st_comp);
        ctx->launches += 2;
        ok(cudaMemsetAsync(ctx->d_fmt_flags, 0, sizeof(int), ctx->st_comp));
        if (ok(cudaStreamSynchronize(ctx->st_comp))) {
          cudaFree(d_covm);
          d_covm = nullptr;
          if (cudaMalloc(&d_txt, (size_t)cap) == cudaSuccess && cudaMalloc(&d_off, sizeof(int64_t) * (rows + 1)) == cudaSuccess) {
            launch_format_time(d_m32, rows, G, false, d_off, d_txt, ctx->d_fmt_flags, ctx->st_comp);
            ctx->launches += 3;
            ok(cudaMemcpyAsync(&total, d_off + rows, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->st_comp));
            ok(cudaMemcpyAsync(&flag, ctx->d_fmt_flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->st_comp));
            ok(cudaStreamSynchronize(ctx->st_comp));
            if (err == cudaSuccess && flag == 0 && total >= 0 && total <= cap) {
              int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
              if (fd < 0) {
                wrote = false;
              } else {
                try {
                  ensure_sink_bufs(ctx, (int64_t)16 << 20);
                  TextSink sink(fd, ctx);
                  for (int64_t off = 0; off < total && err == cudaSuccess; off += ctx->sink_cap) {
                    const int64_t nb = std::min<int64_t>(ctx->sink_cap, total - off);
                    const int b = sink.acquire();
                    ok(cudaMemcpyAsync(ctx->sink_buf[b], d_txt + off, (size_t)nb, cudaMemcpyDeviceToHost, ctx->st_out));
                    ok(cudaEventRecord(ctx->sink_ev[b], ctx->st_out));
                    sink.push_buffer(b, nb);
                  }
                  wrote = sink.finish();
                } catch (...) {
                  close(fd);
                  cudaFree(d_mean); cudaFree(d_m32); cudaFree(d_txt); cudaFree(d_off);
                  throw;
                }
                if (close(fd) != 0) wrote = false;
              }
            } else {
              total = -1;
            }
          } else {
            cudaGetLastError();   // not enough device memory for the text: the host formatter takes over
          }
          if (total < 0 && err == cudaSuccess) {
            m.resize((size_t)n);
            ok(cudaMemcpy(m.data(), d_m32, sizeof(float) * n, cudaMemcpyDeviceToHost));
          }
          if (flag) zero_now(ctx->d_fmt_flags, sizeof(int));
        }
      }
      cudaFree(d_mean); cudaFree(d_covm); cudaFree(d_m32); cudaFree(d_txt); cudaFree(d_off);
      CUDA_OK(err);
      if (total >= 0) {
        if (!wrote) throw ReadError(std::string("cannot write ") + path);
        return CGC_OK;
      }
    } else {
      if (!ctx->compat_on) {
        throw std::invalid_argument("the literal .mtx needs the float32 sums of every frame: set the option \"mtx_compat\" "
                                    "(traj2dEcsv --mtx-compat) before the first cgc_run");
      }
      m.resize((size_t)S * G * G);
      // the reference's literal arithmetic (src/FMOparse.cpp:189-206 with dEsize == 0), float32 throughout
      CUDA_OK(cudaDeviceSynchronize());
      double n = 0;
      CUDA_OK(cudaMemcpy(&n, ctx->d_cov, sizeof(double), cudaMemcpyDeviceToHost));
      const int T = (int)n;
      std::vector<float> cs((size_t)2 * S * G);
      CUDA_OK(cudaMemcpy(cs.data(), ctx->d_compat, sizeof(float) * cs.size(), cudaMemcpyDeviceToHost));
      for (int k = 0; k < S; k++) {
        float* sum = cs.data() + (size_t)k * G;
        float* row0 = cs.data() + (size_t)S * G + (size_t)k * G;
        for (int i = 0; i < G; i++) {
          sum[i] /= T;
          for (int j = 0; j < G; j++) row0[j] /= T;  // the same row, divided once per i
        }
        for (int i = 0; i < G; i++)
          for (int j = 0; j < G; j++) {
            volatile float prod = sum[i] * sum[j];  // keep the product a rounded float32 (no FMA contraction)
            m[((size_t)k * G + i) * G + j] = row0[j] - prod;
          }
      }
    }
    if (write_matrix_text(path, true, m.data(), (int64_t)S * G, G, false) != 0) throw ReadError(std::string("cannot write ") + path);
    return CGC_OK;
  });
}

}  // extern "C"

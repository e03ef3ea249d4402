// traj2dEcsv — drop-in CLI driver on top of the C-ABI (include/cgromcorr.h).
//
// Same contract as the reference program (reference README.md:25-31, src/FMOparse.cpp:304-378):
//     traj2dEcsv traj_in topology excited_itp csv_out num_frames
// writes csv_out.time (one line per frame and site, cm^-1, "%g") and csv_out.mtx.  Differences, all opt-out/opt-in:
//   * every chromophore is a site (the reference hard-codes 7, src/FMOparse.cpp:133): --sites 7 restores that;
//   * csv_out.mtx holds the per-site population covariance the reference's code structure intends; --mtx-compat writes
//     the reference's literal arithmetic instead (its stride variable stays 0, src/FMOparse.cpp:42,193,202);
//     BYTE-LEVEL drop-in for the stock program's files therefore needs exactly two flags: --sites 7 --mtx-compat;
//   * asking for more frames than the file holds stops cleanly (the reference segfaults, README.md:31);
//   * stdout is quiet unless --verbose (per-frame site totals in cm^-1, the reference's "=<sum> cm-1" log).
//   * traj_in may be a GROMACS .trr (binary) or .xtc (compressed, decoded on the device) when --structure some.gro names
//     a .gro of the same system: the authors' trjconv step (reference scripts/2014-09-trr2dEcsv.sh:22) and its 45-69
//     bytes of text per atom are skipped; an .xtc moves 3-6 bytes per atom and frame to the GPU.
//   * csv_out.time is formatted on the device (cgc_run_write_time); --host-format uses the host formatter instead
//     (cgc_run + cgc_write_time, what the first version did; the bytes are the same);
//   * --checkpoint FILE saves, after every chunk of frames, how far the run got together with the Correlate sums;
//     --resume FILE continues such a run: csv_out.time is cut back to the checkpointed length and appended to, the sums
//     carry on (the reference loses its accumulators when it is killed);
//   * --npy additionally writes csv_out.time.npy: the same rows as raw float32 (frames, sites, numTot) in kcal/mol, the
//     array the pyPCA scripts rebuild from the text (reference scripts/2014-06-csv2hdf5.py:56-74) without the text.
//   * --gpus N (or --devices a,b,c): ONE process drives N GPUs, one host thread and one context per GPU.  The frames are
//     cut into N contiguous ranges — what the authors did by hand with one process per trajectory chunk file in a bash
//     loop (reference scripts/2014-06-gro2dEcsv.sh:3-27) — every GPU writes csv_out.time.part<r>, the Correlate sums are
//     combined by ONE ncclAllReduce(double) (cgc_cov_allreduce), then the parts are joined in rank order into
//     csv_out.time and device 0 writes csv_out.mtx.  The rows are bit-identical to a one-GPU run (K2 adds grid-rounded partial sums: exact, order-independent).
// Extra flags come AFTER the five positionals.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/cgromcorr.h"

static volatile sig_atomic_t g_keep_running = 1;
static cgc_context* volatile g_ctx[16] = {};   // contexts a SIGINT must reach
static void on_sigint(int) {
  const char msg[] = "Received cancellation, exiting after the current batch...\n";
  ssize_t r = write(2, msg, sizeof(msg) - 1);
  (void)r;
  g_keep_running = 0;
  for (int i = 0; i < 16; i++)
    if (g_ctx[i]) cgc_request_stop(g_ctx[i]);
}

// NumPy .npy v1.0 header for a C-order float32 array (frames, sites, num_tot); rewritten in place once `frames` is known
// (the header is padded to a fixed 128 bytes so that the rewrite does not move the data).
static bool write_npy_header(FILE* f, long long frames, int sites, int num_tot) {
  char dict[112];
  int n = snprintf(dict, sizeof dict, "{'descr': '<f4', 'fortran_order': False, 'shape': (%lld, %d, %d), }", frames, sites, num_tot);
  if (n < 0 || n > 116) return false;
  unsigned char head[128];
  memset(head, ' ', sizeof head);
  memcpy(head, "\x93NUMPY\x01\x00", 8);
  head[8] = 118;  // header length (little endian): 128 - 10
  head[9] = 0;
  memcpy(head + 10, dict, (size_t)n);
  head[127] = '\n';
  return fseek(f, 0, SEEK_SET) == 0 && fwrite(head, 1, sizeof head, f) == sizeof head;
}

// Checkpoint file: header + the blob of cgc_cov_save_state.  Written to FILE.tmp and renamed, so a kill leaves either the
// previous checkpoint or the new one.
struct CkptHeader {
  char magic[8];          // "CGCCKPT1"
  int64_t frames_done;    // frames processed since frame `first`
  int64_t first_frame;
  int64_t time_bytes;     // length of csv_out.time at that point
  int32_t sites, num_tot;
  int64_t state_bytes;    // 0: covariance off
};

static bool save_checkpoint(cgc_context* ctx, const std::string& path, long long frames_done, long long first,
                            const std::string& time_path, int sites, int num_tot) {
  CkptHeader h{};
  memcpy(h.magic, "CGCCKPT1", 8);
  h.frames_done = frames_done;
  h.first_frame = first;
  FILE* t = fopen(time_path.c_str(), "rb");
  if (!t) return false;
  fseek(t, 0, SEEK_END);
  h.time_bytes = ftell(t);
  fclose(t);
  h.sites = sites;
  h.num_tot = num_tot;
  h.state_bytes = cgc_cov_state_bytes(ctx);
  std::vector<char> blob((size_t)h.state_bytes);
  if (h.state_bytes > 0 && cgc_cov_save_state(ctx, blob.data(), h.state_bytes) != CGC_OK) return false;
  const std::string tmp = path + ".tmp";
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) return false;
  bool ok = fwrite(&h, sizeof h, 1, f) == 1 && (blob.empty() || fwrite(blob.data(), 1, blob.size(), f) == blob.size());
  ok = fclose(f) == 0 && ok;
  return ok && rename(tmp.c_str(), path.c_str()) == 0;
}

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static const char* error_class(int rc) {
  return rc == CGC_ERR_INPUT_MALFORMED ? "InputFileMalformed"
       : rc == CGC_ERR_READ           ? "ReadError"
       : rc == CGC_ERR_UNSUPPORTED    ? "Unsupported"
       : rc == CGC_ERR_CUDA           ? "CudaError"
                                      : "Error";
}

// the reference dies by an uncaught exception (SIGABRT, shell status 134) on the first two classes
static int exit_status(int rc) { return (rc == CGC_ERR_INPUT_MALFORMED || rc == CGC_ERR_READ) ? 134 : EXIT_FAILURE; }

static int die(cgc_context* ctx, int rc, const char* what) {
  fprintf(stderr, "traj2dEcsv: %s while %s: %s\n", error_class(rc), what, cgc_last_error(ctx));
  if (ctx) cgc_destroy(ctx);
  return exit_status(rc);
}

// Every teardown step the operating system does anyway at exit (unpinning host buffers, unmapping the trajectory,
// destroying the CUDA context) is left to it: the outputs are complete and closed when this is called.
[[noreturn]] static void finish(int status) {
  fflush(stdout);
  fflush(stderr);
  _exit(status);
}

struct Options {
  std::string traj, top, itp, out, structure, ckpt_path, resume_path;
  long long num_frames = 0, first = 0;
  int sites = 0, batch = 0, stage_threads = 0, ring_slots = 0;
  std::vector<int> devices;
  bool compat = false, no_mtx = false, verbose = false, timing = false, npy = false, host_format = false;
  bool halve_batch = false;   // several contexts in one process: half the default batch (see run_multi)
  bool xtc = false;           // traj_in is a GROMACS .xtc (by its extension)
};

// cgc_create + open + options for one device; on failure prints the reason and returns the exit status through *status
static cgc_context* open_context(const Options& o, int device, bool announce, int* status) {
  cgc_context* ctx = nullptr;
  int rc = cgc_create(o.top.c_str(), o.itp.c_str(), device, &ctx);
  if (rc != CGC_OK) {
    *status = die(nullptr, rc, "reading the topology");
    return nullptr;
  }
  if (announce && cgc_topology_warnings(ctx) > 0) {
    printf("Minor issues found with the topology input (%d conflicting duplicate entries; first entry kept)\n",
           cgc_topology_warnings(ctx));
  }
  rc = o.structure.empty() ? cgc_open(ctx, o.traj.c_str())
       : o.xtc           ? cgc_open_xtc(ctx, o.traj.c_str(), o.structure.c_str())
                         : cgc_open_trr(ctx, o.traj.c_str(), o.structure.c_str());
  if (rc != CGC_OK) {
    *status = die(ctx, rc, "opening the trajectory");
    return nullptr;
  }
  if (o.sites > 0 && (rc = cgc_set_option(ctx, "sites", o.sites)) != CGC_OK) {
    *status = die(ctx, rc, "setting --sites");
    return nullptr;
  }
  if (o.compat && (rc = cgc_set_option(ctx, "mtx_compat", 1)) != CGC_OK) {
    *status = die(ctx, rc, "setting --mtx-compat");
    return nullptr;
  }
  if (o.no_mtx && (rc = cgc_set_option(ctx, "covariance", 0)) != CGC_OK) {
    *status = die(ctx, rc, "setting --no-mtx");
    return nullptr;
  }
  int batch = o.batch;
  if (batch <= 0 && o.halve_batch) batch = std::max(1, (int)cgc_get_option(ctx, "batch_frames") / 2);
  if (batch > 0 && (rc = cgc_set_option(ctx, "batch_frames", batch)) != CGC_OK) {
    *status = die(ctx, rc, "setting --batch-frames");
    return nullptr;
  }
  if (o.stage_threads > 0 && (rc = cgc_set_option(ctx, "stage_threads", o.stage_threads)) != CGC_OK) {
    *status = die(ctx, rc, "setting --stage-threads");
    return nullptr;
  }
  if (o.ring_slots > 0 && (rc = cgc_set_option(ctx, "ring_slots", o.ring_slots)) != CGC_OK) {
    *status = die(ctx, rc, "setting --ring-slots");
    return nullptr;
  }
  // buffers before the frame loop (the device text path unless the host formatter was asked for)
  if ((rc = cgc_prepare(ctx, o.host_format ? 0 : 1)) != CGC_OK) {
    *status = die(ctx, rc, "allocating the streaming buffers");
    return nullptr;
  }
  return ctx;
}

// ---------------------------------------------------------------------------------------------------------------------
// one GPU
// ---------------------------------------------------------------------------------------------------------------------
static int run_single(const Options& o) {
  const double t_start = now_s();
  int status = EXIT_FAILURE;
  cgc_context* ctx = open_context(o, o.devices[0], true, &status);
  if (!ctx) return status;
  g_ctx[0] = ctx;
  const double t_opened0 = now_s();
  int rc = CGC_OK;
  int n_atoms = 0, n_chromo = 0, n_groups = 0, num_tot = 0;
  cgc_dims(ctx, &n_atoms, &n_chromo, &n_groups, &num_tot);
  const int S = (int)cgc_get_option(ctx, "sites");
  const bool want_mtx = !o.no_mtx && cgc_cov_size(ctx) > 0;
  if (!o.no_mtx && !want_mtx) {
    fprintf(stderr, "traj2dEcsv: note: %d x %d^2 covariance sums do not fit; csv_out.mtx is not written\n", S, num_tot);
  }
  if (o.verbose) printf("atoms %d  nChromo %d  nGroups %d  sites %d\n", n_atoms, n_chromo, n_groups, S);

  const std::string time_path = o.out + ".time";
  const long long first = o.first, num_frames = o.num_frames;
  long long resumed_frames = 0;
  if (!o.resume_path.empty()) {
    FILE* f = fopen(o.resume_path.c_str(), "rb");
    CkptHeader h{};
    std::vector<char> blob;
    bool ok = f && fread(&h, sizeof h, 1, f) == 1 && memcmp(h.magic, "CGCCKPT1", 8) == 0 && h.sites == S && h.num_tot == num_tot &&
              h.first_frame == first && h.state_bytes == cgc_cov_state_bytes(ctx);
    if (ok && h.state_bytes > 0) {
      blob.resize((size_t)h.state_bytes);
      ok = fread(blob.data(), 1, blob.size(), f) == blob.size() && cgc_cov_load_state(ctx, blob.data(), h.state_bytes) == CGC_OK;
    }
    if (f) fclose(f);
    if (!ok || truncate(time_path.c_str(), (off_t)h.time_bytes) != 0) {
      fprintf(stderr, "traj2dEcsv: cannot resume from %s (missing, damaged, or written for other sites / columns / first frame)\n",
              o.resume_path.c_str());
      cgc_destroy(ctx);
      return EXIT_FAILURE;
    }
    resumed_frames = h.frames_done;
    if (o.verbose) printf("resuming after %lld frames\n", resumed_frames);
  } else if (cgc_write_time(time_path.c_str(), 1, nullptr, 0, num_tot) != CGC_OK) {  // truncate (reference src/FMOparse.cpp:48-50)
    fprintf(stderr, "traj2dEcsv: cannot write %s\n", time_path.c_str());
    cgc_destroy(ctx);
    return EXIT_FAILURE;
  }
  FILE* npy_file = nullptr;
  if (o.npy) {
    npy_file = fopen((time_path + ".npy").c_str(), resumed_frames > 0 ? "r+b" : "wb");
    if (npy_file && resumed_frames > 0) {
      // keep the rows of the frames already done; anything after them is rewritten
      if (fseek(npy_file, 128L + (long)(resumed_frames * (long long)S * num_tot * 4), SEEK_SET) != 0) { fclose(npy_file); npy_file = nullptr; }
    } else if (npy_file && !write_npy_header(npy_file, 0, S, num_tot)) {
      fclose(npy_file);
      npy_file = nullptr;
    }
    if (!npy_file) {
      fprintf(stderr, "traj2dEcsv: cannot write %s.npy\n", time_path.c_str());
      cgc_destroy(ctx);
      return EXIT_FAILURE;
    }
  }
  signal(SIGINT, on_sigint);
  const double t_opened = now_s();
  double t_run = 0, t_wait_writer = 0, t_best_chunk = 1e30;
  long long best_chunk_frames = 0;

  const char* pause_env = getenv("CGC_CLI_CHUNK_PAUSE_MS");   // test hook: lets a test land a SIGINT inside the loop
  const int pause_ms = pause_env ? atoi(pause_env) : 0;
  const bool need_rows = o.host_format || o.verbose || npy_file != nullptr;
  // The whole range goes through ONE cgc_run_write_time call (one pipeline fill and drain) unless something has to
  // happen between chunks: rows handed to a host-side consumer (--verbose, --npy, --host-format), checkpoints, the test hook.
  const bool chunked = need_rows || !o.ckpt_path.empty() || pause_ms > 0;
  const long long chunk = chunked ? 16LL * (long long)cgc_get_option(ctx, "batch_frames") : num_frames;
  const size_t row = (size_t)S * num_tot;
  std::vector<float> buf[2];
  std::thread writer;
  int wrc = CGC_OK;
  long long done_total = resumed_frames;
  int which = 0;
  bool eof = false;
  while (done_total < num_frames && g_keep_running && !eof) {
    if (pause_ms > 0 && done_total > resumed_frames) std::this_thread::sleep_for(std::chrono::milliseconds(pause_ms));
    if (!g_keep_running) break;
    const long long n = std::min(chunk, num_frames - done_total);
    if (need_rows) buf[which].resize((size_t)n * row);
    int64_t done = 0;
    double t0 = now_s();
    if (o.host_format) rc = cgc_run(ctx, first + done_total, n, buf[which].data(), &done);
    else rc = cgc_run_write_time(ctx, first + done_total, n, time_path.c_str(), 0, need_rows ? buf[which].data() : nullptr, &done);
    const double t_chunk = now_s() - t0;
    t_run += t_chunk;
    if (done == n && t_chunk / (double)n < t_best_chunk / (double)std::max<long long>(1, best_chunk_frames)) {
      t_best_chunk = t_chunk;
      best_chunk_frames = n;
    }
    t0 = now_s();
    if (rc != CGC_OK && rc != CGC_EOF && rc != CGC_STOPPED) {
      if (writer.joinable()) writer.join();
      return die(ctx, rc, "processing frames");
    }
    eof = rc == CGC_EOF;
    if (writer.joinable()) writer.join();
    t_wait_writer += now_s() - t0;
    if (wrc != CGC_OK) break;
    const float* data = buf[which].data();
    const long long base = done_total;
    const bool host_format = o.host_format, verbose = o.verbose;
    if (need_rows) {
      writer = std::thread([=, &wrc] {
        if (host_format) wrc = cgc_write_time(time_path.c_str(), 0, data, done * S, num_tot);
        if (npy_file && fwrite(data, sizeof(float), (size_t)done * row, npy_file) != (size_t)done * row) wrc = CGC_ERR_READ;
        if (verbose) {
          for (long long t = 0; t < done; t++) {
            printf("timeCount:%lld Computing site", base + t);
            for (int s = 0; s < S; s++) {
              double sum = 0;
              for (int j = 0; j < num_tot; j++) sum += (double)(float)((double)data[(t * S + s) * (size_t)num_tot + j] * 349.7);
              printf(" %d=%g cm-1", s + 1, sum);
            }
            printf("\n");
          }
        }
      });
    }
    done_total += done;
    which ^= 1;
    if (rc == CGC_STOPPED) break;
    if (!o.ckpt_path.empty()) {
      if (writer.joinable()) writer.join();   // the .time / .npy bytes of this chunk are in their files
      if (npy_file) fflush(npy_file);
      if (!save_checkpoint(ctx, o.ckpt_path, done_total, first, time_path, S, num_tot)) {
        fprintf(stderr, "traj2dEcsv: cannot write the checkpoint %s\n", o.ckpt_path.c_str());
      }
    }
  }
  const double t_loop_end = now_s();
  if (writer.joinable()) writer.join();
  if (npy_file) {
    if (!write_npy_header(npy_file, done_total, S, num_tot)) wrc = CGC_ERR_READ;
    fclose(npy_file);
  }
  const double t_written = now_s();
  if (wrc != CGC_OK) {
    fprintf(stderr, "traj2dEcsv: cannot write %s\n", time_path.c_str());
    cgc_destroy(ctx);
    return EXIT_FAILURE;
  }
  if (done_total < num_frames && eof) {
    fprintf(stderr, "traj2dEcsv: trajectory ended after %lld of %lld frames\n", done_total, num_frames);
  }
  if (want_mtx && done_total > 0) {
    printf("Saving to file %s.mtx\n", o.out.c_str());
    rc = cgc_write_mtx(ctx, (o.out + ".mtx").c_str(), o.compat ? 1 : 0);
    if (rc != CGC_OK) return die(ctx, rc, "writing the .mtx file");
  }
  const double t_mtx = now_s();
  if (o.timing) {
    fprintf(stderr,
            "timing: tables+cuda init+open %.3f s | output files %.3f s | frame loop %.3f s (cgc_run %.3f, "
            "waiting for the row writer %.3f) | last rows %.3f s | .mtx %.3f s | frames %lld | in-loop %.0f frames/s | best chunk: %lld frames in %.3f s = %.0f frames/s\n",
            t_opened0 - t_start, t_opened - t_opened0, t_loop_end - t_opened, t_run, t_wait_writer, t_written - t_loop_end,
            t_mtx - t_written, done_total, (double)(done_total - resumed_frames) / std::max(1e-9, t_loop_end - t_opened),
            best_chunk_frames, t_best_chunk, best_chunk_frames / t_best_chunk);
  }
  finish(0);
}

// ---------------------------------------------------------------------------------------------------------------------
// several GPUs in one process
// ---------------------------------------------------------------------------------------------------------------------
static bool append_file(int dst, const std::string& src_path) {
  int src = open(src_path.c_str(), O_RDONLY);
  if (src < 0) return false;
  struct stat sb;
  bool ok = fstat(src, &sb) == 0;
  off_t left = ok ? sb.st_size : 0;
  while (ok && left > 0) {   // in-kernel copy; falls back to read/write where the file system has no copy_file_range
    ssize_t n = copy_file_range(src, nullptr, dst, nullptr, (size_t)std::min<off_t>(left, (off_t)1 << 30), 0);
    if (n <= 0) break;
    left -= n;
  }
  if (ok && left > 0) {
    std::vector<char> buf((size_t)8 << 20);
    while (left > 0) {
      ssize_t n = read(src, buf.data(), buf.size());
      if (n <= 0) { ok = false; break; }
      ssize_t off = 0;
      while (off < n) {
        ssize_t w = write(dst, buf.data() + off, (size_t)(n - off));
        if (w <= 0) { ok = false; break; }
        off += w;
      }
      if (!ok) break;
      left -= n;
    }
  }
  close(src);
  return ok && left == 0;
}

static int run_multi(const Options& o) {
  const int R = (int)o.devices.size();
  const double t_start = now_s();
  if (o.compat) {
    fprintf(stderr, "traj2dEcsv: --mtx-compat reproduces the reference's sequential float32 sums and cannot be combined "
                    "across GPUs; run it on one GPU\n");
    return EXIT_FAILURE;
  }
  if (o.verbose || o.npy || o.host_format || !o.ckpt_path.empty() || !o.resume_path.empty()) {
    fprintf(stderr, "traj2dEcsv: --verbose, --npy, --host-format, --checkpoint and --resume are one-GPU options\n");
    return EXIT_FAILURE;
  }
  // the NCCL communicator takes about a second to set up: that happens next to the frame loop, not after it
  cgc_comm* comm = nullptr;
  int comm_rc = CGC_OK;
  std::string comm_err;
  std::thread comm_thread;
  if (!o.no_mtx) {
    comm_thread = std::thread([&] {
      comm_rc = cgc_comm_create(o.devices.data(), R, &comm);
      if (comm_rc != CGC_OK) comm_err = cgc_last_error(nullptr);
    });
  }
  struct JoinComm {
    std::thread& t;
    ~JoinComm() { if (t.joinable()) t.join(); }
  } join_comm{comm_thread};
  std::vector<cgc_context*> ctx((size_t)R, nullptr);
  std::vector<int> status((size_t)R, 0);
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  Options per = o;
  if (per.stage_threads <= 0) per.stage_threads = (int)std::max(2u, std::min(16u, hw / (unsigned)R));
  // R contexts in one process: every page-locked buffer is mapped into all R devices, so the ring costs R times what it
  // costs alone, and with hw / R copier threads per GPU a batch is staged slowly anyway: a smaller ring of smaller batches
  if (per.ring_slots <= 0) per.ring_slots = 4;
  per.halve_batch = true;
  {
    std::vector<std::thread> th;
    for (int r = 0; r < R; r++) th.emplace_back([&, r] { ctx[(size_t)r] = open_context(per, o.devices[(size_t)r], r == 0, &status[(size_t)r]); });
    for (auto& t : th) t.join();
  }
  for (int r = 0; r < R; r++)
    if (!ctx[(size_t)r]) return status[(size_t)r] ? status[(size_t)r] : EXIT_FAILURE;
  for (int r = 0; r < R && r < 16; r++) g_ctx[r] = ctx[(size_t)r];
  signal(SIGINT, on_sigint);
  int num_tot = 0;
  cgc_dims(ctx[0], nullptr, nullptr, nullptr, &num_tot);
  const int S = (int)cgc_get_option(ctx[0], "sites");
  const bool want_mtx = !o.no_mtx && cgc_cov_size(ctx[0]) > 0;
  if (!o.no_mtx && !want_mtx) {
    fprintf(stderr, "traj2dEcsv: note: %d x %d^2 covariance sums do not fit; csv_out.mtx is not written\n", S, num_tot);
  }
  // frames actually present (the ranges are cut from what exists: the reference would overrun, README.md:31)
  int64_t present = 0;
  int rc = cgc_frame_count(ctx[0], &present);
  if (rc != CGC_OK) return die(ctx[0], rc, "indexing the trajectory");
  const long long total = std::max<long long>(0, std::min<long long>(o.num_frames, (long long)present - o.first));
  if (total < o.num_frames) fprintf(stderr, "traj2dEcsv: trajectory ended after %lld of %lld frames\n", total, o.num_frames);
  const double t_opened = now_s();

  std::vector<int> rrc((size_t)R, CGC_OK);
  std::vector<long long> rdone((size_t)R, 0);
  std::vector<double> rsec((size_t)R, 0.0);
  {
    std::vector<std::thread> th;
    for (int r = 0; r < R; r++) {
      th.emplace_back([&, r] {
        const long long lo = total * r / R, hi = total * (r + 1) / R;
        const std::string part = o.out + ".time.part" + std::to_string(r);
        int64_t done = 0;
        const double t0 = now_s();
        // always create (truncate) the part, also for an empty range
        int prc = cgc_write_time(part.c_str(), 1, nullptr, 0, num_tot);
        if (prc == CGC_OK && hi > lo) prc = cgc_run_write_time(ctx[(size_t)r], o.first + lo, hi - lo, part.c_str(), 0, nullptr, &done);
        rsec[(size_t)r] = now_s() - t0;
        rrc[(size_t)r] = prc;
        rdone[(size_t)r] = done;
      });
    }
    for (auto& t : th) t.join();
  }
  const double t_loop_end = now_s();
  long long done_total = 0;
  bool stopped = false;
  for (int r = 0; r < R; r++) {
    if (rrc[(size_t)r] != CGC_OK && rrc[(size_t)r] != CGC_EOF && rrc[(size_t)r] != CGC_STOPPED) {
      fprintf(stderr, "traj2dEcsv: %s on device %d while processing frames: %s\n", error_class(rrc[(size_t)r]), o.devices[(size_t)r],
              cgc_last_error(ctx[(size_t)r]));
      return exit_status(rrc[(size_t)r]);
    }
    stopped = stopped || rrc[(size_t)r] == CGC_STOPPED;
    done_total += rdone[(size_t)r];
  }
  if (comm_thread.joinable()) comm_thread.join();
  if (want_mtx && done_total > 0) {
    if (comm_rc != CGC_OK) {
      fprintf(stderr, "traj2dEcsv: %s while setting up NCCL: %s\n", error_class(comm_rc), comm_err.c_str());
      return exit_status(comm_rc);
    }
    rc = cgc_cov_allreduce(ctx.data(), R, comm);   // the one collective: n, sum(x-c), sum (x-c)(x-c)^T
    if (rc != CGC_OK) return die(ctx[0], rc, "combining the Correlate sums (NCCL all-reduce)");
  }
  const double t_reduced = now_s();
  // csv_out.mtx (device 0) while this thread joins the parts
  int mrc = CGC_OK;
  std::thread mtx_writer;
  if (want_mtx && done_total > 0) {
    printf("Saving to file %s.mtx\n", o.out.c_str());
    mtx_writer = std::thread([&] { mrc = cgc_write_mtx(ctx[0], (o.out + ".mtx").c_str(), 0); });
  }
  bool joined = true;
  if (stopped) {
    fprintf(stderr, "traj2dEcsv: stopped early: every csv_out.time.part<r> holds the first frames of its own range; they are left as they are\n");
  } else {
    int dst = open((o.out + ".time").c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    joined = dst >= 0;
    for (int r = 0; r < R && joined; r++) {
      const std::string part = o.out + ".time.part" + std::to_string(r);
      joined = append_file(dst, part);
      if (joined) unlink(part.c_str());
    }
    if (dst >= 0 && close(dst) != 0) joined = false;
  }
  const double t_joined = now_s();
  if (mtx_writer.joinable()) mtx_writer.join();
  const double t_mtx = now_s();
  if (!joined) {
    fprintf(stderr, "traj2dEcsv: cannot write %s.time\n", o.out.c_str());
    return EXIT_FAILURE;
  }
  if (mrc != CGC_OK) return die(ctx[0], mrc, "writing the .mtx file");
  if (o.timing) {
    double slowest = 0;
    for (double s : rsec) slowest = std::max(slowest, s);
    fprintf(stderr,
            "timing: %d GPUs | tables+cuda init+open %.3f s | frame loop %.3f s (slowest rank %.3f) | all-reduce %.3f s | join parts %.3f s | "
            ".mtx (overlapped with the join) %.3f s | frames %lld | in-loop %.0f frames/s\n",
            R, t_opened - t_start, t_loop_end - t_opened, slowest, t_reduced - t_loop_end, t_joined - t_reduced, t_mtx - t_reduced,
            done_total, (double)done_total / std::max(1e-9, t_loop_end - t_opened));
  }
  finish(0);
}

int main(int argc, char** argv) {
  if (argc < 6) {
    printf("Five inputs required for this program.\n");
    printf("Order of args: fmo_traj_path, fmo_top_path, bcx_itp_path, output_path, num_frames "
           "[--sites N] [--mtx-compat] [--no-mtx] [--device D] [--gpus N | --devices a,b,..] [--first-frame F] [--batch-frames B] "
           "[--stage-threads T] [--ring-slots K] [--verbose] [--structure frame0.gro (traj_in is then a .trr or .xtc)] [--npy] [--host-format] "
           "[--checkpoint FILE] [--resume FILE]\n"
           "Byte-for-byte the stock program's files: add --sites 7 --mtx-compat (defaults: every chromophore is a site, "
           "csv_out.mtx holds the covariance).\n");
    return EXIT_FAILURE;
  }
  Options o;
  o.traj = argv[1]; o.top = argv[2]; o.itp = argv[3]; o.out = argv[4];
  o.num_frames = atoll(argv[5]);
  int device = 0, gpus = 0;
  for (int i = 6; i < argc; i++) {
    std::string a = argv[i];
    auto val = [&](const char* name) -> const char* {
      if (i + 1 >= argc) {
        fprintf(stderr, "traj2dEcsv: %s needs a value\n", name);
        exit(EXIT_FAILURE);
      }
      return argv[++i];
    };
    if (a == "--sites") o.sites = atoi(val("--sites"));
    else if (a == "--device") device = atoi(val("--device"));
    else if (a == "--gpus") gpus = atoi(val("--gpus"));
    else if (a == "--devices") {
      std::string v = val("--devices");
      size_t p = 0;
      while (p < v.size()) {
        size_t q = v.find(',', p);
        if (q == std::string::npos) q = v.size();
        if (q > p) o.devices.push_back(atoi(v.substr(p, q - p).c_str()));
        p = q + 1;
      }
    }
    else if (a == "--first-frame") o.first = atoll(val("--first-frame"));
    else if (a == "--batch-frames") o.batch = atoi(val("--batch-frames"));
    else if (a == "--stage-threads") o.stage_threads = atoi(val("--stage-threads"));
    else if (a == "--ring-slots") o.ring_slots = atoi(val("--ring-slots"));
    else if (a == "--structure") o.structure = val("--structure");
    else if (a == "--checkpoint") o.ckpt_path = val("--checkpoint");
    else if (a == "--resume") o.resume_path = val("--resume");
    else if (a == "--mtx-compat") o.compat = true;
    else if (a == "--no-mtx") o.no_mtx = true;
    else if (a == "--verbose") o.verbose = true;
    else if (a == "--timing") o.timing = true;
    else if (a == "--npy") o.npy = true;
    else if (a == "--host-format") o.host_format = true;
    else {
      fprintf(stderr, "traj2dEcsv: unknown flag %s\n", a.c_str());
      return EXIT_FAILURE;
    }
  }
  if (const char* e = getenv("CGC_SITES")) if (!o.sites) o.sites = atoi(e);
  if (o.devices.empty()) {
    if (gpus > 1) for (int d = 0; d < gpus; d++) o.devices.push_back(d);
    else o.devices.push_back(device);
  }
  if (o.devices.size() > 16) {
    fprintf(stderr, "traj2dEcsv: at most 16 devices\n");
    return EXIT_FAILURE;
  }
  const bool is_trr = o.traj.size() > 4 && o.traj.compare(o.traj.size() - 4, 4, ".trr") == 0;
  o.xtc = o.traj.size() > 4 && o.traj.compare(o.traj.size() - 4, 4, ".xtc") == 0;
  if ((is_trr || o.xtc) && o.structure.empty()) {
    fprintf(stderr, "traj2dEcsv: a %s trajectory holds no atom names: pass --structure <file.gro> (any .gro frame of the same system)\n",
            o.xtc ? ".xtc" : ".trr");
    return EXIT_FAILURE;
  }
  return o.devices.size() > 1 ? run_multi(o) : run_single(o);
}

// traj2dEcsv — drop-in CLI driver on top of the C-ABI (include/cgromcorr.h).
//
// Same contract as the reference program (reference README.md:25-31, src/FMOparse.cpp:304-378):
//     traj2dEcsv traj_in topology excited_itp csv_out num_frames
// writes csv_out.time (one line per frame and site, cm^-1, "%g") and csv_out.mtx.  Differences, all opt-out/opt-in:
//   * every chromophore is a site (the reference hard-codes 7, src/FMOparse.cpp:133): --sites 7 restores that;
//   * csv_out.mtx holds the per-site population covariance the reference's code structure intends; --mtx-compat writes
//     the reference's literal arithmetic instead (its stride variable stays 0, src/FMOparse.cpp:42,193,202);
//   * asking for more frames than the file holds stops cleanly (the reference segfaults, README.md:31);
//   * stdout is quiet unless --verbose (per-frame site totals in cm^-1, the reference's "=<sum> cm-1" log).
//   * traj_in may be a GROMACS .trr (binary) when --structure some.gro names a .gro of the same system: the authors'
//     trjconv step (reference scripts/2014-09-trr2dEcsv.sh:22) and its 45-69 bytes of text per atom are skipped.
//   * csv_out.time is formatted on the device (cgc_run_write_time); --host-format uses the host formatter instead
//     (cgc_run + cgc_write_time, what the first version did; the bytes are the same);
//   * --checkpoint FILE saves, after every chunk of frames, how far the run got together with the Correlate sums;
//     --resume FILE continues such a run: csv_out.time is cut back to the checkpointed length and appended to, the sums
//     carry on (the reference loses its accumulators when it is killed);
//   * --npy additionally writes csv_out.time.npy: the same rows as raw float32 (frames, sites, numTot) in kcal/mol, the
//     array the pyPCA scripts rebuild from the text (reference scripts/2014-06-csv2hdf5.py:56-74) without the text.
// Extra flags come AFTER the five positionals.
#include <unistd.h>

#include <algorithm>
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
static void on_sigint(int) {
  const char msg[] = "Received cancellation, exiting after the current batch...\n";
  ssize_t r = write(2, msg, sizeof(msg) - 1);
  (void)r;
  g_keep_running = 0;
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

static int die(cgc_context* ctx, int rc, const char* what) {
  const char* cls = rc == CGC_ERR_INPUT_MALFORMED ? "InputFileMalformed"
                  : rc == CGC_ERR_READ           ? "ReadError"
                  : rc == CGC_ERR_UNSUPPORTED    ? "Unsupported"
                  : rc == CGC_ERR_CUDA           ? "CudaError"
                                                 : "Error";
  fprintf(stderr, "traj2dEcsv: %s while %s: %s\n", cls, what, cgc_last_error(ctx));
  if (ctx) cgc_destroy(ctx);
  // the reference dies by an uncaught exception (SIGABRT, shell status 134) on the first two classes
  return (rc == CGC_ERR_INPUT_MALFORMED || rc == CGC_ERR_READ) ? 134 : EXIT_FAILURE;
}

int main(int argc, char** argv) {
  if (argc < 6) {
    printf("Five inputs required for this program.\n");
    printf("Order of args: fmo_traj_path, fmo_top_path, bcx_itp_path, output_path, num_frames "
           "[--sites N] [--mtx-compat] [--no-mtx] [--device D] [--first-frame F] [--batch-frames B] [--verbose] "
           "[--structure frame0.gro (traj_in is then a .trr)] [--npy] [--host-format] [--checkpoint FILE] [--resume FILE]\n");
    return EXIT_FAILURE;
  }
  const std::string traj = argv[1], top = argv[2], itp = argv[3], out = argv[4];
  const long long num_frames = atoll(argv[5]);
  int sites = 0, device = 0, batch = 0;
  long long first = 0;
  std::string structure, ckpt_path, resume_path;
  bool compat = false, no_mtx = false, verbose = false, timing = false, npy = false, host_format = false;
  for (int i = 6; i < argc; i++) {
    std::string a = argv[i];
    auto val = [&](const char* name) -> const char* {
      if (i + 1 >= argc) {
        fprintf(stderr, "traj2dEcsv: %s needs a value\n", name);
        exit(EXIT_FAILURE);
      }
      return argv[++i];
    };
    if (a == "--sites") sites = atoi(val("--sites"));
    else if (a == "--device") device = atoi(val("--device"));
    else if (a == "--first-frame") first = atoll(val("--first-frame"));
    else if (a == "--batch-frames") batch = atoi(val("--batch-frames"));
    else if (a == "--structure") structure = val("--structure");
    else if (a == "--checkpoint") ckpt_path = val("--checkpoint");
    else if (a == "--resume") resume_path = val("--resume");
    else if (a == "--mtx-compat") compat = true;
    else if (a == "--no-mtx") no_mtx = true;
    else if (a == "--verbose") verbose = true;
    else if (a == "--timing") timing = true;
    else if (a == "--npy") npy = true;
    else if (a == "--host-format") host_format = true;
    else {
      fprintf(stderr, "traj2dEcsv: unknown flag %s\n", a.c_str());
      return EXIT_FAILURE;
    }
  }
  if (const char* e = getenv("CGC_SITES")) if (!sites) sites = atoi(e);

  const double t_start = now_s();
  cgc_context* ctx = nullptr;
  int rc = cgc_create(top.c_str(), itp.c_str(), device, &ctx);
  if (rc != CGC_OK) return die(nullptr, rc, "reading the topology");
  if (cgc_topology_warnings(ctx) > 0) {
    printf("Minor issues found with the topology input (%d conflicting duplicate entries; first entry kept)\n",
           cgc_topology_warnings(ctx));
  }
  const double t_created = now_s();
  const bool is_trr = traj.size() > 4 && traj.compare(traj.size() - 4, 4, ".trr") == 0;
  if (is_trr && structure.empty()) {
    fprintf(stderr, "traj2dEcsv: a .trr trajectory holds no atom names: pass --structure <file.gro> (any .gro frame of the same system)\n");
    cgc_destroy(ctx);
    return EXIT_FAILURE;
  }
  rc = structure.empty() ? cgc_open(ctx, traj.c_str()) : cgc_open_trr(ctx, traj.c_str(), structure.c_str());
  if (rc != CGC_OK) return die(ctx, rc, "opening the trajectory");
  if (sites > 0 && (rc = cgc_set_option(ctx, "sites", sites)) != CGC_OK) return die(ctx, rc, "setting --sites");
  if (no_mtx && (rc = cgc_set_option(ctx, "covariance", 0)) != CGC_OK) return die(ctx, rc, "setting --no-mtx");
  if (batch > 0 && (rc = cgc_set_option(ctx, "batch_frames", batch)) != CGC_OK) return die(ctx, rc, "setting --batch-frames");
  int n_atoms = 0, n_chromo = 0, n_groups = 0, num_tot = 0;
  cgc_dims(ctx, &n_atoms, &n_chromo, &n_groups, &num_tot);
  const int S = (int)cgc_get_option(ctx, "sites");
  const bool want_mtx = !no_mtx && cgc_cov_size(ctx) > 0;
  if (!no_mtx && !want_mtx) {
    fprintf(stderr, "traj2dEcsv: note: %d x %d^2 covariance sums do not fit; csv_out.mtx is not written\n", S, num_tot);
  }
  if (verbose) printf("atoms %d  nChromo %d  nGroups %d  sites %d\n", n_atoms, n_chromo, n_groups, S);

  const std::string time_path = out + ".time";
  long long resumed_frames = 0;
  if (!resume_path.empty()) {
    FILE* f = fopen(resume_path.c_str(), "rb");
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
              resume_path.c_str());
      cgc_destroy(ctx);
      return EXIT_FAILURE;
    }
    resumed_frames = h.frames_done;
    if (verbose) printf("resuming after %lld frames\n", resumed_frames);
  } else if (cgc_write_time(time_path.c_str(), 1, nullptr, 0, num_tot) != CGC_OK) {  // truncate (reference src/FMOparse.cpp:48-50)
    fprintf(stderr, "traj2dEcsv: cannot write %s\n", time_path.c_str());
    cgc_destroy(ctx);
    return EXIT_FAILURE;
  }
  FILE* npy_file = nullptr;
  if (npy) {
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

  // frames per cgc_run call: long enough to amortise the call's batch-size ramp, short enough that formatting chunk i
  // overlaps computing chunk i+1 from early on
  const long long chunk = 16LL * (long long)cgc_get_option(ctx, "batch_frames");
  const size_t row = (size_t)S * num_tot;
  std::vector<float> buf[2];
  std::thread writer;
  int wrc = CGC_OK;
  long long done_total = resumed_frames;
  int which = 0;
  bool eof = false;
  const char* pause_env = getenv("CGC_CLI_CHUNK_PAUSE_MS");   // test hook: lets a test land a SIGINT inside the loop
  const int pause_ms = pause_env ? atoi(pause_env) : 0;
  while (done_total < num_frames && g_keep_running && !eof) {
    if (pause_ms > 0 && done_total > resumed_frames) std::this_thread::sleep_for(std::chrono::milliseconds(pause_ms));
    if (!g_keep_running) break;
    const long long n = std::min(chunk, num_frames - done_total);
    if (host_format || verbose || npy_file != nullptr) buf[which].resize((size_t)n * row);
    int64_t done = 0;
    double t0 = now_s();
    const bool need_rows = host_format || verbose || npy_file != nullptr;
    if (host_format) rc = cgc_run(ctx, first + done_total, n, buf[which].data(), &done);
    else rc = cgc_run_write_time(ctx, first + done_total, n, time_path.c_str(), 0, need_rows ? buf[which].data() : nullptr, &done);
    const double t_chunk = now_s() - t0;
    t_run += t_chunk;
    if (done == n && t_chunk / (double)n < t_best_chunk / (double)std::max<long long>(1, best_chunk_frames)) {
      t_best_chunk = t_chunk;
      best_chunk_frames = n;
    }
    t0 = now_s();
    if (rc != CGC_OK && rc != CGC_EOF) {
      if (writer.joinable()) writer.join();
      return die(ctx, rc, "processing frames");
    }
    eof = rc == CGC_EOF;
    if (writer.joinable()) writer.join();
    t_wait_writer += now_s() - t0;
    if (wrc != CGC_OK) break;
    const float* data = buf[which].data();
    const long long base = done_total;
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
    done_total += done;
    which ^= 1;
    if (!ckpt_path.empty()) {
      if (writer.joinable()) writer.join();   // the .time / .npy bytes of this chunk are in their files
      if (npy_file) fflush(npy_file);
      if (!save_checkpoint(ctx, ckpt_path, done_total, first, time_path, S, num_tot)) {
        fprintf(stderr, "traj2dEcsv: cannot write the checkpoint %s\n", ckpt_path.c_str());
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
    printf("Saving to file %s.mtx\n", out.c_str());
    rc = cgc_write_mtx(ctx, (out + ".mtx").c_str(), compat ? 1 : 0);
    if (rc != CGC_OK) return die(ctx, rc, "writing the .mtx file");
  }
  const double t_mtx = now_s();
  cgc_destroy(ctx);
  if (timing) {
    fprintf(stderr,
            "timing: tables+cuda init %.3f s | open (frame-0 pass, device tables) %.3f s | frame loop %.3f s (cgc_run %.3f, "
            "waiting for the .time writer %.3f) | last .time chunk %.3f s | .mtx %.3f s | teardown %.3f s | frames %lld | best chunk: %lld frames in %.3f s = %.0f frames/s\n",
            t_created - t_start, t_opened - t_created, t_loop_end - t_opened, t_run, t_wait_writer, t_written - t_loop_end,
            t_mtx - t_written, now_s() - t_mtx, done_total, best_chunk_frames, t_best_chunk,
            best_chunk_frames / t_best_chunk);
  }
  return 0;
}

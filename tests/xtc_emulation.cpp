// TEST INFRASTRUCTURE — CPU emulation of the two .xtc kernels (cgromcorrfmo_b200/csrc/xtc_kernel.cu).
//
// The kernels are written as per-lane / per-thread pieces in csrc/xtc_bits.h; this file runs those SAME pieces on the host,
// the 32 lanes of the index warp and the 256 threads of a decode CTA one after the other between the points where the device
// code has a ballot / shuffle / __syncthreads.  It exists because the build container has no GPU: the stream arithmetic,
// the speculative walk, the checkpoint table, the shared-memory staging indices and the write-back are all exercised by the
// CPU test-suite (tests/test_xtc.py) before the kernels ever run on a B200.  Never linked into the product library.
//
//   g++ -O2 -std=c++17 -shared -fPIC -o libxtcemu.so tests/xtc_emulation.cpp
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../cgromcorrfmo_b200/csrc/xtc_bits.h"

using namespace cgc::xtc;

namespace {

// xtc_index_kernel: one warp per frame.  Returns the number of steps the walk took.
int64_t index_frame_warp(const char* text, int64_t text_bytes, int64_t off, int n_atoms, int n_ck, U4* table, int* err) {
  FrameHdr h;
  const uint32_t* hw = frame_block(text, text_bytes, off, n_atoms, h);
  if (hw == nullptr) {
    raise_flag(err, kDecBadXtcBit);
    return 0;
  }
  std::vector<uint32_t> ring_words(kRingWords, kRingPoison);   // the kernel's shared memory
  Ring r;
  ring_init(r, ring_words.data(), hw + 10, (h.nbytes + 3u) >> 2);
  Walk w{0u, 0u, (uint32_t)h.smallidx, 0u, kNoAtom};
  const uint32_t natoms = (uint32_t)n_atoms, large_bits = (uint32_t)h.large_bits;
  bool ok = true, wide = false;
  int64_t steps = 0;
  while (ok && w.atom < natoms) {
    const uint32_t cw = w.pos >> 5;
    while (ring_can_issue(r, cw)) {
      for (uint32_t lane = 0; lane < 32; lane++) ring_issue_lane(r, lane);
      ring_issued(r);
    }
    if (ring_must_wait(r, cw, wide ? kRingLookWide : kRingLook)) ring_wait(r);
    steps++;
    if (wide) {
      const uint32_t len = large_bits + 1u + w.nsm * w.smallidx, gs = 1u + w.nsm;
      uint32_t set[kWide], valid[kWide];
      for (uint32_t k = 0; k < kWide; k++) {
        set[k] = valid[k] = 0u;
        for (uint32_t lane = 0; lane < 32; lane++) {
          const uint32_t g = lane + 32u * k;
          const bool exists = w.atom + g * gs < natoms;
          const uint32_t flag = wide_flag(r, w, large_bits, len, g);
          if (exists) valid[k] |= 1u << lane;
          if (exists && flag != 0u) set[k] |= 1u << lane;
        }
      }
      const uint32_t n = wide_count(set, valid);
      for (uint32_t lane = 0; lane < 32; lane++) wide_lane_emit(lane, w, len, gs, n, table);
      ok = wide_advance(w, len, gs, n, natoms);
      wide = n == 32u * kWide;
      continue;
    }
    LaneView v[32];
    uint32_t set = 0, vmask = 0;
    for (uint32_t lane = 0; lane < 32; lane++) {
      v[lane] = walk_lane(r, w, large_bits, natoms, lane);
      const bool valid = v[lane].atom < natoms;
      if (valid) vmask |= 1u << lane;
      if (valid && v[lane].flag != 0u) set |= 1u << lane;
    }
    const uint32_t nv = (uint32_t)__builtin_popcount(vmask);
    const uint32_t winner = walk_winner(set, nv);
    for (uint32_t lane = 0; lane < 32; lane++) walk_lane_emit(lane, v[lane], w, winner, table);
    wide = set == 0u && nv == 32u;
    ok = walk_advance(w, winner, v[winner].npos, v[winner].natom, v[winner].nstate);
  }
  if (ok && w.pos > 8u * h.nbytes) ok = false;
  if (!ok) raise_flag(err, kDecBadXtcBit);
  (void)n_ck;
  return steps;
}

// xtc_index_scalar_kernel: one lane per frame
int64_t index_frame_scalar(const char* text, int64_t text_bytes, int64_t off, int n_atoms, U4* table, int* err) {
  FrameHdr h;
  const uint32_t* hw = frame_block(text, text_bytes, off, n_atoms, h);
  if (hw == nullptr) {
    raise_flag(err, kDecBadXtcBit);
    return 0;
  }
  const Stream s{hw + 10, 0u, (h.nbytes + 3u) >> 2, true};
  Cursor c{0u, 0u, h.smallidx, 0};
  uint32_t prev_atom = kNoAtom;
  const uint32_t natoms = (uint32_t)n_atoms;
  bool ok = true;
  int64_t steps = 0;
  while (ok && c.atom < natoms) {
    const Probe p = probe_group(s, c, h.large_bits, 0u);
    const IndexStep st = plan_index_step(p.flag, 1);
    index_lane_emit(0u, p, c, prev_atom, st, table);
    ok = index_advance(c, prev_atom, st, 1, p, h.large_bits, natoms);
    steps++;
  }
  if (!ok || c.pos > 8u * h.nbytes) raise_flag(err, kDecBadXtcBit);
  return steps;
}

}  // namespace

extern "C" {

// The whole of launch_xtc_decode on the host.  xyz8: float [n_frames][n_pad][3] in blocks of 8 (pre-filled by the caller
// with a sentinel, so that the test sees which words were written); sites: float [n_frames][sites_stride][4].
// stats (may be null): [0] steps of the index walk over all frames, [1] CTAs whose stream span was staged,
// [2] CTAs that read the stream from "global memory", [3] checkpoints written.
int emu_xtc_decode(const char* text, int64_t text_bytes, const int64_t* coords_off, int n_frames, int n_atoms, int n_pad,
                   const int32_t* slot, const float* slot_dq, int sites_stride, float* xyz8, float* sites, int scalar_index,
                   int max_stream_words /* 0: kXtcStreamWords; smaller values force the unstaged path */, int64_t* stats) {
  int err = 0;
  const int n_ck = (n_atoms + kCkAtoms - 1) / kCkAtoms;
  std::vector<U4> ck((size_t)n_ck * n_frames);
  memset(ck.data(), 0xFF, ck.size() * sizeof(U4));
  int64_t steps = 0;
  for (int f = 0; f < n_frames; f++) {
    U4* table = ck.data() + (size_t)f * n_ck;
    steps += scalar_index ? index_frame_scalar(text, text_bytes, coords_off[f], n_atoms, table, &err)
                          : index_frame_warp(text, text_bytes, coords_off[f], n_atoms, n_ck, table, &err);
  }
  int64_t written = 0;
  for (const U4& r : ck) written += r.y != kNoAtom;
  DecodeArgs a;
  a.text = text; a.text_bytes = text_bytes; a.coords_off = coords_off;
  a.n_atoms = n_atoms; a.n_pad = n_pad; a.n_ck = n_ck;
  a.ck = ck.data(); a.slot = slot; a.slot_dq = slot_dq; a.sites_stride = sites_stride;
  a.stream_words = max_stream_words > 0 ? max_stream_words : kXtcStreamWords;   // smaller values force the unstaged path
  a.xyz8 = xyz8; a.sites = reinterpret_cast<F4*>(sites); a.err_flags = &err;
  std::vector<float> stage((size_t)3 * kXtcStagePitch);
  std::vector<uint32_t> sw((size_t)a.stream_words);
  const uint32_t ctas = (uint32_t)((n_atoms + kXtcWindow - 1) / kXtcWindow);
  int64_t n_staged = 0, n_global = 0;
  for (int f = 0; f < n_frames; f++) {
    for (uint32_t b = 0; b < ctas; b++) {
      // fresh "shared memory" per CTA, filled with a pattern: a read of a word nobody wrote shows up in the results
      for (auto& v : stage) v = -7.0e33f;
      for (auto& v : sw) v = 0xDEADBEEFu;
      DecodeCta cta;
      cta.stage = stage.data();
      cta.sw = sw.data();
      MyCheckpoint mine[kXtcThreads];
      for (uint32_t tid = 0; tid < kXtcThreads; tid++) mine[tid] = decode_cta_my_checkpoint(a, f, b, tid);
      decode_cta_setup(a, cta, f, b);                                  // thread 0; __syncthreads
      if (cta.a_first != kNoAtom) (cta.staged ? n_staged : n_global)++;
      for (uint32_t tid = 0; tid < kXtcThreads; tid++) decode_cta_stage_stream(cta, tid);          // __syncthreads
      for (uint32_t tid = 0; tid < kXtcThreads; tid++) decode_cta_unpack(a, cta, b, tid, mine[tid]);   // __syncthreads
      for (uint32_t tid = 0; tid < kXtcThreads; tid++) decode_cta_write_back(a, cta, f, b, b == ctas - 1, tid);
    }
  }
  if (stats) {
    stats[0] = steps;
    stats[1] = n_staged;
    stats[2] = n_global;
    stats[3] = written;
  }
  return err;
}

// pieces on their own, for the property tests
uint32_t emu_value_of_stream_bits(uint32_t y, int rem) { return value_of_stream_bits(y, rem); }
void emu_divmod64(uint64_t n, uint32_t d, uint64_t* q, uint32_t* r) { *q = divmod64(n, d, ~(uint64_t)0 / d, *r); }
void emu_divmod32(uint32_t n, uint32_t d, uint32_t* q, uint32_t* r) { *q = divmod32(n, d, 0xFFFFFFFFu / d, *r); }
void emu_divmod96(uint64_t* lo, uint32_t* hi, uint32_t d, uint32_t* r) { divmod96(*lo, *hi, d, *r); }
float emu_to_coord(int k, float precision) { return to_coord(k, precision); }
int emu_header(const uint32_t* raw, int* out /* [ok, bitsize, large_bits, b0, b1, b2, smallidx] */) {
  const FrameHdr h = parse_header(raw);
  out[0] = h.ok; out[1] = h.bitsize; out[2] = h.large_bits; out[3] = h.bits3[0]; out[4] = h.bits3[1]; out[5] = h.bits3[2];
  out[6] = h.smallidx;
  return h.ok;
}

}  // extern "C"

#ifdef XTC_EMU_MAIN
// Stand-alone run for the sanitizer build (tests/test_xtc.py): xtcemu_asan file.xtc n_fuzz
// Every buffer is allocated at its exact size, so that the address sanitizer sees any read or write beside it: the clean
// file in every mode first, then n_fuzz copies with a few flipped bits (results ignored: only the accesses matter).
#include <cstdio>
#include <random>

static uint32_t be32_at(const std::vector<char>& b, size_t p) {
  return ((uint32_t)(unsigned char)b[p] << 24) | ((uint32_t)(unsigned char)b[p + 1] << 16) | ((uint32_t)(unsigned char)b[p + 2] << 8) |
         (uint32_t)(unsigned char)b[p + 3];
}

static int run_once(const std::vector<char>& blob, const std::vector<int64_t>& offs, int n_atoms, int scalar, int max_words,
                    std::vector<float>* xyz_out) {
  const int n_pad = (n_atoms + 2047) / 2048 * 2048;
  const int nf = (int)offs.size();
  const int stride = 41;
  char* text = (char*)malloc(blob.size());
  memcpy(text, blob.data(), blob.size());
  int32_t* slot = (int32_t*)malloc(sizeof(int32_t) * (size_t)n_pad);
  for (int i = 0; i < n_pad; i++) slot[i] = (i < n_atoms && i % 97 == 3 && i / 97 < stride) ? i / 97 : -1;
  float* slot_dq = (float*)malloc(sizeof(float) * stride);
  for (int i = 0; i < stride; i++) slot_dq[i] = 0.01f * i;
  float* xyz8 = (float*)malloc(sizeof(float) * 3 * (size_t)n_pad * nf);
  float* sites = (float*)malloc(sizeof(float) * 4 * (size_t)stride * nf);
  int64_t* coff = (int64_t*)malloc(sizeof(int64_t) * (size_t)nf);
  for (int i = 0; i < nf; i++) coff[i] = offs[(size_t)i];
  const int err = emu_xtc_decode(text, (int64_t)blob.size(), coff, nf, n_atoms, n_pad, slot, slot_dq, stride, xyz8, sites, scalar,
                                 max_words, nullptr);
  if (xyz_out) xyz_out->assign(xyz8, xyz8 + 3 * (size_t)n_pad * nf);
  free(text); free(slot); free(slot_dq); free(xyz8); free(sites); free(coff);
  return err;
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  std::vector<char> blob;
  char buf[65536];
  size_t n;
  while ((n = fread(buf, 1, sizeof buf, f)) > 0) blob.insert(blob.end(), buf, buf + n);
  fclose(f);
  const int n_fuzz = atoi(argv[2]);
  std::vector<int64_t> offs;
  size_t pos = 0;
  int n_atoms = 0;
  while (pos + 92 <= blob.size()) {
    n_atoms = (int)be32_at(blob, pos + 4);
    const uint32_t nbytes = be32_at(blob, pos + 88);
    offs.push_back((int64_t)pos + 52);
    pos += 92 + ((nbytes + 3) & ~3u);
  }
  if (offs.empty() || pos != blob.size()) return 3;
  std::vector<float> a, b;
  if (run_once(blob, offs, n_atoms, 0, 0, &a) != 0) return 4;
  if (run_once(blob, offs, n_atoms, 1, 0, &b) != 0 || a != b) return 5;
  if (run_once(blob, offs, n_atoms, 0, 48, &b) != 0 || a != b) return 6;
  printf("clean runs ok: %d frames of %d atoms\n", (int)offs.size(), n_atoms);
  std::mt19937_64 rng(12345);
  int flagged = 0;
  for (int it = 0; it < n_fuzz; it++) {
    std::vector<char> m = blob;
    const int flips = 1 + (int)(rng() % 4);
    for (int k = 0; k < flips; k++) {
      // mostly the stream, sometimes the ten header fields of a frame
      size_t p = (rng() % 5 == 0) ? (size_t)offs[rng() % offs.size()] + rng() % 40 : 52 + rng() % (m.size() - 52);
      m[p] = (char)(m[p] ^ (1 << (rng() % 8)));
    }
    flagged += run_once(m, offs, n_atoms, (int)(it & 1), (it % 3 == 0) ? 40 : 0, nullptr) != 0;
  }
  printf("fuzz: %d of %d damaged copies flagged\n", flagged, n_fuzz);
  return 0;
}
#endif

// K1x: GROMACS .xtc frames (compressed coordinates) -> float32 coordinates in the pair kernel's layout + site blocks, on
// the device (sm_100a).  Stream format, the shared arithmetic and the per-lane / per-thread pieces of both kernels:
// xtc_bits.h (the CPU tests run those same pieces lane by lane, tests/xtc_emulation.cpp).
//
// The bit stream of a frame is serial by construction: where a group of atoms starts depends on the run length and the
// small-number table index every group before it left behind.  Two kernels take it apart:
//
//   xtc_index_kernel   ONE WARP PER FRAME walks the group headers only and leaves a checkpoint (bit position, atom, table
//                      index, run length) for every 16 atoms.  The walk is speculative: a group with a clear flag bit
//                      changes neither run length nor table index, so 32 lanes look at the flag bits of the next 32 groups
//                      as they would lie if all of them were clear; a ballot finds the first set flag, every group before
//                      it is confirmed at once, the flagged one is taken serially.  Synthetic and solvent-heavy frames
//                      advance close to 32 groups per step, bonded runs one or two.
//   xtc_decode_kernel  one thread per checkpoint, 256 checkpoints (4096 atoms) per CTA.  The CTA first stages the part of
//                      the stream its checkpoints span into shared memory (coalesced, byte order fixed on the way), every
//                      thread then unpacks its ~16 atoms out of shared memory into a staging array there, and the CTA
//                      writes the coordinate blocks and site entries back coalesced.
//
// Outputs are exactly those of decode_kernel (kernels.cu): xyz8 blocks [x0..x7][y0..y7][z0..z7] with zeroed tile padding,
// and float4(-x, -y, -z, dQ) per dQ != 0 site atom.
#include "kernels.cuh"
#include "xtc_bits.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace cgc {

using namespace xtc;

static_assert(kDecBadXtcBit == kDecBadXtc, "flag bit of xtc_bits.h out of step with kernels.cuh");
static_assert(sizeof(U4) == sizeof(uint4) && sizeof(F4) == sizeof(float4), "record layouts");

__global__ void __launch_bounds__(32)
xtc_index_kernel(const char* __restrict__ text, int64_t text_bytes, const int64_t* __restrict__ coords_off, int n_atoms,
                 int n_ck, U4* __restrict__ ck, int* __restrict__ err_flags) {
  __shared__ uint32_t ring_words[kRingWords];
  const int f = blockIdx.x;
  const uint32_t lane = threadIdx.x;
  FrameHdr h;
  const uint32_t* hw = frame_block(text, text_bytes, coords_off[f], n_atoms, h);
  if (hw == nullptr) {
    if (lane == 0) raise_flag(err_flags, kDecBadXtcBit);
    return;
  }
  Ring r;
  ring_init(r, ring_words, hw + 10, (h.nbytes + 3u) >> 2);
  Walk w{0u, 0u, (uint32_t)h.smallidx, 0u, kNoAtom};
  U4* table = ck + (size_t)f * n_ck;
  const uint32_t natoms = (uint32_t)n_atoms, large_bits = (uint32_t)h.large_bits;
  bool ok = true, wide = false;
  while (ok && w.atom < natoms) {
    const uint32_t cw = w.pos >> 5;
    while (ring_can_issue(r, cw)) {
      ring_issue_lane(r, lane);
      ring_issued(r);
    }
    if (ring_must_wait(r, cw, wide ? kRingLookWide : kRingLook)) ring_wait(r);
    if (wide) {
      const uint32_t len = large_bits + 1u + w.nsm * w.smallidx, gs = 1u + w.nsm;
      uint32_t set[kWide], valid[kWide];
#pragma unroll
      for (uint32_t k = 0; k < kWide; k++) {
        const uint32_t g = lane + 32u * k;
        const bool exists = w.atom + g * gs < natoms;
        const uint32_t flag = wide_flag(r, w, large_bits, len, g);
        set[k] = __ballot_sync(0xFFFFFFFFu, exists && flag != 0u);
        valid[k] = __ballot_sync(0xFFFFFFFFu, exists);
      }
      const uint32_t n = wide_count(set, valid);
      wide_lane_emit(lane, w, len, gs, n, table);
      ok = wide_advance(w, len, gs, n, natoms);
      wide = n == 32u * kWide;
      continue;
    }
    const LaneView v = walk_lane(r, w, large_bits, natoms, lane);
    const bool valid = v.atom < natoms;
    const uint32_t set = __ballot_sync(0xFFFFFFFFu, valid && v.flag != 0u);
    const uint32_t nv = (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, valid));
    const uint32_t winner = walk_winner(set, nv);
    walk_lane_emit(lane, v, w, winner, table);
    wide = set == 0u && nv == 32u;   // 32 clear flags: look further next time
    ok = walk_advance(w, winner, __shfl_sync(0xFFFFFFFFu, v.npos, (int)winner), __shfl_sync(0xFFFFFFFFu, v.natom, (int)winner),
                      __shfl_sync(0xFFFFFFFFu, v.nstate, (int)winner));
  }
  asm volatile("cp.async.wait_all;" ::: "memory");   // nothing may still be landing in shared memory when the CTA ends
  if (ok && w.pos > 8u * h.nbytes) ok = false;
  if (!ok && lane == 0) raise_flag(err_flags, kDecBadXtcBit);
}

// One lane per frame, no speculation: the plain serial walk (CGC_XTC_SCALAR_INDEX=1; the cross-check of the speculative
// walk and the comparison in DESIGN.md).  It is the same step with a warp of one.
__global__ void __launch_bounds__(32)
xtc_index_scalar_kernel(const char* __restrict__ text, int64_t text_bytes, const int64_t* __restrict__ coords_off, int n_atoms,
                        int n_ck, int n_frames, U4* __restrict__ ck, int* __restrict__ err_flags) {
  const int f = blockIdx.x * 32 + threadIdx.x;
  if (f >= n_frames) return;
  FrameHdr h;
  const uint32_t* hw = frame_block(text, text_bytes, coords_off[f], n_atoms, h);
  if (hw == nullptr) {
    raise_flag(err_flags, kDecBadXtcBit);
    return;
  }
  const Stream s{hw + 10, 0u, (h.nbytes + 3u) >> 2, true};
  Cursor c{0u, 0u, h.smallidx, 0};
  uint32_t prev_atom = kNoAtom;
  U4* table = ck + (size_t)f * n_ck;
  const uint32_t natoms = (uint32_t)n_atoms;
  bool ok = true;
  while (ok && c.atom < natoms) {
    const Probe p = probe_group(s, c, h.large_bits, 0u);
    const IndexStep st = plan_index_step(p.flag, 1);
    index_lane_emit(0u, p, c, prev_atom, st, table);
    ok = index_advance(c, prev_atom, st, 1, p, h.large_bits, natoms);
  }
  if (!ok || c.pos > 8u * h.nbytes) raise_flag(err_flags, kDecBadXtcBit);
}

__global__ void __launch_bounds__(kXtcThreads, 3)
xtc_decode_kernel(DecodeArgs a) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  __shared__ DecodeCta cta;
  const int f = blockIdx.y;
  const uint32_t tid = threadIdx.x;
  const MyCheckpoint mine = decode_cta_my_checkpoint(a, f, blockIdx.x, tid);
  if (tid == 0) {
    cta.stage = reinterpret_cast<float*>(sm_raw);
    cta.sw = reinterpret_cast<uint32_t*>(sm_raw + sizeof(float) * 3 * kXtcStagePitch);
    decode_cta_setup(a, cta, f, blockIdx.x);
  }
  __syncthreads();
  decode_cta_stage_stream(cta, tid);
  __syncthreads();
  decode_cta_unpack(a, cta, blockIdx.x, tid, mine);
  __syncthreads();
  decode_cta_write_back(a, cta, f, blockIdx.x, blockIdx.x == gridDim.x - 1, tid);
}

size_t xtc_scratch_bytes(const DeviceSystem& sys, int n_frames) {
  const size_t n_ck = ((size_t)sys.n_atoms + kCkAtoms - 1) / kCkAtoms;
  return sizeof(U4) * n_ck * (size_t)n_frames;
}

void launch_xtc_index(const DeviceSystem& sys, const char* text, int64_t text_bytes, const int64_t* coords_off, int n_frames,
                      void* scratch, int* err_flags, cudaStream_t st) {
  if (n_frames <= 0 || sys.n_atoms <= 0) return;
  const int n_ck = (sys.n_atoms + kCkAtoms - 1) / kCkAtoms;
  U4* ck = reinterpret_cast<U4*>(scratch);
  cudaMemsetAsync(ck, 0xFF, xtc_scratch_bytes(sys, n_frames), st);   // every checkpoint "none" (atom = kNoAtom)
  const char* e = getenv("CGC_XTC_SCALAR_INDEX");
  if (e && *e && *e != '0') {
    xtc_index_scalar_kernel<<<(unsigned)((n_frames + 31) / 32), 32, 0, st>>>(text, text_bytes, coords_off, sys.n_atoms, n_ck, n_frames,
                                                                              ck, err_flags);
  } else {
    xtc_index_kernel<<<(unsigned)n_frames, 32, 0, st>>>(text, text_bytes, coords_off, sys.n_atoms, n_ck, ck, err_flags);
  }
}

void launch_xtc_unpack(const DeviceSystem& sys, const char* text, int64_t text_bytes, const int64_t* coords_off, int n_frames,
                       const void* scratch, float* xyz8, float4* sites, int* err_flags, cudaStream_t st) {
  if (n_frames <= 0 || sys.n_atoms <= 0) return;
  const int n_ck = (sys.n_atoms + kCkAtoms - 1) / kCkAtoms;
  // Stream words a 4096-atom window spans on average in this batch, with a twelfth to spare (a window of isolated atoms in
  // a frame that also holds bonded runs is longer than the average): 23 KB hold them for boxes up to ~20 nm at precision
  // 1000 (three CTAs per SM); frames that need more get up to 48 KB (two CTAs per SM); a window that still does not fit
  // reads its stream from global memory.
  const int64_t per_window = text_bytes / n_frames * kXtcWindow / std::max(1, sys.n_atoms) / 4;
  const int64_t want_words = per_window + per_window / 12 + 64;
  int stream_words = kXtcStreamWords;
  if (want_words > kXtcStreamWords) stream_words = (int)std::min<int64_t>(kXtcStreamWordsMax, (want_words + 255) / 256 * 256);
  const size_t smem = sizeof(float) * 3 * kXtcStagePitch + sizeof(uint32_t) * (size_t)stream_words;
  static bool granted[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!granted[dev & 15]) {
    cudaFuncSetAttribute(xtc_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(sizeof(float) * 3 * kXtcStagePitch + sizeof(uint32_t) * kXtcStreamWordsMax));
    granted[dev & 15] = true;
  }
  DecodeArgs a;
  a.text = text; a.text_bytes = text_bytes; a.coords_off = coords_off;
  a.n_atoms = sys.n_atoms; a.n_pad = sys.n_pad; a.n_ck = n_ck;
  a.ck = reinterpret_cast<const U4*>(scratch); a.slot = sys.slot; a.slot_dq = sys.slot_dq; a.sites_stride = sys.n_sites * sys.site_cap;
  a.stream_words = stream_words;
  a.xyz8 = xyz8; a.sites = reinterpret_cast<F4*>(sites); a.err_flags = err_flags;
  const dim3 grid((unsigned)((sys.n_atoms + kXtcWindow - 1) / kXtcWindow), (unsigned)n_frames);
  xtc_decode_kernel<<<grid, kXtcThreads, smem, st>>>(a);
}

void launch_xtc_decode(const DeviceSystem& sys, const char* text, int64_t text_bytes, const int64_t* coords_off, int n_frames,
                       void* scratch, float* xyz8, float4* sites, int* err_flags, cudaStream_t st) {
  launch_xtc_index(sys, text, text_bytes, coords_off, n_frames, scratch, err_flags, st);
  launch_xtc_unpack(sys, text, text_bytes, coords_off, n_frames, scratch, xyz8, sites, err_flags, st);
}

}  // namespace cgc

// GROMACS .xtc compressed coordinates: the arithmetic shared by the device kernels (xtc_kernel.cu), the host-side frame
// index and encoder (xtc_host.cpp) and the lane-by-lane emulation the CPU tests run (tests/xtc_emulation.cpp).
//
// Not a reference format: the authors' pipeline converts binary trajectories to .gro text with trjconv before traj2dEcsv
// (reference scripts/2014-09-trr2dEcsv.sh:22); SURVEY.md section 8f-3 asks for a binary trr/xtc reader to remove the
// 45-69 bytes of text per atom and frame.  The stream layout is that of GROMACS' xdr3dfcoord
// (src/gromacs/fileio/libxdrf.cpp, published also as xdrfile.c), restated here from its algorithm:
//   * bits are consumed most significant first;
//   * a packed triple is the number (a * s1 + b) * s2 + c written least significant BYTE first into nbits bits (the last
//     byte carries the 1..8 bits that are left);
//   * every atom or run starts with a "large" triple relative to minint in bitlength(s0*s1*s2) bits — or three separate
//     fields of bitlength(s_d) bits when a size exceeds 2^24 - 1 — then one flag bit; flag set: 5 bits r with
//     is_smaller = r % 3 - 1 and run = r - r % 3; flag clear: the run length of the previous group stays;
//   * run / 3 "small" triples of `smallidx` bits follow, each relative to the previous atom and offset by
//     magicints[smallidx] / 2, sizes magicints[smallidx]; the first small atom and the large one change places on output;
//   * after the group smallidx += is_smaller.
// A decoded integer k becomes the float32 nearest to k / precision: the value (float)atof yields on the "%8.3f" text
// trjconv writes for the same frame (reference src/grom2atomFMO.cpp:448-450), so a precision-1000 .xtc and its .gro twin
// give bit-identical coordinates.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define XTC_HD __host__ __device__ __forceinline__
#else
#define XTC_HD inline
#endif

namespace cgc {
namespace xtc {

constexpr int kFirstIdx = 9;        // first usable entry of the magic-number table
constexpr int kLastIdx = 72;        // last entry
constexpr int kMaxRunAtoms = 8;     // small atoms per group at most (run <= 24)
constexpr int kCkAtoms = 16;        // the index pass leaves one checkpoint per 16 atoms
constexpr uint32_t kNoAtom = 0xFFFFFFFFu;

#define CGC_XTC_MAGICINTS                                                                                                  \
  0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512, 645, 812,   \
      1024, 1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384, 20642, 26007, 32768, 41285, 52015,  \
      65536, 82570, 104031, 131072, 165140, 208063, 262144, 330280, 416127, 524287, 660561, 832255, 1048576, 1321122,      \
      1664510, 2097152, 2642245, 3329021, 4194304, 5284491, 6658042, 8388607, 10568983, 13316085, 16777216

static const uint32_t kMagicHost[kLastIdx + 1] = {CGC_XTC_MAGICINTS};
#if defined(__CUDACC__)
static __device__ __constant__ uint32_t kMagicDev[kLastIdx + 1] = {CGC_XTC_MAGICINTS};
#endif

XTC_HD uint32_t magic(int i) {
#if defined(__CUDA_ARCH__)
  return kMagicDev[i];
#else
  return kMagicHost[i];
#endif
}

XTC_HD uint32_t bswap32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __byte_perm(v, 0, 0x0123);
#else
  return __builtin_bswap32(v);
#endif
}
XTC_HD int bitlen32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return 32 - __clz((int)v);
#else
  return v ? 32 - __builtin_clz(v) : 0;
#endif
}
XTC_HD int bitlen64(uint64_t v) {
#if defined(__CUDA_ARCH__)
  return 64 - __clzll((long long)v);
#else
  return v ? 64 - __builtin_clzll(v) : 0;
#endif
}
XTC_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(a, b);
#else
  return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}
XTC_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

// The bit stream of one frame (or of the part of it a CTA staged into shared memory) as 32-bit words.  `swap`: the words
// are the file's bytes as loaded (big-endian content in a little-endian register); otherwise they are already in value
// order.  Word indices are clamped to [lo, hi): a speculative or damaged position reads a wrong value, never another
// buffer.
struct Stream {
  const uint32_t* w;    // w[0] is word `lo` of the frame's stream
  uint32_t lo, hi;      // valid word range, hi > lo
  bool swap;
};
XTC_HD uint32_t stream_word(const Stream& s, uint32_t idx) {
  uint32_t at = idx - s.lo;                       // a word below lo wraps to a huge offset: one unsigned minimum covers both ends
  const uint32_t last = s.hi - s.lo - 1u;
  at = at < last ? at : last;
  const uint32_t v = s.w[at];
  return s.swap ? bswap32(v) : v;
}
// the 32 bits that start at bit position p
XTC_HD uint32_t peek32(const Stream& s, uint32_t p) {
  const uint32_t idx = p >> 5, sh = p & 31u;
  const uint32_t a = stream_word(s, idx), b = stream_word(s, idx + 1);
#if defined(__CUDA_ARCH__)
  return __funnelshift_l(b, a, sh);
#else
  return sh ? (a << sh) | (b >> (32 - sh)) : a;
#endif
}
// n bits (1..32) starting at p, right-aligned
XTC_HD uint32_t read_bits(const Stream& s, uint32_t p, int n) { return peek32(s, p) >> (32 - n); }

// `rem` (1..32) stream bits, left-aligned in y, hold the bytes of a packed number least significant first, 8 bits at a
// time, the last byte with the 1..8 bits that are left: the value they spell.
XTC_HD uint32_t value_of_stream_bits(uint32_t y, int rem) {
  const int nfull = (rem - 1) >> 3;        // whole bytes before the last one (0..3)
  const int last = rem - 8 * nfull;        // bits of the last byte, 1..8
  const uint32_t v = bswap32(y);           // byte j of the stream at bits 8j..8j+7; the last one left-aligned in its byte
  const uint32_t keep = nfull ? (0xFFFFFFFFu >> (32 - 8 * nfull)) : 0u;
  return (v & keep) | ((((v >> (8 * nfull)) & 0xFFu) >> (8 - last)) << (8 * nfull));
}
// A packed number of nbits (1..64) bits at p: two chunks of up to 32 stream bits.
XTC_HD uint64_t read_packed64(const Stream& s, uint32_t p, int nbits) {
  if (nbits <= 32) {
    const uint32_t y = peek32(s, p);
    return value_of_stream_bits(nbits == 32 ? y : (y & (0xFFFFFFFFu << (32 - nbits))), nbits);
  }
  const uint32_t w0 = bswap32(peek32(s, p));
  const int rem = nbits - 32;
  const uint32_t y = peek32(s, p + 32u);
  const uint32_t w1 = value_of_stream_bits(rem == 32 ? y : (y & (0xFFFFFFFFu << (32 - rem))), rem);
  return (uint64_t)w0 | ((uint64_t)w1 << 32);
}
// A packed number of nbits (1..96) bits at p.
XTC_HD void read_packed(const Stream& s, uint32_t p, int nbits, uint64_t& lo, uint32_t& hi) {
  if (nbits <= 64) {
    lo = read_packed64(s, p, nbits);
    hi = 0u;
    return;
  }
  uint32_t w[3] = {0u, 0u, 0u};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int c = 0; c < 3; c++) {
    const int rem = nbits - 32 * c;
    if (rem >= 32) {
      w[c] = bswap32(peek32(s, p + 32u * (uint32_t)c));
    } else if (rem > 0) {
      w[c] = value_of_stream_bits(peek32(s, p + 32u * (uint32_t)c) & (0xFFFFFFFFu << (32 - rem)), rem);
    }
  }
  lo = (uint64_t)w[0] | ((uint64_t)w[1] << 32);
  hi = w[2];
}

// n / d and n % d for d < 2^24 from m = floor((2^64 - 1) / d): floor(n * m / 2^64) is the quotient or one less.
XTC_HD uint64_t divmod64(uint64_t n, uint32_t d, uint64_t m, uint32_t& r) {
  uint64_t q = mulhi64(n, m);
  uint64_t rr = n - q * d;
  if (rr >= d) { q++; rr -= d; }
  r = (uint32_t)rr;
  return q;
}
XTC_HD uint32_t divmod32(uint32_t n, uint32_t d, uint32_t m, uint32_t& r) {   // m = floor((2^32 - 1) / d)
  uint32_t q = mulhi32(n, m);
  uint32_t rr = n - q * d;
  if (rr >= d) { q++; rr -= d; }
  r = rr;
  return q;
}
// (hi:lo) / d for a 96-bit numerator, three 32-bit limbs, schoolbook (only streams whose packed numbers exceed 64 bits)
XTC_HD void divmod96(uint64_t& lo, uint32_t& hi, uint32_t d, uint32_t& r) {
  uint64_t t = hi;
  const uint32_t q2 = (uint32_t)(t / d);
  t = ((t % d) << 32) | (lo >> 32);
  const uint32_t q1 = (uint32_t)(t / d);
  t = ((t % d) << 32) | (lo & 0xFFFFFFFFu);
  const uint32_t q0 = (uint32_t)(t / d);
  r = (uint32_t)(t % d);
  hi = q2;
  lo = ((uint64_t)q1 << 32) | q0;
}

// What the header of a frame's coordinate block says (ten 32-bit fields), plus what follows from it.
struct FrameHdr {
  int natoms;
  float precision;
  int minint[3];
  uint32_t sizeint[3];
  int bitsize;          // bits of a large triple packed as one number; 0: three separate fields
  int bits3[3];         // widths of the separate fields
  int large_bits;       // bits a large triple takes either way
  int smallidx;
  uint32_t nbytes;      // bytes of bit stream
  uint64_t m1, m2;      // floor((2^64 - 1) / sizeint[1]), .. / sizeint[2]
  int ok;               // 0: impossible values
};
XTC_HD FrameHdr parse_header(const uint32_t* raw /* the ten fields as loaded from the file */) {
  FrameHdr h;
  h.natoms = (int)bswap32(raw[0]);
  union { uint32_t u; float f; } cv;
  cv.u = bswap32(raw[1]);
  h.precision = cv.f;
  int maxint[3];
  for (int d = 0; d < 3; d++) {
    h.minint[d] = (int)bswap32(raw[2 + d]);
    maxint[d] = (int)bswap32(raw[5 + d]);
  }
  h.smallidx = (int)bswap32(raw[8]);
  h.nbytes = bswap32(raw[9]);
  h.ok = h.natoms > 9 && h.smallidx >= kFirstIdx && h.smallidx <= kLastIdx && h.nbytes > 0 && h.nbytes < (1u << 28) &&
         h.precision > 0.f;
  for (int d = 0; d < 3; d++) {
    const int64_t sz = (int64_t)maxint[d] - (int64_t)h.minint[d] + 1;
    if (sz <= 0 || sz > 0xFFFFFFFFll) h.ok = 0;
    h.sizeint[d] = (uint32_t)(sz > 0 ? sz : 1);
  }
  if ((h.sizeint[0] | h.sizeint[1] | h.sizeint[2]) > 0xFFFFFFu) {
    h.bitsize = 0;
    for (int d = 0; d < 3; d++) h.bits3[d] = bitlen32(h.sizeint[d]);
    h.large_bits = h.bits3[0] + h.bits3[1] + h.bits3[2];
  } else {
    const uint64_t p01 = (uint64_t)h.sizeint[0] * h.sizeint[1];
    const uint64_t lo = p01 * h.sizeint[2], hi = mulhi64(p01, h.sizeint[2]);
    h.bitsize = hi ? 64 + bitlen64(hi) : bitlen64(lo);
    h.bits3[0] = h.bits3[1] = h.bits3[2] = 0;
    h.large_bits = h.bitsize;
  }
  h.m1 = ~(uint64_t)0 / h.sizeint[1];
  h.m2 = ~(uint64_t)0 / h.sizeint[2];
  return h;
}

// k / precision rounded once to float32.  |k| < 2^24: the float division of two exactly represented operands is that
// value; the two-FMA form for precision 1000 equals it for every such k (the .gro path's div1000, checked exhaustively,
// oracle/cdc_oracle.c cgo_div1000_*).  Larger |k| go through double.
XTC_HD float to_coord(int k, float precision) {
  const uint32_t a = k < 0 ? (uint32_t)(-(int64_t)k) : (uint32_t)k;
  float r;
  if (a < (1u << 24)) {
#if defined(__CUDA_ARCH__)
    if (precision == 1000.0f) {
      const float n = (float)a, rc = 1.0f / 1000.0f;
      const float q = __fmul_rn(n, rc);
      r = __fmaf_rn(__fmaf_rn(-q, 1000.0f, n), rc, q);
    } else {
      r = __fdiv_rn((float)a, precision);
    }
#else
    r = (float)a / precision;
#endif
  } else {
    r = (float)((double)a / (double)precision);
  }
  return k < 0 ? -r : r;
}

// The state the stream is in before a group: where its large triple starts, which atom it is, the table index the small
// triples use and the run length a clear flag bit would keep.
struct Cursor {
  uint32_t pos, atom;
  int smallidx, run;
};
// checkpoint record: x = pos, y = atom (kNoAtom: none), z = smallidx | run << 8
XTC_HD void pack_cursor(const Cursor& c, uint32_t out[4]) {
  out[0] = c.pos;
  out[1] = c.atom;
  out[2] = (uint32_t)c.smallidx | ((uint32_t)c.run << 8);
  out[3] = 0;
}
XTC_HD Cursor unpack_cursor(uint32_t pos, uint32_t atom, uint32_t z) {
  Cursor c;
  c.pos = pos;
  c.atom = atom;
  c.smallidx = (int)(z & 0xFFu);
  c.run = (int)((z >> 8) & 0xFFu);
  return c;
}

// ---- index pass ------------------------------------------------------------------------------------------------------
// What one lane sees when it assumes that the l groups before its own all had a clear flag bit (so that each of them
// took `len` bits and `gs` atoms): its group's flag bit and the five bits behind it.
struct Probe {
  uint32_t pos;      // where the lane's group starts
  uint32_t atom;     // its first atom
  uint32_t flag, r5;
};
XTC_HD Probe probe_group(const Stream& s, const Cursor& c, int large_bits, uint32_t lane) {
  const uint32_t nsm = (uint32_t)c.run / 3u;
  const uint32_t len = (uint32_t)large_bits + 1u + nsm * (uint32_t)c.smallidx, gs = 1u + nsm;
  Probe p;
  p.pos = c.pos + lane * len;
  p.atom = c.atom + lane * gs;
  const uint32_t v = peek32(s, p.pos + (uint32_t)large_bits);
  p.flag = v >> 31;
  p.r5 = (v >> 26) & 31u;
  return p;
}
// does the group that starts at `atom` (the one before it started at prev_atom, kNoAtom: there is none) open a new
// 16-atom window — i.e. is it the checkpoint of window atom / kCkAtoms?
XTC_HD bool opens_window(uint32_t atom, uint32_t prev_atom) {
  return prev_atom == kNoAtom || (prev_atom / kCkAtoms) < (atom / kCkAtoms);
}
// The cursor after a group with a SET flag whose five bits are r5 (the group itself still uses c.smallidx).  Returns
// false when the bits describe an impossible group.
XTC_HD bool after_flagged(const Cursor& at, uint32_t r5, int large_bits, uint32_t natoms, Cursor& next) {
  const int is_smaller = (int)(r5 % 3u) - 1;
  const int run = (int)(r5 - r5 % 3u);
  const uint32_t nsm = (uint32_t)run / 3u;
  next.pos = at.pos + (uint32_t)large_bits + 6u + nsm * (uint32_t)at.smallidx;
  next.atom = at.atom + 1u + nsm;
  next.smallidx = at.smallidx + is_smaller;
  next.run = run;
  return run <= 3 * kMaxRunAtoms && next.atom <= natoms && next.smallidx >= kFirstIdx && next.smallidx <= kLastIdx;
}

// ---- decode pass -----------------------------------------------------------------------------------------------------
// Unpack three numbers below s0 (unchecked), s1, s2 from a packed value.  m1/m2: reciprocals for divmod64; used while the
// value fits 64 bits, which every stream of a box below 2^21 units per side does.
XTC_HD void unpack3(uint64_t lo, uint32_t hi, uint32_t s1, uint32_t s2, uint64_t m1, uint64_t m2, uint32_t out[3]) {
  if (hi != 0u) {
    divmod96(lo, hi, s2, out[2]);
    divmod96(lo, hi, s1, out[1]);
    out[0] = (uint32_t)lo;
    return;
  }
  uint64_t q = divmod64(lo, s2, m2, out[2]);
  q = divmod64(q, s1, m1, out[1]);
  out[0] = (uint32_t)q;
}

// Decode the groups from cursor `c` up to (not including) the group that starts at atom `end_atom`; `limit` bounds the
// atoms that may be emitted at all (a damaged table never makes a thread write outside its window).  emit(atom, k[3])
// receives every atom's integer coordinates in file order.  Returns false when the stream turned out to be impossible.
template <typename Emit>
XTC_HD bool decode_groups(const Stream& s, const FrameHdr& h, Cursor c, uint32_t end_atom, uint32_t limit, Emit&& emit) {
  uint32_t small_size = 0, small_m32 = 0;
  uint64_t small_m64 = 0;
  int cached_idx = -1;
  while (c.atom < end_atom) {
    uint32_t v[3];
    if (h.bitsize) {
      uint64_t lo;
      uint32_t hi;
      read_packed(s, c.pos, h.bitsize, lo, hi);
      unpack3(lo, hi, h.sizeint[1], h.sizeint[2], h.m1, h.m2, v);
    } else {
      uint32_t p = c.pos;
      for (int d = 0; d < 3; d++) {
        v[d] = h.bits3[d] ? read_bits(s, p, h.bits3[d]) : 0u;
        p += (uint32_t)h.bits3[d];
      }
    }
    uint32_t p = c.pos + (uint32_t)h.large_bits;
    int large[3];
    for (int d = 0; d < 3; d++) large[d] = (int)(v[d] + (uint32_t)h.minint[d]);
    const uint32_t fb = peek32(s, p);
    p += 1u;
    int is_smaller = 0;
    if (fb >> 31) {
      const uint32_t r5 = (fb >> 26) & 31u;
      is_smaller = (int)(r5 % 3u) - 1;
      c.run = (int)(r5 - r5 % 3u);
      p += 5u;
    }
    const uint32_t nsm = (uint32_t)c.run / 3u;
    if (nsm > (uint32_t)kMaxRunAtoms || c.atom + 1u + nsm > limit) return false;
    if (nsm == 0u) {
      emit(c.atom, large);
    } else {
      if (cached_idx != c.smallidx) {
        cached_idx = c.smallidx;
        small_size = magic(c.smallidx);
        small_m32 = 0xFFFFFFFFu / small_size;
        small_m64 = c.smallidx > 32 ? ~(uint64_t)0 / small_size : 0;
      }
      const int half = (int)(small_size >> 1);
      int prev[3] = {large[0], large[1], large[2]};
      for (uint32_t k = 0; k < nsm; k++) {
        uint32_t d3[3];
        if (c.smallidx <= 32) {
          uint32_t q = (uint32_t)read_packed64(s, p, c.smallidx);
          q = divmod32(q, small_size, small_m32, d3[2]);
          d3[0] = divmod32(q, small_size, small_m32, d3[1]);
        } else {
          uint64_t lo;
          uint32_t hi;
          read_packed(s, p, c.smallidx, lo, hi);
          unpack3(lo, hi, small_size, small_size, small_m64, small_m64, d3);
        }
        p += (uint32_t)c.smallidx;
        int cur[3];
        for (int d = 0; d < 3; d++) cur[d] = (int)d3[d] + prev[d] - half;
        if (k == 0u) {
          emit(c.atom, cur);          // the first small atom comes out before the large one
          emit(c.atom + 1u, large);
        } else {
          emit(c.atom + 1u + k, cur);
        }
        for (int d = 0; d < 3; d++) prev[d] = cur[d];
      }
    }
    c.atom += 1u + nsm;
    c.pos = p;
    c.smallidx += is_smaller;
    if (c.smallidx < kFirstIdx || c.smallidx > kLastIdx) return c.atom >= end_atom;
  }
  return true;
}


// ======================================================================================================================
// The two kernels of xtc_kernel.cu, written as per-lane / per-thread pieces.  The device kernels and the CPU emulation
// (tests/xtc_emulation.cpp, which runs the lanes of a warp and the threads of a CTA one after the other between the
// barriers) execute these same functions; only the collectives (ballot, shuffle, __syncthreads) differ.
// ======================================================================================================================
struct alignas(16) U4 { uint32_t x, y, z, w; };     // a checkpoint record (layout of CUDA's uint4)
struct alignas(16) F4 { float x, y, z, w; };        // a site entry (layout of CUDA's float4)

constexpr int kDecBadXtcBit = 8;        // == kDecBadXtc (kernels.cuh)
XTC_HD void raise_flag(int* flags, int bit) {
#if defined(__CUDA_ARCH__)
  atomicOr(flags, bit);
#else
  *flags |= bit;
#endif
}

// Locate and check the coordinate block of a frame inside the batch text: its ten header fields, then the stream.
// Returns null when the block cannot be used.
XTC_HD const uint32_t* frame_block(const char* text, int64_t text_bytes, int64_t off, int n_atoms, FrameHdr& h) {
  h.ok = 0;
  h.nbytes = 0;
  if (off < 0 || (off & 3) != 0 || off + 40 > text_bytes) return nullptr;
  const uint32_t* hw = reinterpret_cast<const uint32_t*>(text + off);
  uint32_t raw[10];
  for (int i = 0; i < 10; i++) raw[i] = hw[i];
  h = parse_header(raw);
  if (!h.ok || h.natoms != n_atoms || off + 40 + 4 * (int64_t)((h.nbytes + 3u) >> 2) > text_bytes) {
    h.ok = 0;
    return nullptr;
  }
  return hw;
}

// ---- index pass: one warp per frame ----------------------------------------------------------------------------------
// One step of the walk: every lane has probed its group (probe_group).  `set` = ballot of "my atom exists and my flag bit
// is set", nv = number of lanes whose atom exists (they are lanes 0 .. nv-1).
struct IndexStep {
  int fl;      // first lane whose flag is set, 32: none
  int last;    // the groups of lanes 0 .. last are consumed by this step
};
XTC_HD IndexStep plan_index_step(uint32_t set, int nv) {
  IndexStep st;
#if defined(__CUDA_ARCH__)
  st.fl = set ? __ffs((int)set) - 1 : 32;
#else
  st.fl = set ? __builtin_ctz(set) : 32;
#endif
  st.last = st.fl < nv ? st.fl : nv - 1;
  return st;
}
// what lane `lane` leaves behind in this step: the checkpoint of its group when the group opens a 16-atom window
XTC_HD void index_lane_emit(uint32_t lane, const Probe& p, const Cursor& c, uint32_t prev_atom, const IndexStep& st, U4* table) {
  if ((int)lane > st.last) return;
  const uint32_t gs = 1u + (uint32_t)c.run / 3u;
  const uint32_t before = lane == 0u ? prev_atom : p.atom - gs;
  if (opens_window(p.atom, before)) {
    Cursor at;
    at.pos = p.pos; at.atom = p.atom; at.smallidx = c.smallidx; at.run = c.run;
    uint32_t r[4];
    pack_cursor(at, r);
    U4 rec;
    rec.x = r[0]; rec.y = r[1]; rec.z = r[2]; rec.w = r[3];
    table[p.atom / kCkAtoms] = rec;
  }
}
// the warp-uniform cursor after the step; fp = the probe of lane st.fl (used only when st.fl < nv).  False: impossible stream.
XTC_HD bool index_advance(Cursor& c, uint32_t& prev_atom, const IndexStep& st, int nv, const Probe& fp, int large_bits, uint32_t natoms) {
  const uint32_t nsm = (uint32_t)c.run / 3u, gs = 1u + nsm;
  if (st.fl < nv) {
    Cursor at, next;
    at.pos = fp.pos; at.atom = fp.atom; at.smallidx = c.smallidx; at.run = c.run;
    if (!after_flagged(at, fp.r5, large_bits, natoms, next)) return false;
    prev_atom = at.atom;
    c = next;
    return true;
  }
  const uint32_t len = (uint32_t)large_bits + 1u + nsm * (uint32_t)c.smallidx;
  prev_atom = c.atom + (uint32_t)(nv - 1) * gs;
  c.pos += (uint32_t)nv * len;
  c.atom += (uint32_t)nv * gs;
  return c.atom <= natoms;
}

// ---- the index pass's window on the stream -----------------------------------------------------------------------
// The walk reads a few bytes per step, each step a little further on: from global memory every step would pay a cache-line
// miss.  The warp therefore keeps the stream in a ring of shared-memory words that it fills AHEAD of the walk with
// asynchronous copies (cp.async, 4 bytes each: a frame's stream is only 4-byte aligned), a chunk at a time, as soon as the
// slots of a chunk hold words the walk has passed.  A step may look at words [cw, cw + kRingLook): 32 groups of the
// longest possible kind.  On the host the copies "land" when they are waited for (until then their slots hold a poison
// value), so the emulation fails if the kernel's waits were ever in the wrong place.
constexpr uint32_t kRingWords = 4096;      // 16 KB per warp
constexpr uint32_t kRingChunk = 1024;
constexpr uint32_t kRingLook = 704;        // 31 * (96 + 1 + 8 * 72) + 96 + 32 bits, rounded up, in words
constexpr uint32_t kWide = 4;              // groups per lane in a wide step (below)
constexpr uint32_t kRingLookWide = 2816;   // the same for 32 * kWide groups; kRingWords >= kRingChunk + kRingLookWide
constexpr uint32_t kRingPoison = 0xDEADBEEFu;
static_assert((kRingWords & (kRingWords - 1u)) == 0u, "ring indices are masked");
static_assert(kRingWords >= kRingChunk + kRingLookWide, "a chunk may only overwrite words the widest step no longer looks at");
constexpr uint32_t kMaxGroupBits = 96u + 1u + 8u * 72u;   // large triple + flag + eight small triples of the widest kind
static_assert((31u * kMaxGroupBits + 96u + 32u + 31u) / 32u <= kRingLook, "narrow step: 32 groups of the longest kind");
static_assert(((32u * kWide - 1u) * kMaxGroupBits + 96u + 32u + 31u) / 32u <= kRingLookWide, "wide step: 128 groups of the longest kind");
struct Ring {
  uint32_t* w;           // [kRingWords]
  const uint32_t* src;   // the frame's stream in global memory
  uint32_t nwords;
  uint32_t issued;       // words [0, issued) have been asked for
  uint32_t ready;        // words [0, ready) have landed (those of them that are younger than issued - kRingWords are in the ring)
};
XTC_HD void ring_init(Ring& r, uint32_t* words, const uint32_t* src, uint32_t nwords) {
  r.w = words;
  r.src = src;
  r.nwords = nwords;
  r.issued = r.ready = 0u;
}
// may the next chunk be asked for while the walk stands at word cw?  (its slots hold words below cw by then)
XTC_HD bool ring_can_issue(const Ring& r, uint32_t cw) { return r.issued < r.nwords && r.issued + kRingChunk <= cw + kRingWords; }
// one lane's share of the next chunk
XTC_HD void ring_issue_lane(const Ring& r, uint32_t lane) {
  for (uint32_t j = lane; j < kRingChunk; j += 32u) {
    const uint32_t k = r.issued + j;
    if (k < r.nwords) {
#if defined(__CUDA_ARCH__)
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(r.w + (k & (kRingWords - 1u)));
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(r.src + k) : "memory");
#else
      r.w[k & (kRingWords - 1u)] = kRingPoison;
#endif
    }
  }
}
XTC_HD void ring_issued(Ring& r) {   // after every lane has issued its share
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
  r.issued += kRingChunk;
}
// does the step at word cw need words that are still on their way?
XTC_HD bool ring_must_wait(const Ring& r, uint32_t cw, uint32_t look) { return r.ready < r.issued && cw + look > r.ready; }
XTC_HD void ring_wait(Ring& r) {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncwarp();
#else
  for (uint32_t k = r.ready; k < r.issued && k < r.nwords; k++) r.w[k & (kRingWords - 1u)] = r.src[k];
#endif
  r.ready = r.issued;
}
XTC_HD uint32_t ring_peek32(const Ring& r, uint32_t p) {
  const uint32_t idx = p >> 5, sh = p & 31u;
  const uint32_t a = bswap32(r.w[idx & (kRingWords - 1u)]), b = bswap32(r.w[(idx + 1u) & (kRingWords - 1u)]);
#if defined(__CUDA_ARCH__)
  return __funnelshift_l(b, a, sh);
#else
  return sh ? (a << sh) | (b >> (32 - sh)) : a;
#endif
}
// One step of the speculative walk, written for a short dependent chain (the walk is one warp's latency per step: nothing
// else runs on its scheduler).  The warp-uniform state:
struct Walk {
  uint32_t pos;          // bit position of the next group
  uint32_t atom;         // its first atom
  uint32_t smallidx;     // table index of its small triples
  uint32_t nsm;          // small atoms per group while the flag bits stay clear (run / 3)
  uint32_t prev_atom;    // first atom of the last group consumed; kNoAtom before the first
};
// What lane l derives when it assumes that the l groups before its own had clear flag bits: where its group starts, its flag
// bit, and the walk's state AFTER its group — for a clear flag as well as for a set one, so that the warp only has to pick
// the right lane's answer (three shuffles behind the ballot; no branch).
struct LaneView {
  uint32_t pos, atom;    // the lane's group
  uint32_t flag;         // its flag bit
  uint32_t npos, natom;  // the group after it
  uint32_t nstate;       // smallidx | nsm << 8 | ok << 16 after it
};
XTC_HD LaneView walk_lane(const Ring& r, const Walk& w, uint32_t large_bits, uint32_t natoms, uint32_t lane) {
  const uint32_t len = large_bits + 1u + w.nsm * w.smallidx, gs = 1u + w.nsm;
  LaneView v;
  v.pos = w.pos + lane * len;
  v.atom = w.atom + lane * gs;
  const uint32_t bits = ring_peek32(r, v.pos + large_bits);
  v.flag = bits >> 31;
  const uint32_t r5 = (bits >> 26) & 31u;
  const uint32_t q = (r5 * 11u) >> 5;                       // r5 / 3 for r5 < 32
  const uint32_t nsm2 = v.flag ? q : w.nsm;
  const uint32_t sidx2 = v.flag ? w.smallidx + (r5 - 3u * q) - 1u : w.smallidx;   // += r5 % 3 - 1
  v.npos = v.pos + large_bits + (v.flag ? 6u : 1u) + nsm2 * w.smallidx;
  v.natom = v.atom + 1u + nsm2;
  const uint32_t ok = (nsm2 <= (uint32_t)kMaxRunAtoms && v.natom <= natoms && sidx2 >= (uint32_t)kFirstIdx && sidx2 <= (uint32_t)kLastIdx) ? 1u : 0u;
  v.nstate = sidx2 | (nsm2 << 8) | (ok << 16);
  return v;
}
// the lane whose group ends the step: the first valid lane with a set flag, else the last valid lane
XTC_HD uint32_t walk_winner(uint32_t set, uint32_t nv) {
#if defined(__CUDA_ARCH__)
  const uint32_t fl = set ? (uint32_t)__ffs((int)set) - 1u : 32u;
#else
  const uint32_t fl = set ? (uint32_t)__builtin_ctz(set) : 32u;
#endif
  return fl < nv ? fl : nv - 1u;
}
// the checkpoint a lane leaves when its group is consumed by the step and opens a 16-atom window
XTC_HD void walk_lane_emit(uint32_t lane, const LaneView& v, const Walk& w, uint32_t winner, U4* table) {
  if (lane > winner) return;
  const uint32_t before = lane == 0u ? w.prev_atom : v.atom - (1u + w.nsm);
  if (before == kNoAtom || (before / kCkAtoms) < (v.atom / kCkAtoms)) {
    U4 rec;
    rec.x = v.pos; rec.y = v.atom; rec.z = w.smallidx | ((3u * w.nsm) << 8); rec.w = 0u;
    table[v.atom / kCkAtoms] = rec;
  }
}
// the state after the step, from the winner's view (npos, natom, nstate as shuffled from lane `winner`).  False: impossible.
XTC_HD bool walk_advance(Walk& w, uint32_t winner, uint32_t npos, uint32_t natom, uint32_t nstate) {
  w.prev_atom = w.atom + winner * (1u + w.nsm);
  w.pos = npos;
  w.atom = natom;
  w.smallidx = nstate & 0xFFu;
  w.nsm = (nstate >> 8) & 0xFFu;
  return (nstate >> 16) != 0u;
}

// The wide step, for stretches without a single set flag bit (isolated atoms, or runs of one length: nothing but the
// position changes from group to group): every lane looks at kWide groups, 32 * kWide in all, and the step consumes the
// clear groups in front of the first set flag (or the last atom) — none of the per-lane "state after my group" work of
// the narrow step is needed, because a clear flag leaves the state as it is.  The walk goes wide after a narrow step that
// found 32 clear flags and stays wide while every flag it sees is clear; a set flag is always taken by a narrow step.
XTC_HD uint32_t wide_flag(const Ring& r, const Walk& w, uint32_t large_bits, uint32_t len, uint32_t g) {
  return ring_peek32(r, w.pos + g * len + large_bits) >> 31;
}
// set[k] / valid[k]: ballots over the lanes for group lane + 32 k.  The number of leading groups that exist and are clear.
XTC_HD uint32_t wide_count(const uint32_t set[kWide], const uint32_t valid[kWide]) {
  uint32_t n = 0;
  for (uint32_t k = 0; k < kWide; k++) {
    const uint32_t stop = set[k] | ~valid[k];
    if (stop) {
#if defined(__CUDA_ARCH__)
      return n + (uint32_t)__ffs((int)stop) - 1u;
#else
      return n + (uint32_t)__builtin_ctz(stop);
#endif
    }
    n += 32u;
  }
  return n;
}
XTC_HD void wide_lane_emit(uint32_t lane, const Walk& w, uint32_t len, uint32_t gs, uint32_t n, U4* table) {
  for (uint32_t k = 0; k < kWide; k++) {
    const uint32_t g = lane + 32u * k;
    if (g >= n) return;
    const uint32_t atom = w.atom + g * gs;
    const uint32_t before = g == 0u ? w.prev_atom : atom - gs;
    if (before == kNoAtom || (before / kCkAtoms) < (atom / kCkAtoms)) {
      U4 rec;
      rec.x = w.pos + g * len; rec.y = atom; rec.z = w.smallidx | ((3u * w.nsm) << 8); rec.w = 0u;
      table[atom / kCkAtoms] = rec;
    }
  }
}
// after n clear groups.  False: the last of them passes the atom count (impossible stream).
XTC_HD bool wide_advance(Walk& w, uint32_t len, uint32_t gs, uint32_t n, uint32_t natoms) {
  if (n == 0u) return true;
  w.prev_atom = w.atom + (n - 1u) * gs;
  w.pos += n * len;
  w.atom += n * gs;
  return w.atom <= natoms;
}

// ---- decode pass: one CTA of kXtcThreads threads per window of kXtcWindow atoms ------------------------------------------
constexpr int kXtcThreads = 256;                                  // checkpoints per CTA
constexpr int kXtcWindow = kXtcThreads * kCkAtoms;                // atoms per CTA: 4096
constexpr int kXtcStage = kXtcWindow + 16;                        // atoms staged: the last group may pass the window by up to 8
constexpr int kXtcStagePitch = kXtcStage + kXtcStage / 16 + 1;    // one float of padding per 16 atoms (bank spread)
constexpr int kXtcStreamWords = 5888;                             // stream words staged per CTA by default: 23 KB, three CTAs per SM
constexpr int kXtcStreamWordsMax = 12288;                         // ... when the frames need more per window (two CTAs per SM); longer spans stay in global memory

struct DecodeArgs {
  const char* text;
  int64_t text_bytes;
  const int64_t* coords_off;
  int n_atoms, n_pad, n_ck;
  const U4* ck;
  const int32_t* slot;
  const float* slot_dq;
  int sites_stride;
  int stream_words;           // capacity of the CTA's stream buffer (kXtcStreamWords .. kXtcStreamWordsMax)
  float* xyz8;
  F4* sites;
  int* err_flags;
};
struct DecodeCta {            // the CTA's shared state
  float* stage;               // [3][kXtcStagePitch]
  uint32_t* sw;               // [DecodeArgs::stream_words]
  FrameHdr h;
  const uint32_t* hw;         // the frame's coordinate block (null: unusable)
  uint32_t a_first, a_end;    // atoms this CTA owns: [a_first, a_end); a_first == kNoAtom: none
  uint32_t w_lo, w_hi;        // stream words its checkpoints span
  int staged;                 // the span sits in sw
};
// thread 0, before the first barrier
XTC_HD void decode_cta_setup(const DecodeArgs& a, DecodeCta& cta, int frame, uint32_t cta_in_frame) {
  const uint32_t natoms = (uint32_t)a.n_atoms;
  const uint32_t k0 = cta_in_frame * kXtcThreads;
  const U4* my = a.ck + (size_t)frame * a.n_ck;
  cta.hw = frame_block(a.text, a.text_bytes, a.coords_off[frame], a.n_atoms, cta.h);
  cta.a_first = kNoAtom;
  cta.a_end = 0u;
  cta.w_lo = cta.w_hi = 0u;
  cta.staged = 0;
  if (cta.hw == nullptr) return;
  const uint32_t nwords = (cta.h.nbytes + 3u) >> 2;
  const U4 first = my[k0];
  if (first.y >= natoms) return;
  uint32_t end_pos = 8u * cta.h.nbytes, a_end = natoms;
  if (k0 + kXtcThreads < (uint32_t)a.n_ck) {
    const U4 nx = my[k0 + kXtcThreads];
    if (nx.y < natoms) { a_end = nx.y; end_pos = nx.x; }
  }
  uint32_t w_lo = first.x >> 5;
  uint32_t w_hi = ((end_pos + 31u) >> 5) + 3u;   // a read of 32 bits at the last position touches two words
  if (w_hi > nwords) w_hi = nwords;
  if (w_lo >= w_hi) return;                      // impossible table
  cta.a_first = first.y;
  cta.a_end = a_end;
  cta.w_lo = w_lo;
  cta.w_hi = w_hi;
  cta.staged = (w_hi - w_lo) <= (uint32_t)a.stream_words;
}
// every thread, before the first barrier (the loads are in flight while thread 0 sets the CTA up): its own checkpoint and
// the atom the next one starts at
struct MyCheckpoint {
  U4 me;
  uint32_t next_atom;   // kNoAtom: there is no next checkpoint
};
XTC_HD MyCheckpoint decode_cta_my_checkpoint(const DecodeArgs& a, int frame, uint32_t cta_in_frame, uint32_t tid) {
  const U4* my = a.ck + (size_t)frame * a.n_ck;
  const uint32_t k = cta_in_frame * kXtcThreads + tid;
  MyCheckpoint m;
  m.me.x = 0u; m.me.y = kNoAtom; m.me.z = 0u; m.me.w = 0u;
  m.next_atom = kNoAtom;
  if (k < (uint32_t)a.n_ck) m.me = my[k];
  if (k + 1u < (uint32_t)a.n_ck) m.next_atom = my[k + 1u].y;
  return m;
}
// every thread, between the first and the second barrier: the stream span into shared memory.  On the device the copies are
// asynchronous (cp.async, 4 bytes each: the span is only 4-byte aligned) so that all of them are in flight at once; they
// are waited for just before the second barrier.  The words stay in file byte order (Stream.swap).
XTC_HD void decode_cta_stage_stream(DecodeCta& cta, uint32_t tid) {
  if (cta.a_first == kNoAtom || !cta.staged) return;
  const uint32_t n = cta.w_hi - cta.w_lo;
  const uint32_t* src = cta.hw + 10u + cta.w_lo;
  for (uint32_t i = tid; i < n; i += kXtcThreads) {
#if defined(__CUDA_ARCH__)
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(cta.sw + i);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src + i) : "memory");
#else
    cta.sw[i] = src[i];
#endif
  }
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}
// every thread, between the second and the third barrier: the atoms of its checkpoint into the staging array
XTC_HD void decode_cta_unpack(const DecodeArgs& a, DecodeCta& cta, uint32_t cta_in_frame, uint32_t tid, const MyCheckpoint& m) {
  if (cta.a_first == kNoAtom) return;
  const uint32_t natoms = (uint32_t)a.n_atoms;
  const uint32_t window = cta_in_frame * kXtcWindow;
  const uint32_t k = cta_in_frame * kXtcThreads + tid;
  const U4 me = m.me;
  if (me.y >= natoms) return;
  const uint32_t next_atom = m.next_atom < natoms ? m.next_atom : natoms;
  uint32_t limit = window + (uint32_t)kXtcStage;
  if (limit > natoms) limit = natoms;
  // a checkpoint lies in its own 16-atom window; anything else is a damaged table
  bool ok = me.y / kCkAtoms == k && next_atom > me.y && next_atom <= limit;
  if (ok) {
    Stream s;
    s.swap = true;
    if (cta.staged) { s.w = cta.sw; s.lo = cta.w_lo; s.hi = cta.w_hi; }
    else { s.w = cta.hw + 10; s.lo = 0u; s.hi = (cta.h.nbytes + 3u) >> 2; }
    const float precision = cta.h.precision;
    float* stage = cta.stage;
    ok = decode_groups(s, cta.h, unpack_cursor(me.x, me.y, me.z), next_atom, limit, [&](uint32_t atom, const int* kxyz) {
      const uint32_t j = atom - window;
      const uint32_t at = j + (j >> 4);
      stage[at] = to_coord(kxyz[0], precision);
      stage[kXtcStagePitch + at] = to_coord(kxyz[1], precision);
      stage[2 * kXtcStagePitch + at] = to_coord(kxyz[2], precision);
    });
  }
  if (!ok) raise_flag(a.err_flags, kDecBadXtcBit);
}
// every thread, after the third barrier: coordinate blocks and site entries back to global memory.  The blocks of 8 atoms
// of a window are contiguous in xyz8 ([x0..x7][y0..y7][z0..z7] per block): a thread writes four consecutive words — half a
// row of a block, one 16-byte store when the CTA owns all four atoms — and consecutive threads consecutive quads.
XTC_HD void decode_cta_write_back(const DecodeArgs& a, const DecodeCta& cta, int frame, uint32_t cta_in_frame, bool last_cta,
                                  uint32_t tid) {
  const uint32_t natoms = (uint32_t)a.n_atoms;
  const uint32_t window = cta_in_frame * kXtcWindow;
  const bool have = cta.a_first != kNoAtom;
  float* dst = a.xyz8 + ((size_t)frame * a.n_pad + window) * 3;
  for (uint32_t idx = 4u * tid; idx < 3u * (uint32_t)kXtcStage; idx += 4u * kXtcThreads) {
    const uint32_t blk = idx / 24u, rem = idx - 24u * blk;             // rem: 0, 4, .., 20
    const uint32_t j = blk * 8u + (rem & 7u), atom = window + j;       // the quad holds atoms atom .. atom + 3, one coordinate
    const float* src = cta.stage + (rem >> 3) * kXtcStagePitch + j + (j >> 4);
    if (have && atom >= cta.a_first && atom + 3u < cta.a_end) {
      F4 q;
      q.x = src[0]; q.y = src[1]; q.z = src[2]; q.w = src[3];
      *reinterpret_cast<F4*>(dst + idx) = q;
    } else {
      for (uint32_t e = 0; e < 4u; e++) {
        if (have && atom + e >= cta.a_first && atom + e < cta.a_end) {
          dst[idx + e] = src[e];
        } else if (last_cta && atom + e >= natoms && atom + e < (uint32_t)a.n_pad) {
          dst[idx + e] = 0.f;   // tile padding: its kq is 0 and its column is -1
        }
      }
    }
  }
  if (!have) return;
  // site entries: the slot numbers of the window first (independent loads), then the few entries there are
  constexpr int kRounds = (kXtcStage + kXtcThreads - 1) / kXtcThreads;
  int sl[kRounds];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int rd = 0; rd < kRounds; rd++) {
    const uint32_t j = (uint32_t)rd * kXtcThreads + tid, atom = window + j;
    sl[rd] = (j < (uint32_t)kXtcStage && atom >= cta.a_first && atom < cta.a_end) ? a.slot[atom] : -1;
  }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int rd = 0; rd < kRounds; rd++) {
    if (sl[rd] >= 0) {
      const uint32_t j = (uint32_t)rd * kXtcThreads + tid;
      const uint32_t at = j + (j >> 4);
      F4 e;
      e.x = -cta.stage[at]; e.y = -cta.stage[kXtcStagePitch + at]; e.z = -cta.stage[2 * kXtcStagePitch + at];
      e.w = a.slot_dq[sl[rd]];
      a.sites[(size_t)frame * a.sites_stride + sl[rd]] = e;
    }
  }
}

}  // namespace xtc
}  // namespace cgc

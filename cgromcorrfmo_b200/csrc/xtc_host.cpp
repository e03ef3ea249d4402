// GROMACS .xtc on the host: the frame index (one header per frame — the streams themselves are decoded on the device,
// xtc_kernel.cu) and a frame writer (cgc_xtc_encode; the synthetic benchmarks and tests need .xtc files and no GROMACS
// exists offline).  Stream format: xtc_bits.h.
#include "host_parse.h"
#include "xtc_bits.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace cgc {

namespace {
inline uint32_t be32(const char* p) {
  const unsigned char* u = reinterpret_cast<const unsigned char*>(p);
  return ((uint32_t)u[0] << 24) | ((uint32_t)u[1] << 16) | ((uint32_t)u[2] << 8) | (uint32_t)u[3];
}
inline float be_float(const char* p) {
  const uint32_t b = be32(p);
  float f;
  memcpy(&f, &b, 4);
  return f;
}
}  // namespace

bool locate_xtc_frame(const char* data, int64_t nbytes, int64_t pos, XtcFrame& out, bool* truncated) {
  if (truncated) *truncated = false;
  if (pos >= nbytes) return false;
  auto cut = [&]() {
    if (truncated) *truncated = true;
    return false;
  };
  if (pos + 4 > nbytes) return cut();
  if (be32(data + pos) != 1995u) throw InputFileMalformed("not an .xtc frame: magic number 1995 expected");
  if (pos + 56 > nbytes) return cut();
  const int32_t natoms = (int32_t)be32(data + pos + 4);
  if (natoms <= 0) throw InputFileMalformed(".xtc frame with no atoms");
  if ((int32_t)be32(data + pos + 52) != natoms) throw InputFileMalformed(".xtc frame: the two atom counts differ");
  if (natoms <= 9) throw Unsupported(".xtc frames of at most 9 atoms are stored uncompressed; not read here");
  const int64_t c = pos + 52;
  if (c + 40 > nbytes) return cut();
  uint32_t raw[10];
  memcpy(raw, data + c, sizeof raw);
  const xtc::FrameHdr h = xtc::parse_header(raw);
  if (!h.ok) throw InputFileMalformed(".xtc frame header holds impossible values (precision, coordinate range, table index or byte count)");
  out.begin = pos;
  out.coords = c;
  out.end = c + 40 + (((int64_t)h.nbytes + 3) & ~(int64_t)3);
  out.n_atoms = natoms;
  out.step = (int32_t)be32(data + pos + 8);
  out.time = be_float(data + pos + 12);
  out.precision = h.precision;
  if (out.end > nbytes) return cut();
  return true;
}

// ---------------------------------------------------------------------------------------------------------------------
// writer
// ---------------------------------------------------------------------------------------------------------------------
namespace {

struct BitSink {
  std::vector<char>& out;
  uint64_t acc = 0;   // pending bits, right-aligned
  int n = 0;          // how many
  explicit BitSink(std::vector<char>& o) : out(o) {}
  void bits(int nbits, uint32_t v) {   // nbits <= 32
    if (nbits == 0) return;
    acc = (acc << nbits) | (v & (nbits == 32 ? 0xFFFFFFFFu : ((1u << nbits) - 1u)));
    n += nbits;
    while (n >= 8) {
      out.push_back((char)((acc >> (n - 8)) & 0xFF));
      n -= 8;
    }
  }
  // (a * s1 + b) * s2 + c, least significant byte first, the last byte with the bits that are left
  void packed(int nbits, const uint32_t sizes[3], const uint32_t nums[3]) {
    unsigned __int128 v = ((unsigned __int128)nums[0] * sizes[1] + nums[1]) * sizes[2] + nums[2];
    while (nbits > 8) {
      bits(8, (uint32_t)(v & 0xFF));
      v >>= 8;
      nbits -= 8;
    }
    bits(nbits, (uint32_t)v);
  }
  void flush() {
    if (n > 0) {
      out.push_back((char)((acc << (8 - n)) & 0xFF));
      n = 0;
    }
  }
};

void put_be32(std::vector<char>& out, uint32_t v) {
  out.push_back((char)(v >> 24));
  out.push_back((char)(v >> 16));
  out.push_back((char)(v >> 8));
  out.push_back((char)v);
}
void put_float(std::vector<char>& out, float f) {
  uint32_t b;
  memcpy(&b, &f, 4);
  put_be32(out, b);
}
inline bool within(const int32_t* a, const int32_t* b, int64_t lim) {
  return std::llabs((int64_t)a[0] - b[0]) < lim && std::llabs((int64_t)a[1] - b[1]) < lim && std::llabs((int64_t)a[2] - b[2]) < lim;
}

}  // namespace

void xtc_encode_frame(const int32_t* ints, int natoms, int step, float time, const float box[9], float precision,
                      std::vector<char>& out) {
  if (natoms <= 9) throw Unsupported("xtc writer: frames of at most 9 atoms are stored uncompressed; not written here");
  using xtc::kFirstIdx;
  using xtc::kLastIdx;
  put_be32(out, 1995u);
  put_be32(out, (uint32_t)natoms);
  put_be32(out, (uint32_t)step);
  put_float(out, time);
  for (int i = 0; i < 9; i++) put_float(out, box[i]);
  put_be32(out, (uint32_t)natoms);
  put_float(out, precision);
  std::vector<int32_t> c(ints, ints + (size_t)3 * natoms);   // pairs of atoms change places while encoding
  int32_t lo[3], hi[3];
  for (int d = 0; d < 3; d++) lo[d] = hi[d] = c[d];
  int64_t mindiff = INT64_MAX;
  for (int i = 0; i < natoms; i++) {
    for (int d = 0; d < 3; d++) {
      lo[d] = std::min(lo[d], c[(size_t)3 * i + d]);
      hi[d] = std::max(hi[d], c[(size_t)3 * i + d]);
    }
    if (i > 0) {
      int64_t diff = 0;
      for (int d = 0; d < 3; d++) diff += std::llabs((int64_t)c[(size_t)3 * i + d] - c[(size_t)3 * (i - 1) + d]);
      mindiff = std::min(mindiff, diff);
    }
  }
  uint32_t sizeint[3];
  for (int d = 0; d < 3; d++) {
    put_be32(out, (uint32_t)lo[d]);
  }
  for (int d = 0; d < 3; d++) {
    put_be32(out, (uint32_t)hi[d]);
    sizeint[d] = (uint32_t)((int64_t)hi[d] - lo[d] + 1);
  }
  int bitsize = 0, bits3[3] = {0, 0, 0};
  if ((sizeint[0] | sizeint[1] | sizeint[2]) > 0xFFFFFFu) {
    for (int d = 0; d < 3; d++) bits3[d] = xtc::bitlen32(sizeint[d]);
  } else {
    unsigned __int128 p = (unsigned __int128)sizeint[0] * sizeint[1] * sizeint[2];
    while (p) { bitsize++; p >>= 1; }
  }
  int smallidx = kFirstIdx;
  while (smallidx < kLastIdx && (int64_t)xtc::kMagicHost[smallidx] < mindiff) smallidx++;
  put_be32(out, (uint32_t)smallidx);
  const size_t count_at = out.size();
  put_be32(out, 0u);   // byte count, known at the end
  const size_t stream_at = out.size();

  const int maxidx = std::min(kLastIdx, smallidx + 8), minidx = maxidx - 8;
  int64_t smaller = xtc::kMagicHost[std::max(kFirstIdx, smallidx - 1)] / 2;
  int64_t smallnum = xtc::kMagicHost[smallidx] / 2;
  const int64_t larger = xtc::kMagicHost[maxidx] / 2;
  BitSink w(out);
  int32_t prev[3] = {0, 0, 0};
  int prevrun = -1;
  int i = 0;
  while (i < natoms) {
    int32_t* cur = &c[(size_t)3 * i];
    int is_smaller;
    if (smallidx < maxidx && i >= 1 && within(cur, prev, larger)) is_smaller = 1;
    else if (smallidx > minidx) is_smaller = -1;
    else is_smaller = 0;
    bool is_small = false;
    if (i + 1 < natoms && within(cur, cur + 3, smallnum)) {
      for (int d = 0; d < 3; d++) std::swap(cur[d], cur[3 + d]);   // the second atom of a close pair goes first
      is_small = true;
    }
    uint32_t rel[3];
    for (int d = 0; d < 3; d++) rel[d] = (uint32_t)((int64_t)cur[d] - lo[d]);
    if (bitsize == 0) {
      for (int d = 0; d < 3; d++) w.bits(bits3[d], rel[d]);
    } else {
      w.packed(bitsize, sizeint, rel);
    }
    for (int d = 0; d < 3; d++) prev[d] = cur[d];
    i++;
    if (!is_small && is_smaller == -1) is_smaller = 0;
    uint32_t run[3 * xtc::kMaxRunAtoms];
    int nrun = 0;
    while (is_small && nrun < 3 * xtc::kMaxRunAtoms) {
      const int32_t* a = &c[(size_t)3 * i];
      if (is_smaller == -1) {
        int64_t r2 = 0;
        for (int d = 0; d < 3; d++) r2 += ((int64_t)a[d] - prev[d]) * ((int64_t)a[d] - prev[d]);
        if (r2 >= smaller * smaller) is_smaller = 0;
      }
      for (int d = 0; d < 3; d++) run[nrun++] = (uint32_t)((int64_t)a[d] - prev[d] + smallnum);
      for (int d = 0; d < 3; d++) prev[d] = a[d];
      i++;
      is_small = i < natoms && within(&c[(size_t)3 * i], prev, smallnum);
    }
    if (nrun != prevrun || is_smaller != 0) {
      prevrun = nrun;
      w.bits(1, 1u);
      w.bits(5, (uint32_t)(nrun + is_smaller + 1));
    } else {
      w.bits(1, 0u);
    }
    const uint32_t ss = xtc::kMagicHost[smallidx];
    const uint32_t sizes[3] = {ss, ss, ss};
    for (int k = 0; k < nrun; k += 3) w.packed(smallidx, sizes, &run[k]);
    if (is_smaller != 0) {
      smallidx += is_smaller;
      if (is_smaller < 0) {
        smallnum = smaller;
        smaller = xtc::kMagicHost[smallidx - 1] / 2;
      } else {
        smaller = smallnum;
        smallnum = xtc::kMagicHost[smallidx] / 2;
      }
    }
  }
  w.flush();
  const uint32_t nbytes = (uint32_t)(out.size() - stream_at);
  out[count_at] = (char)(nbytes >> 24);
  out[count_at + 1] = (char)(nbytes >> 16);
  out[count_at + 2] = (char)(nbytes >> 8);
  out[count_at + 3] = (char)nbytes;
  while ((out.size() - stream_at) % 4) out.push_back(0);
}

}  // namespace cgc

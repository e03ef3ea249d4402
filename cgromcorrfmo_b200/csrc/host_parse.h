// Host-side text front end: charge tables from topology/itp, and the frame-0 pass over the .gro trajectory.
// Everything here runs once per trajectory; the per-frame work is on the device (kernels.cu).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace cgc {

// Mirrors of the reference's three exception classes (reference src/grom2atomFMO.h:12-27); they never cross the C-ABI.
struct InputFileMalformed : std::logic_error { using std::logic_error::logic_error; };
struct InputFileMinorMalformed : std::logic_error { using std::logic_error::logic_error; };
struct ReadError : std::runtime_error { using std::runtime_error::runtime_error; };
struct Unsupported : std::runtime_error { using std::runtime_error::runtime_error; };

// The reference's three parallel vectors (mergeNameTable, massTable, chargeTable) plus a first-wins hash index that
// replaces both its O(T^2) duplicate scan (reference src/grom2atomFMO.cpp:188-207) and the std::map rebuilt per site
// per frame in AtomDataLookup_v2 (reference src/grom2atomFMO.cpp:565-568).
struct ChargeTable {
  std::vector<std::string> names;
  std::vector<float> masses;
  std::vector<float> charges;
  std::unordered_map<std::string, int> index;
  int warnings = 0;  // "conflicts with prior entry" notes

  int find(const std::string& key) const {
    auto it = index.find(key);
    return it == index.end() ? -1 : it->second;
  }
};

// ParseTopology (reference src/grom2atomFMO.cpp:78-234).  Appends to `t`.  Throws InputFileMalformed where the
// reference does; throws InputFileMinorMalformed at the end of a file that produced conflict notes (entries parsed so
// far stay in the table, and — as in the reference — the exception unwinds through any including file).
void parse_topology(const std::string& filename, ChargeTable& t, bool excited);

// main()'s two calls with Minor exceptions swallowed (reference src/FMOparse.cpp:355-366).
void load_tables(const std::string& topology, const std::string& excited_itp, ChargeTable& t);

// Result of the frame-0 pass.
struct FrameLayout {
  int n_atoms = 0;
  int line_bytes = 0;   // bytes per atom line including '\n'
  int x_col = 20;       // column of the first coordinate field
  int field_w = 8;      // width of one coordinate field
  int ndec = 3;         // digits after the decimal point
  int n_chromo = 0, n_groups = 0;
  std::vector<int32_t> groups;     // reference atomGroupi
  std::vector<std::string> keys;   // reference atomTypei
  std::vector<float> x0;           // frame-0 coordinates via (float)atof of the tokens, [3*n_atoms]
};

// One frame located inside the text.
struct FrameSpan {
  int64_t begin = 0;        // offset of the title line
  int64_t atoms = 0;        // offset of the first atom line
  int64_t end = 0;          // offset one past the last box line (== begin of the next frame)
};

// Locate the frame that starts at `pos` (title line containing "t=", count line, n_atoms lines of line_bytes, box
// line(s)).  n_atoms/line_bytes == 0 means "not known yet" (frame 0).  Returns false at a clean end of text.
// Throws InputFileMalformed for a broken header, ReadError for a stray line before any header.
bool locate_frame(const char* text, int64_t nbytes, int64_t pos, int n_atoms, int line_bytes, FrameSpan& out,
                  int* n_atoms_found, bool* truncated);

// ---- GROMACS .trr (binary, XDR big-endian) --------------------------------------------------------------------------
// Not a reference format: the authors convert .trr to .gro text with trjconv before traj2dEcsv
// (reference scripts/2014-09-trr2dEcsv.sh:22); reading the .trr directly removes the 45-69 B/atom text tax.
// Frame = header (magic 1993, version string "GMX_trn_file", 13 int32: ir/e/box/vir/pres/top/sym/x/v/f sizes, natoms,
// step, nre; then t and lambda as float or double) followed by box, vir, pres (9 reals each when present), x, v, f
// (3*natoms reals each when present).  The precision is box_size / 9, else x_size / (3*natoms)  (GROMACS
// src/gromacs/fileio/trrio.cpp, do_trr_frame_header / nFloatSize).
struct TrrFrame {
  int64_t begin = 0;     // offset of the magic number
  int64_t x_off = 0;     // offset of the coordinate array
  int64_t end = 0;       // offset one past the frame
  int n_atoms = 0;
  int real_bytes = 4;    // 4 (single precision) or 8
  int64_t step = 0;
  double time = 0.0;
};
// Parse the frame header at `pos`.  Returns false at a clean end of data (pos == nbytes) or when the frame is cut short
// (*truncated = true).  Throws InputFileMalformed for a bad magic/version/size field, Unsupported for a frame without
// coordinates.
bool locate_trr_frame(const char* data, int64_t nbytes, int64_t pos, TrrFrame& out, bool* truncated);

// ---- GROMACS .xtc (compressed coordinates, XDR big-endian; layout and stream format: xtc_bits.h) ----------------------
// Frame = int magic 1995, int natoms, int step, float time, float box[9], then the coordinate block: int natoms, float
// precision, int minint[3], int maxint[3], int smallidx, int nbytes, nbytes of bit stream padded to a multiple of 4.
struct XtcFrame {
  int64_t begin = 0;     // offset of the magic number
  int64_t coords = 0;    // offset of the coordinate block (its atom count field): what the device kernels are pointed at
  int64_t end = 0;       // offset one past the frame
  int n_atoms = 0;
  int64_t step = 0;
  double time = 0.0;
  float precision = 0.f;
};
// Parse the frame header at `pos`.  Returns false at a clean end of data or when the frame is cut short (*truncated).
// Throws InputFileMalformed for a bad magic number or impossible fields, Unsupported for frames of at most 9 atoms (they
// are stored uncompressed).
bool locate_xtc_frame(const char* data, int64_t nbytes, int64_t pos, XtcFrame& out, bool* truncated);
// Append one frame holding the integer coordinates ints[3 * natoms] (units of 1 / precision nm) to `out`, compressed the
// way GROMACS' writer chooses runs and table indices.  natoms > 9.
void xtc_encode_frame(const int32_t* ints, int natoms, int step, float time, const float box[9], float precision,
                      std::vector<char>& out);

// Frame-0 pass with the reference's whitespace-token rules (reference src/grom2atomFMO.cpp:396-462), then the
// fixed-column layout check.  Throws Unsupported when the atom lines are not fixed-column.
void scan_first_frame(const char* text, int64_t nbytes, FrameLayout& out, FrameSpan& span);

// memcpy for the staging copies (page cache -> pinned ring): with `streaming` the destination is written with
// non-temporal stores (AVX2, when the CPU has it), which spares the read-for-ownership of the destination lines — a
// third of the copy's DRAM traffic on a host whose memory system is what bounds the file path.
void stream_copy(char* dst, const char* src, size_t n, bool streaming);

}  // namespace cgc

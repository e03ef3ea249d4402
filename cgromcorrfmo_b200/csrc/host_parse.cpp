// Host-side text front end (see host_parse.h).  Runs once per trajectory.
#include "host_parse.h"

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <cctype>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

namespace cgc {

namespace {

struct Tok {
  const char* p;
  int n;
  std::string str() const { return std::string(p, n); }
  bool is(const char* s) const { return (int)strlen(s) == n && memcmp(p, s, n) == 0; }
};

// whitespace split, same separators as operator>> (isspace)
inline void split_ws(const char* s, int n, std::vector<Tok>& out) {
  out.clear();
  int i = 0;
  while (i < n) {
    while (i < n && isspace((unsigned char)s[i])) i++;
    if (i >= n) break;
    int b = i;
    while (i < n && !isspace((unsigned char)s[i])) i++;
    out.push_back(Tok{s + b, i - b});
  }
}

// "float x = std::atof(tok)": correctly rounded double, then narrowed (reference src/grom2atomFMO.cpp:158-162,448-450)
inline float f32_atof(const Tok& t) {
  char buf[64];
  int n = t.n < 63 ? t.n : 63;
  memcpy(buf, t.p, n);
  buf[n] = 0;
  return (float)atof(buf);
}

// "iss >> float": libstdc++ converts with strtof (reference src/grom2atomFMO.cpp:179)
inline float f32_stream(const Tok& t) {
  char buf[64];
  int n = t.n < 63 ? t.n : 63;
  memcpy(buf, t.p, n);
  buf[n] = 0;
  return strtof(buf, nullptr);
}

inline bool strict_int(const Tok& t, long& v) {
  if (t.n == 0 || t.n > 18) return false;
  int i = 0;
  if (t.p[0] == '+' || t.p[0] == '-') i = 1;
  if (i == t.n) return false;
  for (int k = i; k < t.n; k++)
    if (t.p[k] < '0' || t.p[k] > '9') return false;
  v = strtol(std::string(t.p, t.n).c_str(), nullptr, 10);
  return true;
}

std::string read_file(const std::string& path, bool& ok) {
  std::ifstream f(path, std::ios::binary);
  ok = f.is_open();
  if (!ok) return std::string();
  std::stringstream ss;
  ss << f.rdbuf();
  return ss.str();
}

std::string include_dir(const std::string& filename) {
  // The reference builds parent_path(filename) + "/" + name (src/grom2atomFMO.cpp:123-124), which turns a topology
  // path without a directory part into "/name".  Deliberate deviation: such includes resolve against ".".
  size_t k = filename.find_last_of('/');
  if (k == std::string::npos) return ".";
  if (k == 0) return "";
  return filename.substr(0, k);
}

}  // namespace

void parse_topology(const std::string& filename, ChargeTable& t, bool excited) {
  bool ok = false;
  std::string text = read_file(filename, ok);
  if (!ok) throw InputFileMalformed("Requested topology file " + filename + " could not be opened.");
  std::string notes;
  bool in_atoms = false;
  std::vector<Tok> tok;
  size_t pos = 0;
  while (pos < text.size()) {
    size_t nl = text.find('\n', pos);
    if (nl == std::string::npos) nl = text.size();
    const char* lp = text.data() + pos;
    int ln = (int)(nl - pos);
    pos = nl + 1;
    if (const void* sc = memchr(lp, ';', ln)) ln = (int)((const char*)sc - lp);  // ';' starts a comment (:98)
    split_ws(lp, ln, tok);
    bool header_line = false;
    if (tok.size() >= 3 && tok[0].is("[") && tok[2].is("]")) {  // section header (:107-119)
      if (tok[1].is("atoms")) {
        in_atoms = true;
        header_line = true;
      } else if (in_atoms) {
        in_atoms = false;
      }
    }
    if (!excited && tok.size() >= 1 && tok[0].is("#include")) {  // (:121-126)
      std::string name = tok.size() >= 2 ? tok[1].str() : std::string();
      std::string clean;
      for (char c : name)
        if (c != '"') clean.push_back(c);
      parse_topology(include_dir(filename) + "/" + clean, t, false);
    }
    if (!in_atoms || header_line || ln == 0) continue;  // (:128)

    std::string mol, atom;
    float q = 0.f, m = 0.f, q_ex = 0.f, m_ex = 0.f;
    long id = 0;
    if (excited) {
      // FEP line: 8 columns, or 11 with a B state (:149; the reference counts one trailing empty token, hence 9/12)
      if (tok.size() != 8 && tok.size() != 11) {
        std::ostringstream e;
        e << "Bad Topology file, found wrong number of line entries for FEP topology (expected either 8 or 11 and "
             "received <" << tok.size() + 1 << ">\nBad line = " << std::string(lp, ln);
        throw InputFileMalformed(e.str());
      }
      id = atoi(tok[0].str().c_str());
      mol = tok[3].str();
      atom = tok[4].str();
      q = f32_atof(tok[6]);
      m = f32_atof(tok[7]);
      if (tok.size() == 11) {
        q_ex = f32_atof(tok[9]);
        m_ex = f32_atof(tok[10]);
      } else {
        q_ex = q;
        m_ex = m;
      }
      if (m_ex != m) {
        throw InputFileMalformed("Bad Topology file, atom in " + mol + " has different A and B masses\nBad line: " +
                                 std::string(lp, ln));
      }
      q_ex = q_ex - q;  // only the charge difference enters the CDC sum (:175), float arithmetic
    } else {
      // nr type resnr residue atom cgnr charge mass (:179); a line whose first token is not an integer yields nothing
      if (tok.empty() || !strict_int(tok[0], id)) continue;
      if (tok.size() >= 4) mol = tok[3].str();
      if (tok.size() >= 5) atom = tok[4].str();
      if (tok.size() >= 7) q = f32_stream(tok[6]);
      if (tok.size() >= 8) m = f32_stream(tok[7]);
    }
    std::string key = mol + "," + atom + std::to_string(id);  // (:183)
    int at = t.find(key);
    if (at >= 0) {  // duplicate: first wins, conflicts are only noted (:188-207)
      if (t.masses[at] != m) {
        notes += "Mass of " + key + " conflicts with prior entry\n";
        t.warnings++;
      }
      if (t.charges[at] != q) {
        notes += "Charge of " + key + " conflicts with prior entry\n";
        t.warnings++;
      }
    } else if (!atom.empty()) {  // (:209)
      t.index.emplace(key, (int)t.names.size());
      t.names.push_back(key);
      t.masses.push_back(m);
      t.charges.push_back(q);
      if (excited) {
        std::string key_ex = key + ",EXC";  // (:185,214-218)
        t.index.emplace(key_ex, (int)t.names.size());
        t.names.push_back(key_ex);
        t.masses.push_back(m_ex);
        t.charges.push_back(q_ex);
      }
    }
  }
  if (!notes.empty()) throw InputFileMinorMalformed(notes);
}

void load_tables(const std::string& topology, const std::string& excited_itp, ChargeTable& t) {
  try {
    parse_topology(excited_itp, t, true);
  } catch (const InputFileMinorMalformed&) {
  }
  try {
    parse_topology(topology, t, false);
  } catch (const InputFileMinorMalformed&) {
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// trajectory
// ---------------------------------------------------------------------------------------------------------------------
namespace {

inline int64_t line_end(const char* text, int64_t nbytes, int64_t pos) {  // offset of '\n' or nbytes
  const void* p = memchr(text + pos, '\n', (size_t)(nbytes - pos));
  return p ? (const char*)p - text : nbytes;
}

// std::regex_match(line, ".*t=.*") (reference src/grom2atomFMO.cpp:360,366): '.' does not match '\r'
inline bool is_title(const char* s, int64_t n) {
  if (n < 2) return false;
  if (memchr(s, '\r', (size_t)n)) return false;
  for (int64_t i = 0; i + 1 < n; i++)
    if (s[i] == 't' && s[i + 1] == '=') return true;
  return false;
}

inline bool all_space(const char* s, int64_t n) {
  for (int64_t i = 0; i < n; i++)
    if (!isspace((unsigned char)s[i])) return false;
  return true;
}

}  // namespace

bool locate_frame(const char* text, int64_t nbytes, int64_t pos, int n_atoms, int line_bytes, FrameSpan& out,
                  int* n_atoms_found, bool* truncated) {
  if (truncated) *truncated = false;
  if (pos >= nbytes) return false;
  int64_t e = line_end(text, nbytes, pos);
  if (!is_title(text + pos, e - pos)) {
    if (all_space(text + pos, nbytes - pos)) return false;  // trailing blank lines: clean end
    throw ReadError("Mysterious line appeared!");           // (:486-488)
  }
  out.begin = pos;
  if (e >= nbytes) {
    throw InputFileMalformed("Encountered EOF when number-of-atoms was expected after time-header line.");
  }
  int64_t c0 = e + 1;
  if (c0 >= nbytes) {
    throw InputFileMalformed("Encountered EOF when number-of-atoms was expected after time-header line.");
  }
  int64_t c1 = line_end(text, nbytes, c0);
  {
    std::vector<Tok> tok;
    split_ws(text + c0, (int)(c1 - c0), tok);
    long v = 0;
    if (tok.size() != 1 || !strict_int(tok[0], v) || v < 0 || v > 2000000000L) {  // Convert<int> (:377)
      throw InputFileMalformed(
          "Encountered non-integer entry when number-of-atoms was excpected after time-header line.");
    }
    if (n_atoms_found) *n_atoms_found = (int)v;
    if (n_atoms > 0 && v != n_atoms) {
      throw InputFileMalformed("Number of atoms changes between frames; the group layout of frame 0 no longer applies.");
    }
    if (n_atoms <= 0) n_atoms = (int)v;
  }
  out.atoms = c1 + 1;
  int64_t p = out.atoms;
  if (line_bytes > 0) {
    p += (int64_t)n_atoms * line_bytes;
    if (p > nbytes) {
      // tolerate a missing final newline on the very last atom line
      if (p - 1 == nbytes) p = nbytes;
      else {
        if (truncated) *truncated = true;
        return false;
      }
    }
    if (n_atoms > 0) {
      // cheap host check of the first and last line; every line is re-checked on the device while decoding
      bool first_ok = text[out.atoms + line_bytes - 1] == '\n';
      bool last_ok = (p == nbytes && text[p - 1] != '\n') || text[p - 1] == '\n';
      if (!first_ok || !last_ok) {
        throw Unsupported("atom lines of this frame do not have the fixed length found in frame 0");
      }
    }
  } else {
    for (int i = 0; i < n_atoms; i++) {
      if (p >= nbytes) {
        if (truncated) *truncated = true;
        return false;
      }
      p = line_end(text, nbytes, p) + 1;
    }
    if (p > nbytes) p = nbytes;
  }
  // box line(s): everything up to the next title line (:472-485)
  while (p < nbytes) {
    int64_t le = line_end(text, nbytes, p);
    if (is_title(text + p, le - p)) break;
    p = le + 1;
  }
  out.end = p > nbytes ? nbytes : p;
  return true;
}

void scan_first_frame(const char* text, int64_t nbytes, FrameLayout& L, FrameSpan& span) {
  int n = 0;
  bool trunc = false;
  if (!locate_frame(text, nbytes, 0, 0, 0, span, &n, &trunc)) {
    if (trunc) throw InputFileMalformed("first frame of the trajectory is incomplete");
    throw ReadError("no frame header (a title line containing 't=') found in the trajectory");
  }
  L.n_atoms = n;
  L.groups.assign(n, 0);
  L.keys.assign(n, std::string());
  L.x0.assign((size_t)3 * n, 0.f);
  if (n == 0) throw InputFileMalformed("trajectory frame with zero atoms");

  // --- layout from the first atom line: three decimal points after the name/serial columns, equally spaced
  int64_t p = span.atoms;
  int64_t e = line_end(text, nbytes, p);
  L.line_bytes = (int)(e - p) + 1;
  {
    int dots[3], nd = 0;
    for (int64_t i = p + 20; i < e && nd < 3; i++)
      if (text[i] == '.') dots[nd++] = (int)(i - p);
    if (nd < 3 || dots[1] - dots[0] != dots[2] - dots[1]) throw Unsupported("cannot find three equally spaced coordinate fields");
    L.field_w = dots[1] - dots[0];
    L.ndec = L.field_w - 5;
    L.x_col = dots[0] - (L.field_w - L.ndec - 1);
    if (L.ndec < 1 || L.ndec > 9 || L.x_col != 20 || L.x_col + 3 * L.field_w > L.line_bytes - 1) {
      throw Unsupported("coordinate columns are not in GROMACS fixed format (%5d%-5s%5s%5d then 3 x %W.Df)");
    }
  }

  // --- group ids, keys, frame-0 coordinates
  int last_mol = 0, g_minus = 0, g_plus = 0, id_env = 1, id_bcl = 1;
  std::vector<Tok> tok;
  for (int a = 0; a < n; a++) {
    int64_t le = line_end(text, nbytes, p);
    int len = (int)(le - p);
    if (len + 1 != L.line_bytes) {
      throw Unsupported("atom line " + std::to_string(a + 1) + " has a different length than the first one: not fixed-column");
    }
    const char* s = text + p;
    float fx[3];
    for (int d = 0; d < 3; d++) fx[d] = f32_atof(Tok{s + L.x_col + d * L.field_w, L.field_w});
    split_ws(s, len, tok);
    // The reference's tokeniser view of this line is usable when tokens 3..5 are the coordinates.
    bool tok_ok = tok.size() >= 6;
    if (tok_ok) {
      for (int d = 0; d < 3; d++) {
        float v = f32_atof(tok[3 + d]);
        if (memcmp(&v, &fx[d], 4) != 0) tok_ok = false;
      }
    }
    long resnr = 0;
    std::string mol, atom;
    if (tok_ok) {
      // leading digits of token 0 are the residue number, the rest its name; token 1 is the atom name (:409-415,446)
      int i = 0;
      while (i < tok[0].n && tok[0].p[i] >= '0' && tok[0].p[i] <= '9') i++;
      resnr = i ? atol(std::string(tok[0].p, i).c_str()) : 0;
      mol.assign(tok[0].p + i, tok[0].n - i);
      atom = tok[1].str();
    } else {
      // Extension beyond the reference (it aborts on such lines): strict fixed columns, e.g. 5-digit atom serials
      // that touch the atom name.
      std::vector<Tok> f;
      split_ws(s, 5, f);
      long v = 0;
      if (f.size() != 1 || !strict_int(f[0], v)) throw InputFileMalformed("atom line " + std::to_string(a + 1) + ": bad residue number");
      resnr = v;
      split_ws(s + 5, 5, f);
      mol = f.empty() ? std::string() : f[0].str();
      split_ws(s + 10, 5, f);
      atom = f.empty() ? std::string() : f[0].str();
    }
    int g, id;
    if (mol == "BCL" || mol == "BCX") {  // the reference is built with -DBCX_VALID (GNUmakefile:4; :417-421)
      mol = "BCX";
      if (resnr != last_mol) {
        g_minus--;
        id_bcl = 1;
      }
      g = g_minus;
      id = id_bcl++;
    } else {
      if (resnr != last_mol) g_plus++;
      g = g_plus;
      id = id_env++;
    }
    last_mol = (int)resnr;
    L.groups[a] = g;
    L.keys[a] = mol + "," + atom + std::to_string(id);  // (:460)
    for (int d = 0; d < 3; d++) L.x0[(size_t)3 * a + d] = fx[d];
    p = le + 1;
  }
  L.n_groups = g_plus;
  L.n_chromo = -g_minus;
}


// ---------------------------------------------------------------------------------------------------------------------
// .trr frame headers
// ---------------------------------------------------------------------------------------------------------------------
namespace {
inline uint32_t be32(const char* p) {
  const unsigned char* u = reinterpret_cast<const unsigned char*>(p);
  return ((uint32_t)u[0] << 24) | ((uint32_t)u[1] << 16) | ((uint32_t)u[2] << 8) | (uint32_t)u[3];
}
}  // namespace

bool locate_trr_frame(const char* data, int64_t nbytes, int64_t pos, TrrFrame& out, bool* truncated) {
  if (truncated) *truncated = false;
  if (pos >= nbytes) return false;
  auto cut = [&]() {
    if (truncated) *truncated = true;
    return false;
  };
  if (pos + 12 > nbytes) return cut();
  if (be32(data + pos) != 1993u) throw InputFileMalformed("not a .trr frame: magic number 1993 expected");
  const uint32_t slen = be32(data + pos + 4);     // strlen + 1 as written by GROMACS
  const uint32_t len = be32(data + pos + 8);      // XDR string length
  if (len > 64 || slen > 65) throw InputFileMalformed(".trr version string has an impossible length");
  int64_t p = pos + 12;
  const int64_t padded = ((int64_t)len + 3) & ~(int64_t)3;
  if (p + padded + 13 * 4 > nbytes) return cut();
  if (len != 12 || memcmp(data + p, "GMX_trn_file", 12) != 0) {
    throw InputFileMalformed(".trr version string is not \"GMX_trn_file\"");
  }
  p += padded;
  int32_t h[13];
  for (int i = 0; i < 13; i++) h[i] = (int32_t)be32(data + p + 4 * i);
  p += 13 * 4;
  const int32_t ir_size = h[0], e_size = h[1], box_size = h[2], vir_size = h[3], pres_size = h[4], top_size = h[5],
                sym_size = h[6], x_size = h[7], v_size = h[8], f_size = h[9], natoms = h[10], step = h[11];
  for (int i = 0; i < 10; i++)
    if (h[i] < 0) throw InputFileMalformed(".trr header holds a negative section size");
  if (natoms <= 0) throw InputFileMalformed(".trr frame with no atoms");
  int real_bytes = 0;
  if (box_size) real_bytes = box_size / 9;
  else if (x_size) real_bytes = x_size / (3 * natoms);
  else if (v_size) real_bytes = v_size / (3 * natoms);
  else if (f_size) real_bytes = f_size / (3 * natoms);
  if (real_bytes != 4 && real_bytes != 8) throw InputFileMalformed(".trr frame: cannot tell single from double precision");
  if (x_size == 0) throw Unsupported(".trr frame without coordinates (x_size == 0): write the trajectory with nstxout set");
  if ((int64_t)x_size != (int64_t)3 * natoms * real_bytes) throw InputFileMalformed(".trr header: x_size does not match natoms");
  if (p + 2 * real_bytes > nbytes) return cut();
  if (real_bytes == 4) {
    uint32_t b = be32(data + p);
    float t;
    memcpy(&t, &b, 4);
    out.time = t;
  } else {
    uint64_t b = ((uint64_t)be32(data + p) << 32) | be32(data + p + 4);
    double t;
    memcpy(&t, &b, 8);
    out.time = t;
  }
  p += 2 * real_bytes;  // t, lambda
  p += (int64_t)ir_size + e_size + box_size + vir_size + pres_size + top_size + sym_size;
  out.begin = pos;
  out.x_off = p;
  out.end = p + (int64_t)x_size + v_size + f_size;
  out.n_atoms = natoms;
  out.real_bytes = real_bytes;
  out.step = step;
  if (out.end > nbytes) return cut();
  return true;
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) static void copy_nt_avx2(char* dst, const char* src, size_t n) {
  size_t head = (32 - ((uintptr_t)dst & 31)) & 31;
  if (head > n) head = n;
  memcpy(dst, src, head);
  dst += head; src += head; n -= head;
  const size_t blocks = n / 128;
  for (size_t i = 0; i < blocks; i++) {
    const __m256i a = _mm256_loadu_si256((const __m256i*)(src + 128 * i));
    const __m256i b = _mm256_loadu_si256((const __m256i*)(src + 128 * i + 32));
    const __m256i c = _mm256_loadu_si256((const __m256i*)(src + 128 * i + 64));
    const __m256i d = _mm256_loadu_si256((const __m256i*)(src + 128 * i + 96));
    _mm256_stream_si256((__m256i*)(dst + 128 * i), a);
    _mm256_stream_si256((__m256i*)(dst + 128 * i + 32), b);
    _mm256_stream_si256((__m256i*)(dst + 128 * i + 64), c);
    _mm256_stream_si256((__m256i*)(dst + 128 * i + 96), d);
  }
  memcpy(dst + 128 * blocks, src + 128 * blocks, n - 128 * blocks);
  _mm_sfence();   // the copy engine reads the buffer next: the streamed lines must be globally visible
}
#endif

void stream_copy(char* dst, const char* src, size_t n, bool streaming) {
#if defined(__x86_64__)
  static const bool avx2 = __builtin_cpu_supports("avx2");
  if (streaming && avx2 && n >= 4096) {
    copy_nt_avx2(dst, src, n);
    return;
  }
#endif
  memcpy(dst, src, n);
}

}  // namespace cgc

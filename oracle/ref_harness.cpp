// TEST INFRASTRUCTURE — not product code.
//
// All-sites harness around the UNMODIFIED reference library (compiled from /root/reference/src/grom2atomFMO.cpp
// where it lies; see oracle/Makefile).  The stock CLI (reference src/FMOparse.cpp:133) hard-codes 7 sites and prints
// 6 significant digits; the library underneath has neither limit.  This driver repeats the call sequence of
// main()/main_Cr() (reference src/FMOparse.cpp:356-366, 89-95, 146, 150) for site = 1..nChromo and dumps raw binary:
//
//   <out>.meta.txt     "natoms nchromo ngroups numtot frames"
//   <out>.groups.i32   int32[natoms]        group ids of frame 0 (ReadOneTimeGrom2Atom output)
//   <out>.keys.txt     one "<mol>,<atom><ordinal>" key per line (frame 0)
//   <out>.x0.f32       float32[3*natoms]    coordinates of frame 0
//   <out>.q.f32        float32[nchromo][natoms]  AtomDataLookup_v2 charges per excited site (frame 0)
//   <out>.dE.f32       float32[frames][nchromo][numtot]  ComputeCDC_v1 output (kcal/mol), raw bits
//   <out>.timing.txt   "read_s lookup_s cdc_s total_s frames topo_s sites" wall-clock of the frame loop (no start-up)
//   <out>.frames.txt   one line per frame: "read_s lookup_s cdc_s"
//
// Usage: ref_harness traj top itp out num_frames [max_sites [dump]]
//   max_sites > 0 evaluates only sites 1..max_sites of every frame (a bounded sample for CPU timing: sites are
//   independent and equally expensive); dump = 0 skips the binary dumps.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#include <algorithm>
#include <cmath>

#include "grom2atomFMO.h"

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

template <typename T>
static void dump(const std::string& path, const T* p, size_t n, const char* mode) {
  FILE* f = fopen(path.c_str(), mode);
  if (!f) { perror(path.c_str()); exit(2); }
  fwrite(p, sizeof(T), n, f);
  fclose(f);
}

int main(int argc, char** argv) {
  if (argc < 6) {
    fprintf(stderr, "usage: %s traj top itp out num_frames\n", argv[0]);
    return 1;
  }
  std::string traj = argv[1], top = argv[2], itp = argv[3], out = argv[4];
  int num_frames = atoi(argv[5]);
  int max_sites = argc > 6 ? atoi(argv[6]) : 0;
  bool do_dump = argc > 7 ? atoi(argv[7]) != 0 : true;

  std::vector<std::string> names;
  std::vector<float> masses, charges;
  double t_start = now_s();
  try { ParseTopology(itp, names, masses, charges, Excited::YES); } catch (InputFileMinorMalformed&) {}
  try { ParseTopology(top, names, masses, charges); } catch (InputFileMinorMalformed&) {}
  double t_topo = now_s() - t_start;

  std::streampos cur, next;
  bool have_pos = false;
  double t_read = 0, t_lookup = 0, t_cdc = 0;
  int nchromo = 0, ngroups = 0, numtot = 0, natoms = 0, done = 0;
  dump<float>(out + ".dE.f32", nullptr, 0, "wb");
  std::ofstream per_frame(out + ".frames.txt");
  int sites_done = 0;
  double t_loop0 = now_s();
  for (int t = 0; t < num_frames; t++) {
    std::vector<float> Xi;
    std::vector<std::string> types;
    std::vector<int> groups;
    float box = 0;
    double t0 = now_s();
    ReadOneTimeGrom2Atom(have_pos ? &cur : 0, next, Xi, types, groups, box, traj);
    cur = next; have_pos = true;
    double f_read = now_s() - t0, f_lookup = 0, f_cdc = 0;
    t_read += f_read;
    if (groups.empty()) break;  // ran off the end of the file (the stock CLI segfaults here)
    ngroups = *std::max_element(groups.begin(), groups.end());
    nchromo = std::abs(*std::min_element(groups.begin(), groups.end()));
    numtot = ngroups + nchromo;
    natoms = (int)groups.size();
    if (t == 0 && do_dump) {
      dump<int>(out + ".groups.i32", groups.data(), groups.size(), "wb");
      dump<float>(out + ".x0.f32", Xi.data(), Xi.size(), "wb");
      std::ofstream k(out + ".keys.txt");
      for (auto& s : types) k << s << "\n";
      dump<float>(out + ".q.f32", nullptr, 0, "wb");
    }
    sites_done = (max_sites > 0 && max_sites < nchromo) ? max_sites : nchromo;
    for (int site = 1; site <= sites_done; site++) {
      std::vector<float> m, sz, q, ground, cdc;
      t0 = now_s();
      AtomDataLookup_v2(types, m, sz, q, groups, site, ground, names, masses, charges);
      f_lookup += now_s() - t0;
      if (t == 0 && do_dump) dump<float>(out + ".q.f32", q.data(), q.size(), "ab");
      t0 = now_s();
      ComputeCDC_v1(site, cdc, Xi, q, groups, ground);
      f_cdc += now_s() - t0;
      if (do_dump) dump<float>(out + ".dE.f32", cdc.data(), cdc.size(), "ab");
    }
    t_lookup += f_lookup;
    t_cdc += f_cdc;
    per_frame << f_read << " " << f_lookup << " " << f_cdc << "\n";
    done++;
  }
  double t_loop = now_s() - t_loop0;
  {
    std::ofstream m(out + ".meta.txt");
    m << natoms << " " << nchromo << " " << ngroups << " " << numtot << " " << done << "\n";
    std::ofstream tm(out + ".timing.txt");
    tm << t_read << " " << t_lookup << " " << t_cdc << " " << t_loop << " " << done << " " << t_topo << " " << sites_done
       << "\n";
  }
  return 0;
}

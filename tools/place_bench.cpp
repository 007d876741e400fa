// place_bench.cpp -- offline driver for the shared-memory placement search (csrc/pbd_placement.cpp) on ONE
// tile problem dumped by the planner:
//   PBD_PLACE_DUMP=/tmp/tile.bin:5 python tools/plan_report.py 56      # the 6th tile the planner places
//   g++ -O2 -std=c++17 tools/place_bench.cpp cs121-softbodysim_b200/csrc/pbd_placement.cpp -o /tmp/place_bench
//   /tmp/place_bench /tmp/tile.bin [effort] [block]
// Prints the gathers' wavefronts per quarter-warp role before / after and the time the search took.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../cs121-softbodysim_b200/csrc/pbd_plan.h"

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: place_bench dump.bin [effort] [block] [relabel]\n"); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 1; }
  uint32_t hdr[3];
  if (fread(hdr, 4, 3, f) != 3) return 1;
  std::vector<pbd::PlaceGroup> groups(hdr[1]);
  std::vector<uint32_t> loc((size_t)hdr[2] * 4), payload(hdr[2]);
  if (fread(groups.data(), sizeof(pbd::PlaceGroup), hdr[1], f) != hdr[1] || fread(loc.data(), 16, hdr[2], f) != hdr[2]) return 1;
  fclose(f);
  for (uint32_t i = 0; i < hdr[2]; ++i) payload[i] = i;
  const int effort = argc > 2 ? atoi(argv[2]) : 1;
  const uint32_t block = argc > 3 ? (uint32_t)atoi(argv[3]) : 0u;
  const bool relabel = argc > 4 && atoi(argv[4]) != 0;   // tets may permute their vertices (fast arithmetic)
  printf("tile: %u vertices, %u groups, %u constraints\n", hdr[0], hdr[1], hdr[2]);
  for (int e : {0, effort}) {
    std::vector<uint32_t> l = loc, p = payload, nl;
    std::vector<uint8_t> pm(hdr[2]);
    pbd::PlaceStats st;
    const auto t0 = std::chrono::steady_clock::now();
    pbd::optimise_placement(hdr[0], groups.data(), hdr[1], l.data(), p.data(), hdr[2], e, block, nl, &st, relabel && e ? pm.data() : nullptr);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    printf("effort %d: edges %.3f (%llu / %llu)  tets %.3f (%llu / %llu)   %.1f ms\n", e, (double)st.wavefronts[0] / std::max<uint64_t>(1, st.ideal[0]),
           (unsigned long long)st.wavefronts[0], (unsigned long long)st.ideal[0], (double)st.wavefronts[1] / std::max<uint64_t>(1, st.ideal[1]),
           (unsigned long long)st.wavefronts[1], (unsigned long long)st.ideal[1], ms);
  }
  return 0;
}

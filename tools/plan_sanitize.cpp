// plan_sanitize.cpp -- host-only driver for the schedule planner (csrc/pbd_plan.cpp + pbd_tileplan.cpp,
// ~2,000 lines of C++ that use up to 16 threads) to be built with -fsanitize=address,undefined and
// -fsanitize=thread (tests/test_sanitizers_cpu.py).  No CUDA: it calls what pbd_plan_create calls.
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=all -pthread \
//       tools/plan_sanitize.cpp cs121-softbodysim_b200/csrc/pbd_plan.cpp cs121-softbodysim_b200/csrc/pbd_tileplan.cpp -o plan_asan
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <cstring>
#include <string>
#include <vector>

#include "../cs121-softbodysim_b200/csrc/pbd_plan.h"

using namespace pbd;

// Kuhn / Freudenthal 6-tet split of an n^3 cube; edges = unique tet edges, first-seen order
static void kuhn(uint32_t n, std::vector<float>& x, std::vector<uint32_t>& tets, std::vector<uint32_t>& edges) {
  const uint32_t m = n + 1;
  x.resize((size_t)m * m * m * 3);
  for (uint32_t k = 0; k < m; ++k)
    for (uint32_t j = 0; j < m; ++j)
      for (uint32_t i = 0; i < m; ++i) {
        const size_t v = ((size_t)k * m + j) * m + i;
        x[3 * v] = (float)i / n + 0.013f * k; x[3 * v + 1] = 0.25f + (float)j / n; x[3 * v + 2] = (float)k / n - 0.007f * i;
      }
  static const int perms[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
  for (uint32_t k = 0; k < n; ++k)
    for (uint32_t j = 0; j < n; ++j)
      for (uint32_t i = 0; i < n; ++i)
        for (auto& p : perms) {
          uint32_t c[3] = {i, j, k}, id[4];
          id[0] = (c[2] * m + c[1]) * m + c[0];
          for (int s = 0; s < 3; ++s) { c[p[s]]++; id[s + 1] = (c[2] * m + c[1]) * m + c[0]; }
          tets.insert(tets.end(), id, id + 4);
        }
  std::vector<std::vector<uint32_t>> seen((size_t)m * m * m);
  static const int pr[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
  for (size_t t = 0; t < tets.size() / 4; ++t)
    for (auto& q : pr) {
      uint32_t a = tets[4 * t + q[0]], b = tets[4 * t + q[1]];
      if (a > b) std::swap(a, b);
      bool dup = false;
      for (uint32_t o : seen[a]) dup |= o == b;
      if (!dup) { seen[a].push_back(b); edges.push_back(a); edges.push_back(b); }
    }
}

int main(int argc, char** argv) {
  const uint32_t n = argc > 1 ? (uint32_t)atoi(argv[1]) : 14;
  std::vector<float> x;
  std::vector<uint32_t> tets, edges;
  kuhn(n, x, tets, edges);
  MeshView m{(uint32_t)(x.size() / 3), (uint32_t)(edges.size() / 2), (uint32_t)(tets.size() / 4), x.data(), edges.data(), tets.data()};
  std::string err;
  if (!validate_mesh(m, err)) { printf("validate: %s\n", err.c_str()); return 1; }
  int rc = 0;
  struct Case { uint32_t order, tileVerts, partitions, lanes, flags; };
  const Case cases[] = {{PBD_ORDER_STRICT, 0, 0, 0, 0}, {PBD_ORDER_INTERLEAVED, 0, 0, 0, 0}, {PBD_ORDER_RIDING, 0, 0, 0, 0},
                        {PBD_ORDER_RIDING, 150, 3, 0, PBD_FLAG_TAGGED_HANDOVER}, {PBD_ORDER_INTERLEAVED, 90, 5, 4, 0},
                        {PBD_ORDER_STRICT, 64, 2, 2, 0},
                        // the placement search with relabelled tets (fast arithmetic), home and shifted tiles
                        {PBD_ORDER_INTERLEAVED, 0, 0, 0, PBD_FLAG_TAGGED_HANDOVER | PBD_FLAG_FAST_ARITH},
                        {PBD_ORDER_INTERLEAVED, 120, 0, 1, PBD_FLAG_TAGGED_HANDOVER | PBD_FLAG_FAST_ARITH}};
  for (const Case& c : cases) {
    pbd_options o;
    memset(&o, 0, sizeof o);
    o.struct_size = sizeof o; o.backend = PBD_BACKEND_TILE; o.order_mode = c.order; o.tile_vertices = c.tileVerts;
    o.partitions = c.partitions; o.lanes_per_tet = c.lanes; o.flags = c.flags;
    Plan plan;
    if (!build_tile_plan(m, o, 148, 227u * 1024u - 8192u, plan, err)) { printf("plan failed: %s\n", err.c_str()); rc = 1; continue; }
    // every constraint scheduled exactly once
    std::vector<uint8_t> seenE(m.E, 0), seenT(m.T, 0);
    for (uint32_t e : plan.edgeOrder) seenE[e]++;
    for (uint32_t t : plan.tetOrder) seenT[t]++;
    for (uint8_t s : seenE) rc |= s != 1;
    for (uint8_t s : seenT) rc |= s != 1;
    // a relabelled tet still names its own four vertices, each once (tetLocal is in role order)
    if (!plan.tetPerm.empty())
      for (uint32_t k = 0; k < m.T; ++k) {
        uint32_t used = 0;
        for (int r = 0; r < 4; ++r) used |= 1u << ((plan.tetPerm[k] >> (2 * r)) & 3u);
        rc |= used != 15u;
      }
    printf("order %u tv %u K %u lanes %u flags %u: %zu tiles, %zu phases, %u riders, %.0f ms\n", c.order, c.tileVerts, c.partitions, c.lanes,
           c.flags, plan.tiles.size(), plan.phases.size(), plan.riders, plan.planMs);
  }
  Plan sp;
  build_stream_plan(m, sp);
  std::vector<float> w, er, tr;
  host_inverse_mass(m, nullptr, 0, w);
  host_rest_state(m, er, tr);
  // NaN / Inf input must be refused before it reaches the planner's sorts
  std::vector<float> bad = x;
  bad[7] = std::numeric_limits<float>::quiet_NaN();
  MeshView mb = m;
  mb.x0 = bad.data();
  rc |= validate_mesh(mb, err) ? 1 : 0;
  printf(rc ? "FAILED\n" : "OK\n");
  return rc;
}

// pbd_plan.cpp -- colouring, validation, host init helpers, the "stream" schedule.
#include "pbd_plan.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>

namespace pbd {

namespace {
double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}
}  // namespace

uint32_t greedy_colour(const uint32_t* ids, uint32_t n, uint32_t arity, uint32_t nVerts,
                       std::vector<uint32_t>& colour) {
  colour.assign(n, 0);
  if (n == 0) return 0;
  // Per-vertex bitmask of colours already used by incident constraints; the word count grows
  // on demand (high-valence vertices of Delaunay-style meshes need > 64 colours).
  uint32_t words = 1;
  std::vector<uint64_t> used((size_t)nVerts * words, 0);
  uint32_t nColours = 0;
  for (uint32_t k = 0; k < n; ++k) {
    const uint32_t* c = ids + (size_t)k * arity;
    uint32_t pick = UINT32_MAX;
    for (;;) {
      for (uint32_t wd = 0; wd < words && pick == UINT32_MAX; ++wd) {
        uint64_t m = 0;
        for (uint32_t j = 0; j < arity; ++j) m |= used[(size_t)c[j] * words + wd];
        if (~m) pick = wd * 64 + (uint32_t)__builtin_ctzll(~m);
      }
      if (pick != UINT32_MAX) break;
      // every colour in `words` words is taken at these vertices: widen the masks
      uint32_t nw = words * 2;
      std::vector<uint64_t> wider((size_t)nVerts * nw, 0);
      for (uint32_t v = 0; v < nVerts; ++v)
        std::memcpy(&wider[(size_t)v * nw], &used[(size_t)v * words], sizeof(uint64_t) * words);
      used.swap(wider);
      words = nw;
    }
    colour[k] = pick;
    nColours = std::max(nColours, pick + 1);
    for (uint32_t j = 0; j < arity; ++j) used[(size_t)c[j] * words + pick / 64] |= 1ull << (pick % 64);
  }
  return nColours;
}

bool validate_mesh(const MeshView& m, std::string& err) {
  if (m.V > 0 && !m.x0) { err = "x0 is null"; return false; }
  if (m.E > 0 && !m.edges) { err = "edgeIds is null"; return false; }
  if (m.T > 0 && !m.tets) { err = "tetIds is null"; return false; }
  // NaN / Inf positions would only propagate in the reference; here they would reach the planner's
  // sorts (not a strict weak ordering with NaN) and llround() of the bounding box: refuse them
  for (size_t i = 0; i < (size_t)m.V * 3; ++i)
    if (!std::isfinite(m.x0[i])) { err = "x0 is not finite (vertex " + std::to_string(i / 3) + ")"; return false; }
  for (size_t i = 0; i < (size_t)m.E * 2; ++i)
    if (m.edges[i] >= m.V) { err = "edge index out of range (edge " + std::to_string(i / 2) + ")"; return false; }
  for (size_t i = 0; i < (size_t)m.T * 4; ++i)
    if (m.tets[i] >= m.V) { err = "tet index out of range (tet " + std::to_string(i / 4) + ")"; return false; }
  return true;
}

// counting sort of constraints by colour, stable: inside a colour the caller's order is kept
static void order_by_colour(const std::vector<uint32_t>& colour, uint32_t nColours,
                            std::vector<uint32_t>& order, std::vector<uint32_t>& off) {
  const uint32_t n = (uint32_t)colour.size();
  off.assign((size_t)nColours + 1, 0);
  for (uint32_t k = 0; k < n; ++k) off[colour[k] + 1]++;
  for (uint32_t c = 0; c < nColours; ++c) off[c + 1] += off[c];
  order.resize(n);
  std::vector<uint32_t> cur(off.begin(), off.end() - 1);
  for (uint32_t k = 0; k < n; ++k) order[cur[colour[k]]++] = k;
}

void build_stream_plan(const MeshView& m, Plan& p) {
  double t0 = now_ms();
  p.V = m.V; p.E = m.E; p.T = m.T;
  p.backend = PBD_BACKEND_STREAM;
  p.orderMode = PBD_ORDER_STRICT;
  uint32_t ne = greedy_colour(m.edges, m.E, 2, m.V, p.edgeColor);
  uint32_t nt = greedy_colour(m.tets, m.T, 4, m.V, p.tetColor);
  order_by_colour(p.edgeColor, ne, p.edgeOrder, p.edgeColorOff);
  order_by_colour(p.tetColor, nt, p.tetOrder, p.tetColorOff);
  p.edgePhase.assign(m.E, 0); p.edgeTile.assign(m.E, 0);
  p.tetPhase.assign(m.T, 0); p.tetTile.assign(m.T, 0);
  p.edgePhases = ne; p.tetPhases = nt;       // one grid-wide phase (= launch) per colour
  p.edgeColorSum = ne; p.tetColorSum = nt;
  p.edgeDev.resize(m.E); p.tetDev.resize(m.T);
  for (uint32_t k = 0; k < m.E; ++k) p.edgeDev[k] = k;
  for (uint32_t k = 0; k < m.T; ++k) p.tetDev[k] = k;
  p.edgeDevCount = m.E; p.tetDevCount = m.T;
  p.slotToVertex.resize(m.V); p.vertexToSlot.resize(m.V);
  for (uint32_t i = 0; i < m.V; ++i) p.slotToVertex[i] = p.vertexToSlot[i] = i;
  p.planMs = now_ms() - t0;
}

// ---- reference init helpers, restated for the host ---------------------------------------
// Compiled for baseline x86-64 without FMA contraction (build.py passes -ffp-contract=off to the
// host compiler), the same arithmetic as the reference's g++ -O3 build.

static inline float signed_volume6(const float* p0, const float* p1, const float* p2, const float* p3) {
  // tet_volume, CProgram/include/PBDServer.h:140-145
  float ax = p1[0] - p0[0], ay = p1[1] - p0[1], az = p1[2] - p0[2];
  float bx = p2[0] - p0[0], by = p2[1] - p0[1], bz = p2[2] - p0[2];
  float cx = p3[0] - p0[0], cy = p3[1] - p0[1], cz = p3[2] - p0[2];
  float nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
  return (nx * cx + ny * cy + nz * cz) / 6.0f;
}

void host_inverse_mass(const MeshView& m, const uint32_t* pinned, uint32_t nPinned, std::vector<float>& w) {
  w.assign(m.V, 0.0f);
  for (uint32_t t = 0; t < m.T; ++t) {
    const uint32_t* id = m.tets + (size_t)t * 4;
    float vol = signed_volume6(m.x0 + 3 * (size_t)id[0], m.x0 + 3 * (size_t)id[1],
                               m.x0 + 3 * (size_t)id[2], m.x0 + 3 * (size_t)id[3]);
    float mv = std::fabs(vol);
    if (mv > 1e-12f) {
      float inv = 4.0f / mv;
      w[id[0]] += inv; w[id[1]] += inv; w[id[2]] += inv; w[id[3]] += inv;
    }
  }
  for (uint32_t k = 0; k < nPinned; ++k)
    if (pinned[k] < m.V) w[pinned[k]] = 0.0f;
}

void host_rest_state(const MeshView& m, std::vector<float>& edgeRest, std::vector<float>& tetRest) {
  edgeRest.resize(m.E);
  for (uint32_t e = 0; e < m.E; ++e) {
    const float* p0 = m.x0 + 3 * (size_t)m.edges[2 * (size_t)e];
    const float* p1 = m.x0 + 3 * (size_t)m.edges[2 * (size_t)e + 1];
    float dx = p1[0] - p0[0], dy = p1[1] - p0[1], dz = p1[2] - p0[2];
    edgeRest[e] = std::sqrt(dx * dx + dy * dy + dz * dz);
  }
  tetRest.resize(m.T);
  for (uint32_t t = 0; t < m.T; ++t) {
    const uint32_t* id = m.tets + (size_t)t * 4;
    tetRest[t] = signed_volume6(m.x0 + 3 * (size_t)id[0], m.x0 + 3 * (size_t)id[1],
                                m.x0 + 3 * (size_t)id[2], m.x0 + 3 * (size_t)id[3]);
  }
}

uint64_t algorithmic_bytes_per_substep(uint32_t V, uint32_t E, uint32_t T, uint32_t iterations) {
  // predict 52 B/vertex + commit 52 B/vertex; per iteration the edge sweep
  // 20 B/edge + 28 B/vertex, the tet sweep 28 B/tet + 28 B/vertex, the ground clamp 28 B/vertex.
  return 104ull * V + (uint64_t)iterations * (20ull * E + 28ull * T + 84ull * V);
}

}  // namespace pbd

// pbd_tileplan.cpp -- the "tile" schedule: shifted shared-memory tiles.
//
// Why: a globally coloured sweep needs one grid-wide barrier per colour (~46 per iteration on the
// Kuhn grid, 155 on default_Tet), which caps an L2-resident mesh at a few % of the HBM roofline
// (SURVEY.md 7 "dependent-phase count vs. bytes").  Here one iteration is K grid-wide phases
// (K = 4 by default; the phases are ordered by per-tile dependency counters, not by a grid barrier,
// see pbd_tile.cu).  Phase p uses vertex partition P_p: the vertices are cut into as many tiles
// as there are SMs (k-d style multi-way splits in rank space), and P_1..P_{K-1} are the same cuts
// CYCLICALLY SHIFTED by p/K of a tile along every axis.  A constraint can run in phase p when all
// its vertices fall into one tile of P_p ("interior"); with four shifts every constraint of a
// grid-like mesh is interior to at least one partition, usually to several.  Each constraint is
// then assigned to ONE of its admissible phases so that every vertex sees about 1/K of its
// incident constraints per phase: the dependent chain of an iteration (sum over phases of the
// local colour count) stays close to the vertex valence instead of multiplying with the phases.
// Inside a (phase, tile) the constraints are coloured locally and swept in shared memory with
// block barriers only.  Constraints interior to no partition (rare) go through residual phases
// that re-partition just their own vertices (graph Voronoi), as many as needed.
//
// Order: phase by phase, tile by tile, [edge colours, then tet colours]; inside a colour the
// constraints are arranged for conflict-free shared-memory banks (any order gives the same result:
// they share no vertex).  Tiles of one phase are vertex-disjoint and colour groups are conflict-free, so running
// them in parallel equals running them sequentially in that order: still Gauss-Seidel over the
// same constraint set, only permuted (disclosed through pbd_get_schedule_order / _sequence).
//   PBD_ORDER_STRICT       every iteration projects all edges (K phases), then all tets (K phases),
//                          like the reference (CProgram/src/Sim.cpp:293-297) up to a permutation
//                          inside each type: the unmodified reference reproduces it bit for bit.
//   PBD_ORDER_INTERLEAVED  a tile visit projects its edges and then its tets (K phases per
//                          iteration instead of 2K); checked against the oracle's sequence sweep.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <thread>

#include "pbd_plan.h"

namespace pbd {

namespace {

constexpr uint32_t NONE = 0xffffffffu;
constexpr uint32_t kMaxPartitions = 8;

// Debug switches of the planner, read from the environment ONCE (A/B runs: tools/gpu_ab_env.sh).
// None of them is needed in production; the defaults are what every measurement in DESIGN.md uses.
struct PlanKnobs {
  bool debug;            // PBD_PLAN_DEBUG=1          stage timings and statistics on stderr
  bool mixed;            // PBD_PLAN_MIXED=0          interleaved order without mixed colour steps
  bool potential;        // PBD_PLAN_NOPOT=1          skip the potential-descent pass of the assignment
  bool repair;           // PBD_PLAN_NOREPAIR=1       skip the per-type peak repair
  bool jointRepair;      // PBD_PLAN_NOJOINTREPAIR=1  skip the joint (edges + tets) peak repair
  bool jointLoad;        // PBD_PLAN_NOJOINTLOAD=1    balance the tets without the edge loads
  bool tileBalance;      // PBD_PLAN_NOBAL=1          skip the tile balance
  int snapLevels;        // PBD_PLAN_NOSNAP=n         k-d levels whose cuts are NOT gap-snapped (default 4)
  int capMargin;         // PBD_PLAN_CAPM=n           tile balance: cap = p99 load - n (default 1)
  int tabu;              // PBD_PLAN_TABU=n           tabu-search iterations per class (-1: built-in budgets)
  int minTile;           // PBD_PLAN_MINTILE=n        smallest tile (vertices) before fewer SMs are used instead (default 1024)
  int riders;            // PBD_PLAN_RIDERS=n         PBD_ORDER_RIDING: most riders per tet (1 or 2, default 2)
  bool sigSort;          // PBD_PLAN_NOSIGSORT=1      home-tile slots in caller order instead of sorted by their shifted tiles
  bool relabel;          // PBD_PLAN_NORELABEL=1      fast arithmetic: keep every tet's vertices in their roles
  int place;             // PBD_PLAN_PLACE=n          shared-memory placement search (pbd_placement.cpp): 0 = off, default 1
  int placeBlockHome;    // PBD_PLAN_PLACE_BH=n       ... a vertex of a home tile moves inside its aligned block of n indices (0: anywhere)
  int placeBlockShifted; // PBD_PLAN_PLACE_BS=n       ... the same for the shifted tiles
};
const PlanKnobs& knobs() {
  static const PlanKnobs k = [] {
    auto flag = [](const char* n) { const char* e = getenv(n); return e && atoi(e) != 0; };
    auto num = [](const char* n, int dflt) { const char* e = getenv(n); return e ? atoi(e) : dflt; };
    PlanKnobs q;
    q.debug = getenv("PBD_PLAN_DEBUG") != nullptr;
    q.mixed = num("PBD_PLAN_MIXED", 1) != 0;
    q.potential = !flag("PBD_PLAN_NOPOT");
    q.repair = !flag("PBD_PLAN_NOREPAIR");
    q.jointRepair = !flag("PBD_PLAN_NOJOINTREPAIR");
    q.jointLoad = !flag("PBD_PLAN_NOJOINTLOAD");
    q.tileBalance = !flag("PBD_PLAN_NOBAL");
    q.snapLevels = num("PBD_PLAN_NOSNAP", 4);
    q.capMargin = num("PBD_PLAN_CAPM", 1);
    q.tabu = num("PBD_PLAN_TABU", -1);
    q.minTile = std::max(32, num("PBD_PLAN_MINTILE", 1024));
    q.riders = num("PBD_PLAN_RIDERS", 2);
    q.sigSort = !flag("PBD_PLAN_NOSIGSORT");
    q.relabel = !flag("PBD_PLAN_NORELABEL");
    q.place = std::max(0, num("PBD_PLAN_PLACE", 1));
    q.placeBlockHome = std::max(0, num("PBD_PLAN_PLACE_BH", 8));
    q.placeBlockShifted = std::max(0, num("PBD_PLAN_PLACE_BS", 32));
    return q;
  }();
  return k;
}

double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// ------------------------------------------------------------------ shifted k-d partitions

// Multi-way k-d split in rank space.  Level L sorts the current vertex set along the L-th longest
// axis of the body frame and cuts it into parts that receive equal shares of the tiles; `offset`
// (a fraction of one part) rotates the cuts cyclically, so the last part wraps around to the first.
// Every cut is then snapped to the widest coordinate gap nearby: on lattice-like meshes the cuts
// fall BETWEEN vertex planes (and the shifted partitions' cuts between other planes), on
// unstructured meshes the snap is a no-op in effect.
static int nosnapLevels() {
  static const int v = knobs().snapLevels;
  return v;
}

struct Partitioner {
  const float* x = nullptr;         // 3V body-frame coordinates
  int axis[3] = {0, 1, 2};
  double ext[3] = {1, 1, 1};
  double offset = 0.0;
  uint32_t grid[3] = {0, 0, 0};     // parts per level when the tile count is a product m0*m1*m2 (0: derive)
  std::vector<uint32_t> idx;        // permutation being partitioned
  std::vector<uint32_t> tileBegin;  // leaf ranges in idx
  uint32_t* tileOf = nullptr;       // per caller vertex

  void leaf(uint32_t lo, uint32_t hi) {
    const uint32_t t = (uint32_t)tileBegin.size();
    tileBegin.push_back(lo);
    std::sort(idx.begin() + lo, idx.begin() + hi);   // caller order inside a tile
    for (uint32_t i = lo; i < hi; ++i) tileOf[idx[i]] = t;
  }

  void split(uint32_t lo, uint32_t hi, uint32_t nTiles, int level) {
    const uint32_t n = hi - lo;
    if (n == 0) return;
    if (nTiles <= 1 || n <= 1 || level > 2) { leaf(lo, hi); return; }
    double want;
    if (level == 2) want = nTiles;
    else if (level == 1) want = std::sqrt((double)nTiles * ext[1] / ext[2]);
    else want = std::cbrt((double)nTiles * ext[0] * ext[0] / (ext[1] * ext[2]));
    uint32_t parts = grid[level] ? grid[level] : (uint32_t)std::llround(want);
    parts = std::max(1u, std::min(parts, std::min(nTiles, n)));
    if (parts == 1 && level < 2) { split(lo, hi, nTiles, level + 1); return; }
    const int ax = axis[level];
    auto coord = [&](uint32_t i) { return x[3 * (size_t)idx[lo + i] + ax]; };
    const float q = (nosnapLevels() >> level & 1) ? (float)(ext[level] * 1e-4) : 0.0f;
    std::sort(idx.begin() + lo, idx.begin() + hi, [&](uint32_t a, uint32_t b) {
      float ca = x[3 * (size_t)a + ax], cb = x[3 * (size_t)b + ax];
      if (q > 0.0f) { ca = std::floor(ca / q + 0.5f); cb = std::floor(cb / q + 0.5f); }
      return ca < cb || (ca == cb && a < b);   // total order -> the split is unique
    });
    // boundaries in sorted-rank space (cyclic), snapped to the widest gap within +-window
    const uint32_t shift = (uint32_t)((offset * n) / parts) % n;
    // The last level is cut at the exact rank instead (coordinates quantised so that a lattice plane
    // splits along vertex-index order, i.e. along a line): tiles of one column then hold the same
    // number of vertices, which evens out the tiles' work (measured +5 %), while the snapped first
    // two levels keep the cuts of all slabs and columns aligned.
    const int nosnap = nosnapLevels();   // bit L = level L unsnapped
    const uint32_t window = (nosnap >> level & 1) ? 0u : n / parts / 8;
    std::vector<uint32_t> cut(parts), tiles(parts);
    uint32_t cum = 0;
    for (uint32_t j = 0; j < parts; ++j) {
      tiles[j] = nTiles / parts + (j < nTiles % parts ? 1u : 0u);
      uint32_t b = (shift + (uint32_t)(((uint64_t)n * cum) / nTiles)) % n;
      cum += tiles[j];
      if (window && b != 0) {
        const uint32_t from = b > window ? b - window : 1u, to = std::min(n - 1, b + window);
        float best = coord(b) - coord(b - 1);
        uint32_t bestR = b;
        for (uint32_t r = from; r <= to; ++r) {
          const float gap = coord(r) - coord(r - 1);
          const uint32_t dr = r > b ? r - b : b - r, db = bestR > b ? bestR - b : b - bestR;
          if (gap > best * 1.0001f || (gap >= best && dr < db)) { best = gap; bestR = r; }
        }
        b = bestR;
      }
      cut[j] = b;
    }
    // parts are the cyclic ranges [cut[j], cut[j+1]); rotate so that part 0 starts the array
    for (uint32_t j = 1; j < parts; ++j)   // keep the cuts strictly increasing in cyclic order
      if (((cut[j] + n - cut[0]) % n) <= ((cut[j - 1] + n - cut[0]) % n)) cut[j] = (cut[j - 1] + 1) % n;
    const uint32_t c0 = cut[0];
    if (c0) std::rotate(idx.begin() + lo, idx.begin() + lo + c0, idx.begin() + hi);
    for (uint32_t j = 0; j < parts; ++j) {
      const uint32_t b = (cut[j] + n - c0) % n;
      const uint32_t e = j + 1 < parts ? (cut[j + 1] + n - c0) % n : n;
      if (e > b) split(lo + b, lo + e, tiles[j], level + 1);
    }
  }
};

// Body frame: the rotation that minimises the volume of the axis-aligned bounding box (searched
// over Euler angles on the extreme points of the body).  For box-like bodies the k-d cuts then
// run parallel to the faces, so the cuts of different slabs line up and the shifted partitions
// stay a fixed distance apart.  Returns row-major R (frame coordinates = R * x); identity unless
// the box shrinks by more than 2 %.
void body_frame(const float* x, uint32_t V, double R[9]) {
  for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
  if (V < 8) return;
  // extreme points along 96 fixed directions (Fibonacci sphere)
  std::vector<uint32_t> ext;
  const int D = 96;
  for (int d = 0; d < D; ++d) {
    const double z = 1.0 - (2.0 * d + 1.0) / D, r = std::sqrt(std::max(0.0, 1.0 - z * z)), ph = d * 2.399963229728653;
    const double dx = r * std::cos(ph), dy = r * std::sin(ph), dz = z;
    uint32_t lo = 0, hi = 0;
    double vlo = INFINITY, vhi = -INFINITY;
    for (uint32_t v = 0; v < V; ++v) {
      const double t = dx * x[3 * (size_t)v] + dy * x[3 * (size_t)v + 1] + dz * x[3 * (size_t)v + 2];
      if (t < vlo) { vlo = t; lo = v; }
      if (t > vhi) { vhi = t; hi = v; }
    }
    ext.push_back(lo);
    ext.push_back(hi);
  }
  std::sort(ext.begin(), ext.end());
  ext.erase(std::unique(ext.begin(), ext.end()), ext.end());
  auto rot = [](double a, double b, double c, double M[9]) {   // Rz(c) * Ry(b) * Rx(a)
    const double ca = std::cos(a), sa = std::sin(a), cb = std::cos(b), sb = std::sin(b), cc = std::cos(c), sc = std::sin(c);
    M[0] = cc * cb; M[1] = cc * sb * sa - sc * ca; M[2] = cc * sb * ca + sc * sa;
    M[3] = sc * cb; M[4] = sc * sb * sa + cc * ca; M[5] = sc * sb * ca - cc * sa;
    M[6] = -sb;     M[7] = cb * sa;                M[8] = cb * ca;
  };
  auto volume = [&](const double M[9]) {
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t v : ext)
      for (int i = 0; i < 3; ++i) {
        const double t = M[3 * i] * x[3 * (size_t)v] + M[3 * i + 1] * x[3 * (size_t)v + 1] + M[3 * i + 2] * x[3 * (size_t)v + 2];
        mn[i] = std::min(mn[i], t);
        mx[i] = std::max(mx[i], t);
      }
    return (mx[0] - mn[0]) * (mx[1] - mn[1]) * (mx[2] - mn[2]);
  };
  double I[9];
  rot(0, 0, 0, I);
  const double v0 = volume(I);
  if (!(v0 > 0.0) || !std::isfinite(v0)) return;
  const double H = 1.5707963267948966;
  double best[3] = {0, 0, 0}, bestV = v0, M[9];
  const int G = 15;
  for (int i = 0; i < G; ++i)
    for (int j = 0; j < G; ++j)
      for (int k = 0; k < G; ++k) {
        const double a = H * i / G, b = H * j / G - H / 2, c = H * k / G;
        rot(a, b, c, M);
        const double v = volume(M);
        if (v < bestV) { bestV = v; best[0] = a; best[1] = b; best[2] = c; }
      }
  for (double step = H / G / 2; step > 1e-5; step *= 0.5)
    for (int pass = 0; pass < 2; ++pass)
      for (int q = 0; q < 3; ++q)
        for (int sgn = -1; sgn <= 1; sgn += 2) {
          double t[3] = {best[0], best[1], best[2]};
          t[q] += sgn * step;
          rot(t[0], t[1], t[2], M);
          const double v = volume(M);
          if (v < bestV) { bestV = v; best[q] = t[q]; }
        }
  if (bestV < 0.98 * v0) rot(best[0], best[1], best[2], R);
}

// ------------------------------------------------------------------ residual re-partitioning

// Constraints of one type in SLOT numbering.
struct CSet {
  const uint32_t* ids = nullptr;   // n * arity slots
  uint32_t n = 0, arity = 0;
  const uint32_t* at(uint32_t k) const { return ids + (size_t)k * arity; }
};

struct GrownTiles {
  std::vector<std::vector<uint32_t>> verts;   // per tile: slots, ascending
};

// Partition the vertices touched by the residual constraints `res` into tiles of at most `cap`
// vertices (aiming at `target`), growing them over the residual's own connectivity so that cuts
// avoid the previous cuts.  tileOfSlot (size V, all NONE on entry) receives the result and is
// reset by the caller afterwards.
void grow_tiles(const CSet& cs, const std::vector<uint32_t>& res, uint32_t cap, uint32_t target,
                std::vector<uint32_t>& compactOf /* size V scratch, all NONE */, GrownTiles& out,
                std::vector<uint32_t>& tileOfSlot) {
  std::vector<uint32_t> slots;
  for (uint32_t k : res)
    for (uint32_t j = 0; j < cs.arity; ++j) {
      const uint32_t s = cs.at(k)[j];
      if (compactOf[s] == NONE) { compactOf[s] = 0; slots.push_back(s); }
    }
  std::sort(slots.begin(), slots.end());
  const uint32_t U = (uint32_t)slots.size();
  for (uint32_t u = 0; u < U; ++u) compactOf[slots[u]] = u;
  std::vector<uint32_t> off(U + 1, 0);
  for (uint32_t k : res)
    for (uint32_t j = 0; j < cs.arity; ++j) off[compactOf[cs.at(k)[j]] + 1]++;
  for (uint32_t u = 0; u < U; ++u) off[u + 1] += off[u];
  std::vector<uint32_t> adj(off[U]), cur(off.begin(), off.end() - 1);
  for (uint32_t k : res)
    for (uint32_t j = 0; j < cs.arity; ++j) adj[cur[compactOf[cs.at(k)[j]]]++] = k;

  auto for_neighbours = [&](uint32_t u, auto&& fn) {
    for (uint32_t a = off[u]; a < off[u + 1]; ++a) {
      const uint32_t* id = cs.at(adj[a]);
      for (uint32_t j = 0; j < cs.arity; ++j) {
        const uint32_t w = compactOf[id[j]];
        if (w != u) fn(w);
      }
    }
  };

  std::vector<uint32_t> tileOf(U, NONE);
  std::vector<uint32_t> tileSize;
  std::vector<uint32_t> comp(U, NONE), dist(U), queue;
  queue.reserve(U);

  std::vector<uint8_t> pending(U, 1);
  uint32_t binTile = NONE;   // tile currently collecting small components
  for (int round = 0; round < 64; ++round) {
    bool any = false;
    std::fill(comp.begin(), comp.end(), NONE);
    for (uint32_t s0 = 0; s0 < U; ++s0) {
      if (!pending[s0] || comp[s0] != NONE) continue;
      any = true;
      queue.clear();
      queue.push_back(s0);
      comp[s0] = s0;
      for (size_t h = 0; h < queue.size(); ++h)
        for_neighbours(queue[h], [&](uint32_t w) {
          if (pending[w] && comp[w] == NONE) { comp[w] = s0; queue.push_back(w); }
        });
      std::vector<uint32_t> members(queue.begin(), queue.end());
      const uint32_t n = (uint32_t)members.size();
      if (n <= cap) {
        // small component: keep it whole; pack several into one tile
        if (binTile == NONE || tileSize[binTile] + n > target) {
          binTile = (uint32_t)tileSize.size();
          tileSize.push_back(0);
        }
        for (uint32_t u : members) { tileOf[u] = binTile; pending[u] = 0; }
        tileSize[binTile] += n;
        continue;
      }
      // large component: seeds by farthest-point sampling over hop distance, then
      // capacity-limited multi-source BFS (graph Voronoi)
      const uint32_t k = (n + target - 1) / target;
      for (uint32_t u : members) dist[u] = NONE;
      std::vector<uint32_t> seeds;
      uint32_t far = members.back();   // last vertex reached by the BFS: a peripheral vertex
      for (uint32_t si = 0; si < k; ++si) {
        seeds.push_back(far);
        queue.clear();
        queue.push_back(far);
        dist[far] = 0;
        for (size_t h = 0; h < queue.size(); ++h) {
          const uint32_t u = queue[h], d = dist[u] + 1;
          for_neighbours(u, [&](uint32_t w) {
            if (comp[w] == s0 && pending[w] && dist[w] > d) { dist[w] = d; queue.push_back(w); }
          });
        }
        uint32_t best = 0;
        far = members[0];
        for (uint32_t u : members)
          if (dist[u] > best) { best = dist[u]; far = u; }
        if (best == 0) break;   // every vertex is a seed already
      }
      const uint32_t base = (uint32_t)tileSize.size();
      tileSize.resize(base + seeds.size(), 0);
      queue.clear();
      for (uint32_t si = 0; si < seeds.size(); ++si) {
        tileOf[seeds[si]] = base + si;
        pending[seeds[si]] = 0;
        tileSize[base + si] = 1;
        queue.push_back(seeds[si]);
      }
      for (size_t h = 0; h < queue.size(); ++h) {
        const uint32_t u = queue[h], t = tileOf[u];
        for_neighbours(u, [&](uint32_t w) {
          if (pending[w] && comp[w] == s0 && tileSize[t] < cap) {
            tileOf[w] = t;
            pending[w] = 0;
            tileSize[t]++;
            queue.push_back(w);
          }
        });
      }
      // vertices a full tile could not take stay pending for the next round
    }
    if (!any) break;
  }
  for (uint32_t u = 0; u < U; ++u)
    if (pending[u]) {
      if (binTile == NONE || tileSize[binTile] + 1 > target) { binTile = (uint32_t)tileSize.size(); tileSize.push_back(0); }
      tileOf[u] = binTile;
      tileSize[binTile]++;
      pending[u] = 0;
    }

  out.verts.assign(tileSize.size(), {});
  for (uint32_t t = 0; t < tileSize.size(); ++t) out.verts[t].reserve(tileSize[t]);
  for (uint32_t u = 0; u < U; ++u) {
    out.verts[tileOf[u]].push_back(slots[u]);   // ascending because slots[] is sorted
    tileOfSlot[slots[u]] = tileOf[u];
  }
  for (uint32_t s : slots) compactOf[s] = NONE;  // leave the scratch clean
}

// ------------------------------------------------------------------ tile construction

struct TypeList {
  std::vector<uint32_t> cons;     // constraint ids, sorted by (colour, id) after colour_list()
  std::vector<uint32_t> colour;   // parallel to cons
  uint32_t nColours = 0;
};

struct TileBuild {
  bool contiguous = false;
  uint32_t rangeBegin = 0, rangeCount = 0;   // contiguous tiles
  std::vector<uint32_t> verts;               // gathered tiles: slots, ascending
  TypeList ty[2];                            // 0 = edges, 1 = tets
  bool mixed = false;                        // both lists share ONE colouring: colour s of either type = step s of the visit
  bool presetVerts = false;                  // verts holds the whole partition cell (tagged hand-over: every phase rewrites every vertex)
  uint32_t nRiders = 0;                      // PBD_ORDER_RIDING: edges that ride on this tile's tets (not in ty[0].cons)
  std::vector<uint8_t> tetPerm;              // parallel to ty[1].cons after place_tile() with relabelling: role permutation codes (empty: none)
  std::vector<uint32_t> localPerm;           // contiguous tiles after place_tile(): slot rangeBegin + i moves to rangeBegin + localPerm[i]
};

// Try to empty the highest colour classes: move each of their constraints to a lower colour that
// is free at all its vertices, or that is blocked by a single constraint which can itself move
// to another free lower colour.  `ids` are tile-local vertex indices, arity per constraint.
uint32_t reduce_colours(const uint32_t* ids, uint32_t n, uint32_t arity, uint32_t nVerts, std::vector<uint32_t>& col,
                        uint32_t nColours) {
  if (nColours <= 1 || nColours > 64) return nColours;   // 64-bit masks are enough for tiles; otherwise keep greedy
  std::vector<uint64_t> used(nVerts, 0);                 // colours present at a vertex
  // owner[v * nColours + c] = constraint of colour c at vertex v
  std::vector<uint32_t> owner((size_t)nVerts * nColours, NONE);
  for (uint32_t k = 0; k < n; ++k)
    for (uint32_t j = 0; j < arity; ++j) {
      const uint32_t v = ids[(size_t)k * arity + j];
      used[v] |= 1ull << col[k];
      owner[(size_t)v * nColours + col[k]] = k;
    }
  auto mask_of = [&](uint32_t k) {
    uint64_t mk = 0;
    for (uint32_t j = 0; j < arity; ++j) mk |= used[ids[(size_t)k * arity + j]];
    return mk;
  };
  auto recolour = [&](uint32_t k, uint32_t c) {
    for (uint32_t j = 0; j < arity; ++j) {
      const uint32_t v = ids[(size_t)k * arity + j];
      // a constraint may list a vertex twice: clear/set is idempotent
      used[v] &= ~(1ull << col[k]);
      owner[(size_t)v * nColours + col[k]] = NONE;
    }
    col[k] = c;
    for (uint32_t j = 0; j < arity; ++j) {
      const uint32_t v = ids[(size_t)k * arity + j];
      used[v] |= 1ull << c;
      owner[(size_t)v * nColours + c] = k;
    }
  };
  std::vector<std::vector<uint32_t>> byColour(nColours);
  for (uint32_t k = 0; k < n; ++k) byColour[col[k]].push_back(k);
  uint32_t top = nColours;
  while (top > 1) {
    const uint32_t hc = top - 1;
    const uint64_t lower = (hc >= 64 ? ~0ull : ((1ull << hc) - 1));
    bool emptied = true;
    std::vector<uint32_t> members;
    for (uint32_t k = 0; k < n; ++k)
      if (col[k] == hc) members.push_back(k);
    std::vector<std::pair<uint32_t, uint32_t>> undo;   // (constraint, previous colour)
    for (uint32_t k : members) {
      const uint64_t freeMask = ~mask_of(k) & lower;
      if (freeMask) {
        undo.push_back({k, col[k]});
        recolour(k, (uint32_t)__builtin_ctzll(freeMask));
        continue;
      }
      // one-step swap: a lower colour c blocked by exactly one constraint b that can move elsewhere
      bool moved = false;
      for (uint32_t c = 0; c < hc && !moved; ++c) {
        uint32_t blocker = NONE;
        bool single = true;
        for (uint32_t j = 0; j < arity && single; ++j) {
          const uint32_t o = owner[(size_t)ids[(size_t)k * arity + j] * nColours + c];
          if (o == NONE || o == k) continue;
          if (blocker == NONE) blocker = o; else if (blocker != o) single = false;
        }
        if (!single || blocker == NONE) continue;
        // colours free for the blocker once k has left hc (k still occupies hc at its vertices: exclude hc anyway)
        const uint64_t fb = ~mask_of(blocker) & lower & ~(1ull << c);
        if (!fb) continue;
        undo.push_back({blocker, col[blocker]});
        recolour(blocker, (uint32_t)__builtin_ctzll(fb));
        if (~mask_of(k) & (1ull << c)) {
          undo.push_back({k, col[k]});
          recolour(k, c);
          moved = true;
        }
      }
      if (!moved) { emptied = false; break; }
    }
    if (!emptied) {
      for (size_t i = undo.size(); i-- > 0;) recolour(undo[i].first, undo[i].second);
      break;
    }
    --top;
  }
  return top;
}

// Tabu search (TabuCol, Hertz & de Werra) on a finished colouring: drop the smallest colour class,
// give its members the least conflicting of the remaining colours, then repair conflicts one
// recolouring at a time (best non-tabu move of a conflicting constraint; a move back is tabu for
// a while).  Repeats while it succeeds and the count is above `floorColours` (the largest vertex
// degree: no colouring can do better).  Greedy + iterated greedy usually end one or two colours
// above that bound; each colour is one block barrier per tile visit.  Deterministic (fixed LCG).
uint32_t tabu_reduce(const uint32_t* ids, uint32_t n, uint32_t arity, uint32_t nLocal, std::vector<uint32_t>& col,
                     uint32_t nc, uint32_t floorColours, uint32_t maxIter, uint32_t seed) {
  if (n < 2 || nc < 2 || nc > 64 || nc <= floorColours) return nc;
  // conflict graph (constraints sharing a vertex), CSR, neighbours unique
  std::vector<uint32_t> vOff((size_t)nLocal + 1, 0), vList;
  auto repeated = [&](uint32_t i, uint32_t j) {
    for (uint32_t q = 0; q < j; ++q)
      if (ids[(size_t)i * arity + q] == ids[(size_t)i * arity + j]) return true;
    return false;
  };
  for (uint32_t i = 0; i < n; ++i)
    for (uint32_t j = 0; j < arity; ++j)
      if (!repeated(i, j)) vOff[ids[(size_t)i * arity + j] + 1]++;
  for (uint32_t v = 0; v < nLocal; ++v) vOff[v + 1] += vOff[v];
  vList.resize(vOff[nLocal]);
  {
    std::vector<uint32_t> cur(vOff.begin(), vOff.end() - 1);
    for (uint32_t i = 0; i < n; ++i)
      for (uint32_t j = 0; j < arity; ++j)
        if (!repeated(i, j)) vList[cur[ids[(size_t)i * arity + j]]++] = i;
  }
  std::vector<uint32_t> aOff((size_t)n + 1, 0), adj, tmp;
  for (int pass = 0; pass < 2; ++pass) {
    for (uint32_t i = 0; i < n; ++i) {
      tmp.clear();
      for (uint32_t j = 0; j < arity; ++j) {
        if (repeated(i, j)) continue;
        const uint32_t v = ids[(size_t)i * arity + j];
        for (uint32_t a = vOff[v]; a < vOff[v + 1]; ++a)
          if (vList[a] != i) tmp.push_back(vList[a]);
      }
      std::sort(tmp.begin(), tmp.end());
      tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
      if (pass == 0) aOff[i + 1] = aOff[i] + (uint32_t)tmp.size();
      else std::copy(tmp.begin(), tmp.end(), adj.begin() + aOff[i]);
    }
    if (pass == 0) adj.resize(aOff[n]);
  }
  uint32_t lcg = seed | 1u;
  auto rnd = [&]() { lcg = lcg * 1664525u + 1013904223u; return lcg >> 8; };
  std::vector<uint32_t> work, conf, confPos(n), gamma, tabuUntil;
  while (nc > floorColours && nc >= 2) {
    const uint32_t k = nc - 1;
    // the smallest class becomes colour k (the one to dissolve)
    std::vector<uint32_t> size(nc, 0);
    for (uint32_t i = 0; i < n; ++i) size[col[i]]++;
    uint32_t drop = 0;
    for (uint32_t c = 1; c < nc; ++c)
      if (size[c] < size[drop]) drop = c;
    work = col;
    for (uint32_t i = 0; i < n; ++i) {
      if (work[i] == drop) work[i] = k;
      else if (work[i] == k) work[i] = drop;
    }
    gamma.assign((size_t)n * k, 0);   // gamma[i*k + c] = neighbours of i coloured c
    for (uint32_t i = 0; i < n; ++i)
      if (work[i] < k)
        for (uint32_t a = aOff[i]; a < aOff[i + 1]; ++a) gamma[(size_t)adj[a] * k + work[i]]++;
    for (uint32_t i = 0; i < n; ++i)
      if (work[i] == k) {
        uint32_t best = 0;
        for (uint32_t c = 1; c < k; ++c)
          if (gamma[(size_t)i * k + c] < gamma[(size_t)i * k + best]) best = c;
        work[i] = best;
        for (uint32_t a = aOff[i]; a < aOff[i + 1]; ++a) gamma[(size_t)adj[a] * k + best]++;
      }
    conf.clear();
    std::fill(confPos.begin(), confPos.end(), NONE);
    uint64_t nConf = 0;
    for (uint32_t i = 0; i < n; ++i)
      if (gamma[(size_t)i * k + work[i]]) { confPos[i] = (uint32_t)conf.size(); conf.push_back(i); nConf += gamma[(size_t)i * k + work[i]]; }
    nConf /= 2;
    uint64_t bestSeen = nConf;
    tabuUntil.assign((size_t)n * k, 0);
    auto set_conf = [&](uint32_t i) {
      const bool c = gamma[(size_t)i * k + work[i]] != 0;
      if (c && confPos[i] == NONE) { confPos[i] = (uint32_t)conf.size(); conf.push_back(i); }
      else if (!c && confPos[i] != NONE) {
        const uint32_t last = conf.back();
        conf[confPos[i]] = last;
        confPos[last] = confPos[i];
        conf.pop_back();
        confPos[i] = NONE;
      }
    };
    uint32_t it = 0, lastGain = 0;
    const uint32_t patience = std::max(400u, maxIter / 4);   // give up on a class that stopped yielding
    for (; it < maxIter && nConf && it - lastGain < patience; ++it) {
      int64_t bestDelta = INT64_MAX;
      uint32_t bi = NONE, bc = 0, ties = 0;
      for (uint32_t i : conf) {
        const uint32_t* gi = &gamma[(size_t)i * k];
        const int64_t cur = gi[work[i]];
        for (uint32_t c = 0; c < k; ++c) {
          if (c == work[i]) continue;
          const int64_t d = (int64_t)gi[c] - cur;
          if (tabuUntil[(size_t)i * k + c] > it && !((int64_t)nConf + d < (int64_t)bestSeen)) continue;
          if (d < bestDelta) { bestDelta = d; bi = i; bc = c; ties = 1; }
          else if (d == bestDelta && (rnd() % ++ties) == 0) { bi = i; bc = c; }
        }
      }
      if (bi == NONE) continue;
      const uint32_t old = work[bi];
      for (uint32_t a = aOff[bi]; a < aOff[bi + 1]; ++a) {
        gamma[(size_t)adj[a] * k + old]--;
        gamma[(size_t)adj[a] * k + bc]++;
      }
      work[bi] = bc;
      nConf = (uint64_t)((int64_t)nConf + bestDelta);
      tabuUntil[(size_t)bi * k + old] = it + (uint32_t)(0.6 * conf.size()) + rnd() % 10u;
      set_conf(bi);
      for (uint32_t a = aOff[bi]; a < aOff[bi + 1]; ++a) set_conf(adj[a]);
      if (nConf < bestSeen) { bestSeen = nConf; lastGain = it; }
    }
    if (nConf) break;   // could not dissolve this class within the budget
    col = work;
    nc = k;
  }
  return nc;
}

// Colour n constraints given by their tile-local vertex indices (`arity` per constraint, caller's
// order; a constraint may list a vertex more than once).  col[i] receives the colour of
// constraint i; returns the number of colours.  Largest-degree-first greedy, the recolouring pass
// above, then iterated greedy.
uint32_t colour_ids(const uint32_t* ids, uint32_t n, uint32_t arity, uint32_t nLocal, std::vector<uint32_t>& col,
                    std::vector<uint32_t>& scratch, int maxIter, uint32_t goal, uint32_t seed, uint32_t tabuIter = 0) {
  col.assign(n, 0);
  if (n == 0) return 0;
  auto repeated = [&](uint32_t i, uint32_t j) {   // vertex j of constraint i already listed at an earlier position
    for (uint32_t q = 0; q < j; ++q)
      if (ids[(size_t)i * arity + q] == ids[(size_t)i * arity + j]) return true;
    return false;
  };
  // visiting order: most constrained first (largest vertex degree inside the tile), then caller's order
  std::vector<uint32_t> deg(nLocal, 0);
  for (uint32_t i = 0; i < n; ++i)
    for (uint32_t j = 0; j < arity; ++j)
      if (!repeated(i, j)) deg[ids[(size_t)i * arity + j]]++;
  std::vector<uint32_t> key(n), visit(n);
  for (uint32_t i = 0; i < n; ++i) {
    uint32_t mx = 0, sum = 0;
    for (uint32_t j = 0; j < arity; ++j) {
      if (repeated(i, j)) continue;
      const uint32_t d = deg[ids[(size_t)i * arity + j]];
      mx = std::max(mx, d);
      sum += d;
    }
    key[i] = (std::min(mx, 0xfffu) << 20) | std::min(sum, 0xfffffu);
  }
  std::iota(visit.begin(), visit.end(), 0u);
  std::stable_sort(visit.begin(), visit.end(), [&](uint32_t a, uint32_t b) { return key[a] > key[b]; });
  scratch.resize((size_t)n * arity);
  for (uint32_t i = 0; i < n; ++i)
    for (uint32_t j = 0; j < arity; ++j) scratch[(size_t)i * arity + j] = ids[(size_t)visit[i] * arity + j];
  std::vector<uint32_t> colV;
  uint32_t nc = greedy_colour(scratch.data(), n, arity, nLocal, colV);
  nc = reduce_colours(scratch.data(), n, arity, nLocal, colV, nc);
  // Iterated greedy (Culberson): re-run first-fit with the constraints grouped by their current
  // colour class and the classes permuted -- never needs more colours than before, often fewer --
  // alternating with the recolouring pass, until the vertex-degree lower bound is met or the
  // colour count stops improving.  Deterministic (fixed LCG for the shuffles).
  uint32_t maxDeg = 0;
  for (uint32_t v = 0; v < nLocal; ++v) maxDeg = std::max(maxDeg, deg[v]);
  {
    std::vector<uint32_t> ids2((size_t)n * arity), perm(n), col2, classOrder, classStart;
    uint32_t lcg = seed, stale = 0;
    const uint32_t stop = std::max(maxDeg, goal);
    uint32_t staleMax = maxIter > 48 ? (uint32_t)maxIter : 12u;
    // when the tabu search below will run (it needs <= 64 colours) a few rounds are enough: it closes
    // the remaining gap to the bound faster than more greedy passes (same step counts, ~40 % less planning time)
    // (tile lists only -- thousands per plan; a whole body, coloured once with a long tabu budget, keeps all rounds)
    if (tabuIter && tabuIter < 10000u && nc <= 64) { maxIter = std::min(maxIter, 4); staleMax = std::min(staleMax, 4u); }
    for (int iter = 0; iter < maxIter && nc > stop && stale < staleMax; ++iter) {
      classOrder.resize(nc);
      std::iota(classOrder.begin(), classOrder.end(), 0u);
      std::vector<uint32_t> size(nc, 0);
      for (uint32_t i = 0; i < n; ++i) size[colV[i]]++;
      switch (iter % 4) {
        case 0: std::reverse(classOrder.begin(), classOrder.end()); break;
        case 1: std::stable_sort(classOrder.begin(), classOrder.end(), [&](uint32_t a, uint32_t b) { return size[a] > size[b]; }); break;
        case 2: std::stable_sort(classOrder.begin(), classOrder.end(), [&](uint32_t a, uint32_t b) { return size[a] < size[b]; }); break;
        default:
          for (uint32_t i = nc; i > 1; --i) {
            lcg = lcg * 1664525u + 1013904223u;
            std::swap(classOrder[i - 1], classOrder[(lcg >> 8) % i]);
          }
      }
      std::vector<uint32_t> rank(nc);
      for (uint32_t c = 0; c < nc; ++c) rank[classOrder[c]] = c;
      std::iota(perm.begin(), perm.end(), 0u);
      std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return rank[colV[a]] < rank[colV[b]]; });
      for (uint32_t i = 0; i < n; ++i)
        for (uint32_t j = 0; j < arity; ++j) ids2[(size_t)i * arity + j] = scratch[(size_t)perm[i] * arity + j];
      uint32_t nc2 = greedy_colour(ids2.data(), n, arity, nLocal, col2);
      nc2 = reduce_colours(ids2.data(), n, arity, nLocal, col2, nc2);
      if (nc2 <= nc) {
        stale = nc2 < nc ? 0 : stale + 1;
        nc = nc2;
        for (uint32_t i = 0; i < n; ++i) colV[perm[i]] = col2[i];
      } else {
        ++stale;
      }
    }
  }
  if (tabuIter && nc > std::max(maxDeg, goal)) nc = tabu_reduce(scratch.data(), n, arity, nLocal, colV, nc, std::max(maxDeg, goal), tabuIter, seed);
  for (uint32_t i = 0; i < n; ++i) col[visit[i]] = colV[i];
  return nc;
}

// Even out colour classes so that none holds more than `cap` constraints (what one block pass of
// the sweep takes; a larger class costs a second colour step): constraints of oversized classes
// move to classes with room where all their vertices are free; when the classes cannot hold
// everything (nc * cap < n) new classes are opened.  Colourings with more than 64 classes are left
// alone (the caller splits oversized groups anyway).  Returns the new number of classes.
uint32_t cap_classes(const uint32_t* ids, uint32_t n, uint32_t arity, uint32_t nLocal, std::vector<uint32_t>& col,
                     uint32_t nc, uint32_t cap) {
  if (cap == 0 || nc == 0 || nc > 64) return nc;
  std::vector<uint32_t> size(64, 0);
  for (uint32_t k = 0; k < n; ++k) size[col[k]]++;
  bool over = false;
  for (uint32_t c = 0; c < nc; ++c) over |= size[c] > cap;
  if (!over) return nc;
  uint32_t nc2 = std::max(nc, (n + cap - 1) / cap);
  if (nc2 > 64) return nc;
  std::vector<uint64_t> used(nLocal, 0);
  for (uint32_t k = 0; k < n; ++k)
    for (uint32_t j = 0; j < arity; ++j) used[ids[(size_t)k * arity + j]] |= 1ull << col[k];
  auto mask_of = [&](uint32_t k) {
    uint64_t mk = 0;
    for (uint32_t j = 0; j < arity; ++j) mk |= used[ids[(size_t)k * arity + j]];
    return mk;
  };
  for (int round = 0; round < 4; ++round) {
    for (uint32_t k = 0; k < n; ++k) {
      const uint32_t c = col[k];
      if (size[c] <= cap) continue;
      uint64_t freeMask = ~mask_of(k) & (nc2 >= 64 ? ~0ull : ((1ull << nc2) - 1));
      uint32_t best = NONE;
      while (freeMask) {
        const uint32_t d = (uint32_t)__builtin_ctzll(freeMask);
        freeMask &= freeMask - 1;
        if (size[d] < cap && (best == NONE || size[d] < size[best])) best = d;
      }
      if (best == NONE) continue;
      for (uint32_t j = 0; j < arity; ++j) used[ids[(size_t)k * arity + j]] &= ~(1ull << c);
      for (uint32_t j = 0; j < arity; ++j) used[ids[(size_t)k * arity + j]] |= 1ull << best;
      col[k] = best;
      size[c]--;
      size[best]++;
    }
    over = false;
    for (uint32_t c = 0; c < nc2; ++c) over |= size[c] > cap;
    if (!over || nc2 >= 64) break;
    ++nc2;   // one more class and another round
  }
  // drop classes that stayed empty
  std::vector<uint32_t> remap(64, NONE);
  uint32_t m2 = 0;
  for (uint32_t c = 0; c < nc2; ++c)
    if (size[c]) remap[c] = m2++;
  for (uint32_t k = 0; k < n; ++k) col[k] = remap[col[k]];
  return m2;
}

// sort a tile's constraint list by (colour, id) given the colours of its current order
void sort_by_colour(TypeList& tl, const std::vector<uint32_t>& col, uint32_t nColours) {
  const uint32_t n = (uint32_t)tl.cons.size();
  tl.nColours = nColours;
  std::vector<uint32_t> order(n);
  std::iota(order.begin(), order.end(), 0u);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return col[a] < col[b]; });
  std::vector<uint32_t> c2(n), k2(n);
  for (uint32_t i = 0; i < n; ++i) { c2[i] = tl.cons[order[i]]; k2[i] = col[order[i]]; }
  tl.cons.swap(c2);
  tl.colour.swap(k2);
}

// colour the constraints of one tile locally and sort them by (colour, id)
// `cap`: most constraints one colour step can take (0 = no limit)
void colour_list(const CSet& cs, TypeList& tl, uint32_t nLocal, const std::vector<uint32_t>& localOf,
                 std::vector<uint32_t>& scratch, uint32_t cap, int maxIter = 48, uint32_t goal = 0,
                 uint32_t seed = 0x9e3779b9u, uint32_t tabuBudget = 2000u) {
  const uint32_t tabuIter = knobs().tabu >= 0 ? (uint32_t)knobs().tabu : tabuBudget;
  const uint32_t n = (uint32_t)tl.cons.size();
  std::sort(tl.cons.begin(), tl.cons.end());
  std::vector<uint32_t> ids((size_t)n * cs.arity), col;
  for (uint32_t i = 0; i < n; ++i)
    for (uint32_t j = 0; j < cs.arity; ++j) ids[(size_t)i * cs.arity + j] = localOf[cs.at(tl.cons[i])[j]];
  uint32_t nc = colour_ids(ids.data(), n, cs.arity, nLocal, col, scratch, maxIter, goal, seed, tabuIter);
  nc = cap_classes(ids.data(), n, cs.arity, nLocal, col, nc, cap);
  sort_by_colour(tl, col, nc);
}

// Mixed steps (interleaved order, one thread per constraint): colour a tile's edges AND tets
// together, so that one colour step of the visit projects a vertex-disjoint set of edges (the
// block's first warps) and tets (its last warps) at the same time.  A visit then takes about
// max-joint-vertex-load steps instead of (edge colours + tet colours), and the warps a tet-only
// step leaves idle do edge work.  `threads` = block size: a step holds at most that many threads,
// edges and tets in separate warps.  Afterwards both lists are sorted by (step, id) and
// ty[0].nColours == ty[1].nColours == number of steps.
void colour_joint(const CSet sets[2], TileBuild& tb, uint32_t nLocal, const std::vector<uint32_t>& localOf,
                  std::vector<uint32_t>& scratch, uint32_t threads, int maxIter = 48, uint32_t goal = 0,
                  uint32_t seed = 0x9e3779b9u) {
  TypeList& LE = tb.ty[0];
  TypeList& LT = tb.ty[1];
  std::sort(LE.cons.begin(), LE.cons.end());
  std::sort(LT.cons.begin(), LT.cons.end());
  const uint32_t nT = (uint32_t)LT.cons.size(), nE = (uint32_t)LE.cons.size(), n = nT + nE;
  // joint list: tets first, then edges as (a, b, a, b)
  std::vector<uint32_t> ids((size_t)n * 4), col;
  for (uint32_t i = 0; i < nT; ++i)
    for (uint32_t j = 0; j < 4; ++j) ids[(size_t)i * 4 + j] = localOf[sets[1].at(LT.cons[i])[j]];
  for (uint32_t i = 0; i < nE; ++i)
    for (uint32_t j = 0; j < 4; ++j) ids[(size_t)(nT + i) * 4 + j] = localOf[sets[0].at(LE.cons[i])[j & 1u]];
  const uint32_t tabuIter = knobs().tabu >= 0 ? (uint32_t)knobs().tabu : 4000u;
  uint32_t nc = colour_ids(ids.data(), n, 4, nLocal, col, scratch, maxIter, goal, seed, tabuIter);

  auto pad32 = [](uint32_t x) { return (x + 31u) & ~31u; };
  if (nc <= 64) {
    // ---- even out the steps.  A step costs a fixed latency plus issue time that grows with its
    // warps (a tet ~ twice an edge), and it must fit the block.
    std::vector<uint64_t> used(nLocal, 0);
    for (uint32_t k = 0; k < n; ++k)
      for (uint32_t j = 0; j < 4; ++j) used[ids[(size_t)k * 4 + j]] |= 1ull << col[k];
    std::vector<uint32_t> cntE(64, 0), cntT(64, 0);
    for (uint32_t k = 0; k < n; ++k) (k < nT ? cntT : cntE)[col[k]]++;
    auto cost = [&](uint32_t c) { return cntE[c] + 2u * cntT[c]; };
    auto fits = [&](uint32_t c, bool tet) { return pad32(cntE[c] + (tet ? 0u : 1u)) + pad32(cntT[c] + (tet ? 1u : 0u)) <= threads; };
    auto mask_of = [&](uint32_t k) {
      uint64_t mk = 0;
      for (uint32_t j = 0; j < 4; ++j) mk |= used[ids[(size_t)k * 4 + j]];
      return mk;
    };
    auto move = [&](uint32_t k, uint32_t d) {
      const bool tet = k < nT;
      for (uint32_t j = 0; j < 4; ++j) used[ids[(size_t)k * 4 + j]] &= ~(1ull << col[k]);
      (tet ? cntT : cntE)[col[k]]--;
      col[k] = d;
      for (uint32_t j = 0; j < 4; ++j) used[ids[(size_t)k * 4 + j]] |= 1ull << d;
      (tet ? cntT : cntE)[d]++;
    };
    for (int pass = 0; pass < 6; ++pass) {
      uint32_t moves = 0;
      for (uint32_t k = 0; k < n; ++k) {
        const uint32_t c = col[k], w = k < nT ? 2u : 1u;
        uint64_t freeMask = ~mask_of(k) & (nc >= 64 ? ~0ull : ((1ull << nc) - 1));
        uint32_t best = NONE, bestCost = cost(c) - w;   // the target must end cheaper than the source is after the move
        while (freeMask) {
          const uint32_t d = (uint32_t)__builtin_ctzll(freeMask);
          freeMask &= freeMask - 1;
          if (cost(d) + w < bestCost + 0u && fits(d, k < nT)) { best = d; bestCost = cost(d) + w; }
        }
        if (best != NONE) { move(k, best); ++moves; }
      }
      if (!moves) break;
    }
    // ---- steps that still exceed the block: spill into other free steps, else into a new one
    for (uint32_t c = 0; c < nc && nc < 64; ++c) {
      if (pad32(cntE[c]) + pad32(cntT[c]) <= threads) continue;
      for (uint32_t k = 0; k < n && pad32(cntE[c]) + pad32(cntT[c]) > threads; ++k) {
        if (col[k] != c) continue;
        uint64_t freeMask = ~mask_of(k) & ((1ull << nc) - 1);
        uint32_t best = NONE;
        while (freeMask) {
          const uint32_t d = (uint32_t)__builtin_ctzll(freeMask);
          freeMask &= freeMask - 1;
          if (fits(d, k < nT) && (best == NONE || cost(d) < cost(best))) best = d;
        }
        if (best != NONE) move(k, best);
      }
      if (pad32(cntE[c]) + pad32(cntT[c]) > threads && nc < 64) {
        const uint32_t d = nc++;   // members of one step share no vertex: any subset forms a valid new step
        for (uint32_t k = 0; k < n && pad32(cntE[c]) + pad32(cntT[c]) > threads; ++k)
          if (col[k] == c && fits(d, k < nT)) move(k, d);
      }
    }
  }
  std::vector<uint32_t> colT(col.begin(), col.begin() + nT), colE(col.begin() + nT, col.end());
  sort_by_colour(LT, colT, nc);
  sort_by_colour(LE, colE, nc);
  tb.mixed = true;
}

// Order the constraints of one colour group so that the shared-memory gathers of a warp do not
// collide: a 16-byte vertex load is served per quarter-warp (8 lanes), conflict-free when the 8
// tile-local vertex indices of each role differ modulo 8.  Greedy: fill one quarter-warp at a
// time with the first remaining constraints whose indices are still free in every role.  The
// order inside a colour group never changes the result (its constraints share no vertex).
void bank_order(const CSet& cs, const std::vector<uint32_t>& localOf, uint32_t* cons, uint32_t n) {
  if (n <= 1) return;
  const uint32_t ar = cs.arity;
  // residues of every constraint, one byte per role
  std::vector<uint8_t> res((size_t)n * 4, 0);
  for (uint32_t i = 0; i < n; ++i)
    for (uint32_t r = 0; r < ar; ++r) res[(size_t)i * 4 + r] = (uint8_t)(localOf[cs.at(cons[i])[r]] & 7u);
  std::vector<uint32_t> pick, out(n);
  pack_rows_greedy(res.data(), ar, n, 0x2545f491u, 24, pick);
  for (uint32_t i = 0; i < n; ++i) out[i] = cons[pick[i]];
  std::copy(out.begin(), out.end(), cons);
}

// Shared-memory placement of one finished (coloured) tile: pbd_placement.cpp renumbers the tile's vertices and
// reorders every colour group (same groups, same colours: no result changes).  Gathered tiles take the new
// numbering by reordering their vertex list; a contiguous tile records it in localPerm and the planner
// renumbers the slots of its range afterwards.  effort 0: statistics only.
// relabel: the tets' vertices may change roles (fast arithmetic; the permutation codes go to tb.tetPerm).
void place_tile(const CSet sets[2], TileBuild& tb, std::vector<uint32_t>& localOf, int effort, uint32_t block, bool relabel, PlaceStats& stats) {
  const uint32_t nLocal = tb.contiguous ? tb.rangeCount : (uint32_t)tb.verts.size();
  if (tb.contiguous) for (uint32_t i = 0; i < nLocal; ++i) localOf[tb.rangeBegin + i] = i;
  else for (uint32_t i = 0; i < nLocal; ++i) localOf[tb.verts[i]] = i;
  std::vector<PlaceGroup> groups;
  std::vector<uint32_t> loc, payload;
  size_t typeBegin[3] = {0, 0, 0};
  for (int ty = 0; ty < 2; ++ty) {
    const TypeList& L = tb.ty[ty];
    const CSet& cs = sets[ty];
    for (size_t i = 0; i < L.cons.size();) {
      size_t j = i;
      while (j < L.cons.size() && L.colour[j] == L.colour[i]) ++j;
      groups.push_back({(uint32_t)payload.size(), (uint32_t)(j - i), cs.arity});
      for (size_t k = i; k < j; ++k) {
        payload.push_back(L.cons[k]);
        for (uint32_t r = 0; r < 4; ++r) loc.push_back(r < cs.arity ? localOf[cs.at(L.cons[k])[r]] : NONE);
      }
      i = j;
    }
    typeBegin[ty + 1] = payload.size();
  }
  std::vector<uint32_t> newLocal;
  std::vector<uint8_t> perm(relabel && effort > 0 ? payload.size() : 0);
  optimise_placement(nLocal, groups.data(), (uint32_t)groups.size(), loc.data(), payload.data(), (uint32_t)payload.size(), effort,
                     block, newLocal, &stats, perm.empty() ? nullptr : perm.data());
  if (effort <= 0) return;
  if (!perm.empty()) tb.tetPerm.assign(perm.begin() + typeBegin[1], perm.begin() + typeBegin[2]);
  for (int ty = 0; ty < 2; ++ty)   // (colours are unchanged: the groups keep their places)
    std::copy(payload.begin() + typeBegin[ty], payload.begin() + typeBegin[ty + 1], tb.ty[ty].cons.begin());
  if (tb.contiguous) {
    tb.localPerm = newLocal;
  } else {
    std::vector<uint32_t> moved(nLocal);
    for (uint32_t i = 0; i < nLocal; ++i) moved[newLocal[i]] = tb.verts[i];
    tb.verts.swap(moved);
  }
}

// Finish a tile: vertex list (gathered tiles: the slots its constraints touch), local numbering,
// colouring of both lists.  localOf is scratch of size V.
// mixedThreads != 0: tiles that carry both types get ONE joint colouring (steps of at most that
// many threads, see colour_joint).
bool steps_fit(const TileBuild& tb, uint32_t threads) {
  std::vector<uint32_t> cnt[2];
  for (int ty = 0; ty < 2; ++ty) {
    cnt[ty].assign(tb.ty[ty].nColours, 0);
    for (uint32_t c : tb.ty[ty].colour) cnt[ty][c]++;
  }
  if (cnt[0].size() != cnt[1].size()) return false;
  for (size_t s = 0; s < cnt[0].size(); ++s)
    if (((cnt[0][s] + 31u) & ~31u) + ((cnt[1][s] + 31u) & ~31u) > threads) return false;
  return true;
}

void order_groups(const CSet sets[2], TileBuild& tb, const std::vector<uint32_t>& localOf, int ty) {
  TypeList& L = tb.ty[ty];
  for (size_t i = 0; i < L.cons.size();) {
    size_t j = i;
    while (j < L.cons.size() && L.colour[j] == L.colour[i]) ++j;
    bank_order(sets[ty], localOf, &L.cons[i], (uint32_t)(j - i));
    i = j;
  }
}

void finish_tile(const CSet sets[2], TileBuild& tb, std::vector<uint32_t>& localOf, std::vector<uint32_t>& scratch,
                 const uint32_t caps[2], uint32_t mixedThreads = 0, bool orderGroups = true) {
  uint32_t nLocal;
  if (tb.contiguous) {
    nLocal = tb.rangeCount;
    for (uint32_t i = 0; i < tb.rangeCount; ++i) localOf[tb.rangeBegin + i] = i;
  } else {
    if (!tb.presetVerts) {
      std::vector<uint32_t> used;
      for (int ty = 0; ty < 2; ++ty)
        for (uint32_t k : tb.ty[ty].cons)
          for (uint32_t j = 0; j < sets[ty].arity; ++j) used.push_back(sets[ty].at(k)[j]);
      std::sort(used.begin(), used.end());
      used.erase(std::unique(used.begin(), used.end()), used.end());
      tb.verts.swap(used);
    }
    nLocal = (uint32_t)tb.verts.size();
    for (uint32_t i = 0; i < nLocal; ++i) localOf[tb.verts[i]] = i;
  }
  tb.mixed = false;
  if (mixedThreads && !tb.ty[0].cons.empty() && !tb.ty[1].cons.empty()) {
    colour_joint(sets, tb, nLocal, localOf, scratch, mixedThreads);
    if (!steps_fit(tb, mixedThreads)) tb.mixed = false;   // (more than 64 joint colours: keep the two sweeps apart)
  }
  for (int ty = 0; ty < 2; ++ty) {
    TypeList& L = tb.ty[ty];
    if (L.cons.empty()) continue;
    if (!tb.mixed) colour_list(sets[ty], L, nLocal, localOf, scratch, caps[ty]);
    if (orderGroups) order_groups(sets, tb, localOf, ty);   // (main tiles the placement search will reorder anyway skip this)
  }
}

uint32_t tile_bytes(const TileBuild& tb, bool ride = false) {
  const uint32_t nv = tb.contiguous ? tb.rangeCount : (uint32_t)tb.verts.size();
  const uint32_t rec = tile_record_bytes(tb.contiguous ? 0u : nv, tb.ty[0].nColours, tb.ty[1].nColours,
                                         (uint32_t)tb.ty[0].cons.size() + tb.nRiders, (uint32_t)tb.ty[1].cons.size(), ride);
  return 16u * nv + 2u * rec;   // vertices + double-buffered record block
}

// Residual phases of one type: re-partition until nothing is left.
void build_residual_phases(const CSet sets[2], int ty, std::vector<uint32_t> res, uint32_t V, uint32_t cap,
                           std::vector<std::vector<TileBuild>>& phases, std::vector<uint32_t>& localOf,
                           std::vector<uint32_t>& scratch, const uint32_t caps[2]) {
  const CSet& cs = sets[ty];
  std::vector<uint32_t> compactOf(V, NONE), tileOfSlot(V, NONE);
  const uint32_t target = std::max(16u, (uint32_t)(cap * 0.75));
  while (!res.empty()) {
    GrownTiles g;
    grow_tiles(cs, res, cap, target, compactOf, g, tileOfSlot);
    std::vector<TileBuild> ph(g.verts.size());
    std::vector<uint32_t> next;
    for (uint32_t k : res) {
      const uint32_t* id = cs.at(k);
      const uint32_t t = tileOfSlot[id[0]];
      bool same = t != NONE;
      for (uint32_t j = 1; j < cs.arity && same; ++j) same = tileOfSlot[id[j]] == t;
      if (same) ph[t].ty[ty].cons.push_back(k); else next.push_back(k);
    }
    if (next.size() == res.size()) {
      // no progress (cannot happen with hop-distance Voronoi on a connected residual, but stay
      // safe): peel off a vertex-disjoint set of single-constraint tiles
      for (auto& vs : g.verts)
        for (uint32_t s : vs) tileOfSlot[s] = NONE;
      ph.clear();
      g.verts.clear();
      next.clear();
      for (uint32_t k : res) {
        const uint32_t* id = cs.at(k);
        bool free = true;
        for (uint32_t j = 0; j < cs.arity; ++j) free &= tileOfSlot[id[j]] == NONE;
        if (!free) { next.push_back(k); continue; }
        std::vector<uint32_t> vs(id, id + cs.arity);
        for (uint32_t s : vs) tileOfSlot[s] = (uint32_t)g.verts.size();
        g.verts.push_back(vs);
        ph.emplace_back();
        ph.back().ty[ty].cons.push_back(k);
      }
    }
    std::vector<TileBuild> kept;
    for (auto& tb : ph) {
      if (tb.ty[ty].cons.empty()) continue;
      finish_tile(sets, tb, localOf, scratch, caps);
      kept.push_back(std::move(tb));
    }
    for (auto& vs : g.verts)
      for (uint32_t s : vs) tileOfSlot[s] = NONE;
    phases.push_back(std::move(kept));
    res.swap(next);
  }
}

}  // namespace

void colour_and_order(const uint32_t* ids, uint32_t n, uint32_t arity, uint32_t nVerts, uint32_t maxGroup,
                      std::vector<uint32_t>& order, std::vector<uint32_t>& counts) {
  order.clear();
  counts.clear();
  if (n == 0) return;
  const CSet cs{ids, n, arity};
  std::vector<uint32_t> localOf(nVerts), scratch;
  std::iota(localOf.begin(), localOf.end(), 0u);
  TypeList L;
  L.cons.resize(n);
  std::iota(L.cons.begin(), L.cons.end(), 0u);
  // a whole body is coloured once per topology (the batch backend caches it) and every colour is a
  // block barrier per iteration for every body of that topology: a long tabu search pays
  colour_list(cs, L, nVerts, localOf, scratch, maxGroup, 48, 0, 0x9e3779b9u, 30000u);
  for (size_t i = 0; i < L.cons.size();) {
    size_t j = i;
    while (j < L.cons.size() && L.colour[j] == L.colour[i]) ++j;
    bank_order(cs, localOf, &L.cons[i], (uint32_t)(j - i));
    const uint32_t nGrp = (uint32_t)(j - i), lim = std::max(1u, maxGroup), parts = (nGrp + lim - 1) / lim;
    for (uint32_t q = 0; q < parts; ++q)   // a colour larger than one block pass becomes several groups
      counts.push_back((uint32_t)(((uint64_t)nGrp * (q + 1)) / parts - ((uint64_t)nGrp * q) / parts));
    i = j;
  }
  order = L.cons;
}

bool build_tile_plan(const MeshView& m, const pbd_options& opts, uint32_t nSMs, uint32_t smemBytes, Plan& plan,
                     std::string& err) {
  const double t0 = now_ms();
  // PBD_PLAN_DEBUG: wall time of every planning stage
  double tStage = t0;
  const bool stageLog = knobs().debug;
#define PBD_PLAN_STAGE(name) do { if (stageLog) { const double t1_ = now_ms(); fprintf(stderr, "[plan] %-16s %8.0f ms\n", name, t1_ - tStage); tStage = t1_; } } while (0)
  if (nSMs == 0) nSMs = 148;
  if (smemBytes < 16384) { err = "shared memory too small for a vertex tile"; return false; }
  plan = Plan();
  plan.V = m.V; plan.E = m.E; plan.T = m.T;
  plan.backend = PBD_BACKEND_TILE;
  plan.orderMode = opts.order_mode;
  if (opts.order_mode != PBD_ORDER_STRICT && opts.order_mode != PBD_ORDER_INTERLEAVED && opts.order_mode != PBD_ORDER_RIDING) { err = "unknown order_mode"; return false; }
  const bool riding = opts.order_mode == PBD_ORDER_RIDING;
  if (riding && opts.lanes_per_tet > 1) { err = "PBD_ORDER_RIDING needs one thread per tet (lanes_per_tet <= 1)"; return false; }
  const bool fused = opts.order_mode == PBD_ORDER_INTERLEAVED || riding;
  // tiles (CTAs) per SM.  auto: two half-size tiles of 256 threads with the tagged hand-over on a body that
  // fills every SM (measured +4.6 % fast / +2.2 % exact on the 1M-tet body: one CTA's hand-over latencies
  // overlap the other's sweep); one 512-thread tile otherwise (with done counters two CTAs measured -3 %:
  // MEMBAR.GPU stalls the whole SM's memory pipeline)
  if (nSMs == 0) nSMs = 148;
  const uint32_t perSmPlan = opts.tiles_per_sm ? opts.tiles_per_sm
                             : (((opts.flags & PBD_FLAG_TAGGED_HANDOVER) && !opts.tile_vertices &&
                                 (uint64_t)m.V >= (uint64_t)nSMs * 1024u) ? 2u : 1u);
  // smallest tile before fewer SMs are used instead.  With done counters a tile below ~1024 vertices is all
  // hand-over; the tagged hand-over is cheap enough that 256-vertex tiles on more SMs win (measured on the
  // 100k-tet body: 6,280 -> 8,430 substeps/s, 19 -> 75 tiles per partition)
  const uint32_t minTileAuto = getenv("PBD_PLAN_MINTILE") ? (uint32_t)knobs().minTile
                               : ((opts.flags & PBD_FLAG_TAGGED_HANDOVER) ? 256u : 1024u);   // (the same for every rank of a sharded body and for its single-GPU twin: plans must be identical)
  const bool smallTiles = !opts.tile_vertices && perSmPlan == 1 &&
                          (uint64_t)m.V / std::max<uint64_t>(1, std::min<uint64_t>(nSMs, m.V / std::max(1u, minTileAuto))) < 640u;
  const uint32_t blockThreads = opts.block_threads ? opts.block_threads : ((perSmPlan >= 2 || smallTiles) ? 256u : 512u);
  if (blockThreads % 32 || blockThreads > 512) { err = "block_threads must be a multiple of 32, <= 512"; return false; }
  plan.blockThreads = blockThreads;
  plan.tilesPerSm = perSmPlan;
  // interleaved order, one thread per tet: edges and tets of a tile visit share the colour steps
  const bool noMixed = !knobs().mixed;   // debug: A/B against separate sweeps
  const uint32_t mixedThreads = (fused && opts.lanes_per_tet <= 1 && (!noMixed || riding)) ? blockThreads : 0u;
  // most constraints of one type a colour step can take: one block pass
  const uint32_t caps[2] = {blockThreads, std::max(1u, blockThreads / std::max(1u, opts.lanes_per_tet))};
  if (opts.partitions > kMaxPartitions) { err = "partitions must be <= 8"; return false; }

  // body frame, extents -> axis order for the k-d levels
  Partitioner base;
  std::vector<float> xf((size_t)m.V * 3);
  {
    double R[9];
    body_frame(m.x0, m.V, R);
    for (uint32_t v = 0; v < m.V; ++v)
      for (int i = 0; i < 3; ++i)
        xf[3 * (size_t)v + i] = (float)(R[3 * i] * m.x0[3 * (size_t)v] + R[3 * i + 1] * m.x0[3 * (size_t)v + 1] +
                                        R[3 * i + 2] * m.x0[3 * (size_t)v + 2]);
  }
  base.x = xf.data();
  {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t v = 0; v < m.V; ++v)
      for (int a = 0; a < 3; ++a) {
        const float c = xf[3 * (size_t)v + a];
        mn[a] = std::min(mn[a], c);
        mx[a] = std::max(mx[a], c);
      }
    double e[3];
    for (int a = 0; a < 3; ++a) e[a] = m.V ? std::max(0.0, (double)mx[a] - (double)mn[a]) : 1.0;
    const double big = std::max({e[0], e[1], e[2], 1e-30});
    for (int a = 0; a < 3; ++a) e[a] = std::max(e[a], big * 1e-3);   // flat bodies: keep the ratios finite
    std::sort(base.axis, base.axis + 3, [&](int a, int b) { return e[a] > e[b] || (e[a] == e[b] && a < b); });
    for (int l = 0; l < 3; ++l) base.ext[l] = e[base.axis[l]];
  }

  // ---- tiles per partition: one tile per SM unless the body is small (fewer, fuller tiles) or
  // a tile would not fit in shared memory (more tiles, whole waves)
  uint32_t K1;
  if (opts.tile_vertices) {
    K1 = std::max(1u, (m.V + opts.tile_vertices - 1) / opts.tile_vertices);
  } else {
    const uint32_t perSm = perSmPlan;
    const uint32_t minTile = std::max(32u, minTileAuto / perSm);   // below this a tile is all interface: use fewer SMs instead
    K1 = std::max(1u, std::min(nSMs * perSm, m.V / std::max(1u, minTile)));
    // large bodies: enough tiles (whole waves) that a tile fits in shared memory, estimated from the
    // bytes a tile visit needs (16 B/vertex + double-buffered records, 12 B/edge and 16 B/tet, of
    // ~1/4 of its constraints, with 25 % headroom for uneven tiles) instead of finding out by
    // repeated planning
    const double need = 16.0 * m.V + 2.0 * (12.0 * m.E + 16.0 * m.T) / 4.0;
    const uint32_t byBytes = (uint32_t)std::ceil(1.25 * need / (double)smemBytes);
    if (byBytes > K1) K1 = ((byBytes + nSMs - 1) / nSMs) * nSMs;
    // Two tiles per SM only pay while two record-block pairs fit an SM: bodies larger than one wave keep
    // the tile size of the one-wave case (~640 vertices) and take more waves instead of growing the
    // tiles until one barely fits (measured on the 8.4M-tet body, one GPU: 729 -> 775 substeps/s, and the
    // plan for 2 x 148 SMs 572 -> what two GPUs need to scale)
    if (perSm >= 2) {
      const uint32_t wave = nSMs * perSm, bySize = (m.V + 639u) / 640u;
      if (bySize > K1) K1 = ((bySize + wave - 1) / wave) * wave;
    }
  }

  std::vector<uint32_t> localOf(m.V, NONE), scratch;
  for (int attempt = 0;; ++attempt) {
    if (attempt > 24) { err = "could not fit the tiles into shared memory"; return false; }
    const uint32_t K = K1 <= 1 ? 1u : (opts.partitions ? opts.partitions : 4u);

    // A regular m0 x m1 x m2 arrangement keeps the cuts of all slabs / columns aligned, which is
    // what makes the shifted partitions cover every constraint of a box-like body.  Take the
    // largest product <= K1 whose tiles are not too elongated (when the caller fixed the tile
    // size, or the body is tiny, keep K1 and split unevenly instead).
    base.grid[0] = base.grid[1] = base.grid[2] = 0;
    uint32_t tilesWanted = K1;
    if (!opts.tile_vertices && K1 >= 8) {
      uint32_t bestProd = 0;
      double bestAspect = 0.0;
      for (uint32_t a = 1; a <= K1; ++a)
        for (uint32_t b = 1; a * b <= K1; ++b) {
          const uint32_t c = K1 / (a * b);
          if (c == 0) continue;
          const double d0 = base.ext[0] / a, d1 = base.ext[1] / b, d2 = base.ext[2] / c;
          const double aspect = std::max({d0, d1, d2}) / std::min({d0, d1, d2});
          if (aspect > 1.75) continue;
          const uint32_t prod = a * b * c;
          if (prod > bestProd || (prod == bestProd && aspect < bestAspect)) {
            bestProd = prod; bestAspect = aspect;
            base.grid[0] = a; base.grid[1] = b; base.grid[2] = c;
          }
        }
      if (bestProd * 10 >= K1 * 8) tilesWanted = bestProd;   // give up at most 20 % of the SMs for regularity
      else base.grid[0] = base.grid[1] = base.grid[2] = 0;
    }

    PBD_PLAN_STAGE("frame+sizing");
    // ---- partitions (per caller vertex), P_0 also defines the slot numbering
    std::vector<std::vector<uint32_t>> tileOfV(K, std::vector<uint32_t>(m.V, 0));
    std::vector<uint32_t> tile0Begin, slotToVertex;
    uint32_t nTilesMax = 1;
    {
      // the K partitions are independent of each other (own permutation, own tileOf array): one host thread each
      std::vector<uint32_t> nTilesOf(K, 1u);
      auto split_partition = [&](uint32_t p) {
        Partitioner pt = base;
        pt.offset = (double)p / K;
        pt.idx.resize(m.V);
        std::iota(pt.idx.begin(), pt.idx.end(), 0u);
        pt.tileOf = tileOfV[p].data();
        if (m.V) pt.split(0, m.V, tilesWanted, 0); else pt.tileBegin.push_back(0);
        nTilesOf[p] = (uint32_t)pt.tileBegin.size();
        if (p == 0) {
          tile0Begin = pt.tileBegin;
          tile0Begin.push_back(m.V);
          slotToVertex = std::move(pt.idx);
        }
      };
      std::vector<std::thread> pool;
      if (m.V >= 50000u && std::thread::hardware_concurrency() > 1)
        for (uint32_t p = 1; p < K; ++p) pool.emplace_back(split_partition, p);
      else
        for (uint32_t p = 1; p < K; ++p) split_partition(p);
      split_partition(0);
      for (auto& th : pool) th.join();
      for (uint32_t p = 0; p < K; ++p) nTilesMax = std::max(nTilesMax, nTilesOf[p]);
    }
    const uint32_t nTile0 = (uint32_t)tile0Begin.size() - 1;
    // Slot order inside a home tile: vertices that share their tile in EVERY shifted partition sit next to each
    // other (sorted by that signature, caller order inside a signature).  A shifted tile then takes runs of
    // consecutive slots from each home tile instead of a scattered subset: fewer half-used 32-byte sectors in the
    // 16-byte vertex loads and stores of a visit, whose L2 traffic the frame time follows.
    if (K > 1 && knobs().sigSort)
      for (uint32_t t = 0; t < nTile0; ++t)
        std::sort(slotToVertex.begin() + tile0Begin[t], slotToVertex.begin() + tile0Begin[t + 1], [&](uint32_t a, uint32_t b) {
          for (uint32_t p = 1; p < K; ++p)
            if (tileOfV[p][a] != tileOfV[p][b]) return tileOfV[p][a] < tileOfV[p][b];
          return a < b;
        });
    std::vector<uint32_t> vertexToSlot(m.V);
    for (uint32_t s = 0; s < m.V; ++s) vertexToSlot[slotToVertex[s]] = s;

    // constraints in slot numbering; per-partition tile of a slot
    std::vector<uint32_t> eSlots((size_t)m.E * 2), tSlots((size_t)m.T * 4);
    for (size_t i = 0; i < eSlots.size(); ++i) eSlots[i] = vertexToSlot[m.edges[i]];
    for (size_t i = 0; i < tSlots.size(); ++i) tSlots[i] = vertexToSlot[m.tets[i]];
    const CSet sets[2] = {{eSlots.data(), m.E, 2}, {tSlots.data(), m.T, 4}};
    std::vector<std::vector<uint32_t>> tileOfS(K, std::vector<uint32_t>(m.V));
    for (uint32_t p = 0; p < K; ++p)
      for (uint32_t s = 0; s < m.V; ++s) tileOfS[p][s] = tileOfV[p][slotToVertex[s]];

    PBD_PLAN_STAGE("partitions");
    // ---- assign every constraint to one admissible phase, balancing the per-vertex load
    std::vector<std::vector<TileBuild>> mainPh(K);
    for (uint32_t p = 0; p < K; ++p) mainPh[p].resize(nTilesMax);
    std::vector<uint32_t> resid[2];
    std::vector<uint16_t> load((size_t)m.V * K);
    std::vector<uint8_t> maskA[2], phaseA[2];
    // tile balance: a phase lasts as long as its fullest tile.  Move constraints out of tiles that
    // hold more than the average into emptier admissible tiles, never raising a vertex load above
    // the peak the passes above settled on.  Then hand the constraints to their tiles.
    auto finish_type = [&](int ty) {
      const CSet& cs = sets[ty];
      std::vector<uint8_t>& mask = maskA[ty];
      std::vector<uint8_t>& phaseOf = phaseA[ty];
      {
        // cap = the load 99 % of the (vertex, phase) pairs stay within: balancing must not turn
        // the rare peak into the norm (colours follow the per-tile peak)
        uint32_t peak = 0;
        {
          std::vector<uint64_t> hist(64, 0);
          uint64_t n99 = 0, acc = 0;
          for (size_t i = 0; i < load.size(); ++i)
            if (load[i]) { hist[std::min<uint32_t>(load[i], 63u)]++; ++n99; }
          n99 = n99 - n99 / 100;
          for (uint32_t l = 0; l < 64; ++l) { acc += hist[l]; if (acc >= n99) { peak = l; break; } }
          // one below that: a tile whose vertices all sit at the 99 % load needs ~2 more colours
          const uint32_t margin = (uint32_t)std::max(0, knobs().capMargin);
          peak = peak > margin ? peak - margin : 0u;
        }
        std::vector<std::vector<uint32_t>> cnt(K, std::vector<uint32_t>(nTilesMax, 0));
        uint64_t total = 0;
        for (uint32_t k = 0; k < cs.n; ++k)
          if (mask[k]) { cnt[phaseOf[k]][tileOfS[phaseOf[k]][cs.at(k)[0]]]++; ++total; }
        const uint32_t mean = (uint32_t)(total / std::max<uint64_t>(1, (uint64_t)K * nTilesMax));
        for (int sweep = 0; sweep < (knobs().tileBalance ? 8 : 0); ++sweep) {
          uint32_t moves = 0;
          for (uint32_t k = 0; k < cs.n; ++k) {
            if (mask[k] == 0 || (mask[k] & (mask[k] - 1)) == 0) continue;
            const uint32_t* id = cs.at(k);
            const uint32_t p0 = phaseOf[k], t0 = tileOfS[p0][id[0]];
            if (cnt[p0][t0] <= mean) continue;
            uint32_t bestP = p0, bestCnt = cnt[p0][t0] - 1;   // must end strictly emptier than the source is now
            for (uint32_t p = 0; p < K; ++p) {
              if (p == p0 || !(mask[k] >> p & 1)) continue;
              uint32_t mxl = 0;
              for (uint32_t j = 0; j < cs.arity; ++j) mxl = std::max<uint32_t>(mxl, load[(size_t)id[j] * K + p] + 1u);
              const uint32_t c = cnt[p][tileOfS[p][id[0]]];
              if (mxl <= peak && c < bestCnt) { bestCnt = c; bestP = p; }
            }
            if (bestP != p0) {
              for (uint32_t j = 0; j < cs.arity; ++j) { load[(size_t)id[j] * K + p0]--; load[(size_t)id[j] * K + bestP]++; }
              cnt[p0][t0]--;
              cnt[bestP][tileOfS[bestP][id[0]]]++;
              phaseOf[k] = (uint8_t)bestP;
              ++moves;
            }
          }
          if ((uint64_t)moves * 4000u < cs.n) break;   // converged (or nearly: the tail of a sweep moves a handful)
        }
      }
      for (uint32_t k = 0; k < cs.n; ++k)
        if (mask[k]) mainPh[phaseOf[k]][tileOfS[phaseOf[k]][cs.at(k)[0]]].ty[ty].cons.push_back(k);
      if (knobs().debug) {
        // forced load: constraints with a single admissible phase
        std::vector<uint16_t> forced((size_t)m.V * K, 0);
        for (uint32_t k = 0; k < cs.n; ++k) {
          if (mask[k] == 0 || (mask[k] & (mask[k] - 1)) != 0) continue;
          const uint32_t p = (uint32_t)__builtin_ctz(mask[k]);
          for (uint32_t j = 0; j < cs.arity; ++j) forced[(size_t)cs.at(k)[j] * K + p]++;
        }
        uint32_t mxF = 0, mxL = 0;
        for (size_t i = 0; i < forced.size(); ++i) { mxF = std::max<uint32_t>(mxF, forced[i]); mxL = std::max<uint32_t>(mxL, load[i]); }
        fprintf(stderr, "[plan] type %d max forced load %u, max load %u\n", ty, mxF, mxL);
      }
    };
    // Mixed steps: the step count of a visit follows the JOINT load of its most loaded vertex.  The
    // per-type passes above moved tets only (on top of fixed edge loads); here every constraint
    // of either type that sits on an overloaded (vertex, phase) may move, one level at a time.
    auto joint_repair = [&]() {
      if (!knobs().jointRepair) return;
      std::vector<uint32_t> incOff((size_t)m.V + 1, 0), inc;   // vertex -> (type << 31 | constraint)
      for (int ty = 0; ty < 2; ++ty)
        for (uint32_t k = 0; k < sets[ty].n; ++k)
          if (maskA[ty][k]) for (uint32_t j = 0; j < sets[ty].arity; ++j) incOff[sets[ty].at(k)[j] + 1]++;
      for (uint32_t v = 0; v < m.V; ++v) incOff[v + 1] += incOff[v];
      inc.resize(incOff[m.V]);
      {
        std::vector<uint32_t> cur(incOff.begin(), incOff.end() - 1);
        for (int ty = 0; ty < 2; ++ty)
          for (uint32_t k = 0; k < sets[ty].n; ++k)
            if (maskA[ty][k]) for (uint32_t j = 0; j < sets[ty].arity; ++j) inc[cur[sets[ty].at(k)[j]]++] = ((uint32_t)ty << 31) | k;
      }
      auto peak_at = [&](int ty, uint32_t k, uint32_t p) {
        uint32_t mxl = 0;
        for (uint32_t j = 0; j < sets[ty].arity; ++j) mxl = std::max<uint32_t>(mxl, load[(size_t)sets[ty].at(k)[j] * K + p]);
        return mxl;
      };
      auto move_to = [&](int ty, uint32_t k, uint32_t p) {
        const uint32_t* id = sets[ty].at(k);
        for (uint32_t j = 0; j < sets[ty].arity; ++j) { load[(size_t)id[j] * K + phaseA[ty][k]]--; load[(size_t)id[j] * K + p]++; }
        phaseA[ty][k] = (uint8_t)p;
      };
      auto movable = [&](int ty, uint32_t k) { const uint8_t mk = maskA[ty][k]; return mk != 0 && (mk & (mk - 1)) != 0; };
      auto room_for = [&](int ty, uint32_t k, uint32_t target, uint32_t avoid) {
        for (uint32_t p = 0; p < K; ++p)
          if (p != phaseA[ty][k] && p != avoid && (maskA[ty][k] >> p & 1) && peak_at(ty, k, p) + 1u <= target) return p;
        return NONE;
      };
      uint32_t mxL = 0;
      for (size_t i = 0; i < load.size(); ++i) mxL = std::max<uint32_t>(mxL, load[i]);
      for (uint32_t target = mxL ? mxL - 1 : 0; target >= 1; --target) {
        for (int pass = 0; pass < 3; ++pass) {
          uint32_t moves = 0;
          for (int ty = 0; ty < 2; ++ty)
            for (uint32_t k = 0; k < sets[ty].n; ++k) {
              if (!movable(ty, k) || peak_at(ty, k, phaseA[ty][k]) <= target) continue;
              uint32_t p = room_for(ty, k, target, NONE);
              if (p == NONE) {
                // make room: a phase where exactly one vertex of k is full, and a constraint (either type) there that can leave
                for (uint32_t q = 0; q < K && p == NONE; ++q) {
                  if (q == phaseA[ty][k] || !(maskA[ty][k] >> q & 1)) continue;
                  uint32_t full = NONE, nFull = 0;
                  for (uint32_t j = 0; j < sets[ty].arity; ++j)
                    if (load[(size_t)sets[ty].at(k)[j] * K + q] + 1u > target) { full = sets[ty].at(k)[j]; ++nFull; }
                  if (nFull != 1 || load[(size_t)full * K + q] != target) continue;
                  for (uint32_t a = incOff[full]; a < incOff[full + 1]; ++a) {
                    const int ty2 = (int)(inc[a] >> 31);
                    const uint32_t k2 = inc[a] & 0x7fffffffu;
                    if ((ty2 == ty && k2 == k) || phaseA[ty2][k2] != q || !movable(ty2, k2)) continue;
                    const uint32_t p2 = room_for(ty2, k2, target, phaseA[ty][k]);
                    if (p2 == NONE) continue;
                    move_to(ty2, k2, p2);
                    if (peak_at(ty, k, q) + 1u <= target) { p = q; break; }
                  }
                }
              }
              if (p != NONE) { move_to(ty, k, p); ++moves; }
            }
          if (!moves) break;
        }
        size_t above = 0;
        for (size_t i = 0; i < load.size(); ++i) above += load[i] > target;
        if (above > (size_t)K * nTilesMax) break;
      }
    };
    // PBD_ORDER_RIDING: riderOf[2 * tet + slot] = edge riding on that tet (NONE = free slot),
    // hostOf[edge] = its host tet (NONE = a free edge, scheduled like any edge of the interleaved order)
    std::vector<uint32_t> riderOf, hostOf;
    auto assign_type = [&](int ty, bool onTop) {
      const CSet& cs = sets[ty];
      // mixed steps: a visit's step count follows the JOINT load of a vertex, so the second type is
      // balanced on top of the loads the first one already placed
      if (!onTop) std::fill(load.begin(), load.end(), (uint16_t)0);
      std::vector<uint8_t>& mask = maskA[ty];
      std::vector<uint8_t>& phaseOf = phaseA[ty];
      mask.assign(cs.n, 0);
      phaseOf.assign(cs.n, 0);
      std::vector<uint32_t> bucket[kMaxPartitions + 1];
      for (uint32_t k = 0; k < cs.n; ++k) {
        const uint32_t* id = cs.at(k);
        uint8_t mk = 0;
        for (uint32_t p = 0; p < K; ++p) {
          const uint32_t t = tileOfS[p][id[0]];
          bool same = true;
          for (uint32_t j = 1; j < cs.arity; ++j) same &= tileOfS[p][id[j]] == t;
          if (same) mk |= (uint8_t)(1u << p);
        }
        if (ty == 0 && riding && hostOf[k] != NONE) { mask[k] = 0; continue; }   // a rider: it goes where its host tet goes
        mask[k] = mk;
        bucket[__builtin_popcount(mk)].push_back(k);
      }
      resid[ty] = bucket[0];
      if (knobs().debug) {
        fprintf(stderr, "[plan] type %d admissible-phase popcount histogram:", ty);
        for (uint32_t f = 0; f <= K; ++f) fprintf(stderr, " %zu", bucket[f].size());
        fprintf(stderr, "\n");
      }
      PBD_PLAN_STAGE("  masks");
      // least flexible first; the flexible ones then fill the valleys
      for (uint32_t f = 1; f <= K; ++f)
        for (uint32_t k : bucket[f]) {
          const uint32_t* id = cs.at(k);
          uint32_t bestP = 0, bestMax = NONE, bestSum = NONE;
          for (uint32_t p = 0; p < K; ++p) {
            if (!(mask[k] >> p & 1)) continue;
            uint32_t mxl = 0, sum = 0;
            for (uint32_t j = 0; j < cs.arity; ++j) {
              const uint32_t l = load[(size_t)id[j] * K + p];
              mxl = std::max(mxl, l);
              sum += l;
            }
            if (mxl < bestMax || (mxl == bestMax && sum < bestSum)) { bestP = p; bestMax = mxl; bestSum = sum; }
          }
          for (uint32_t j = 0; j < cs.arity; ++j) load[(size_t)id[j] * K + bestP]++;
          phaseOf[k] = (uint8_t)bestP;
        }
      PBD_PLAN_STAGE("  greedy");
      // potential descent: move a constraint to another admissible phase when that lowers the sum of
      // squared vertex loads -- evens the loads out where the min-max rule below sees only plateaus
      for (int sweep = 0; sweep < (knobs().potential ? 12 : 0); ++sweep) {
        uint32_t moves = 0;
        for (uint32_t k = 0; k < cs.n; ++k) {
          if (mask[k] == 0 || (mask[k] & (mask[k] - 1)) == 0) continue;   // residual or forced
          const uint32_t* id = cs.at(k);
          const uint32_t p0 = phaseOf[k];
          int32_t here = 0;
          for (uint32_t j = 0; j < cs.arity; ++j) here += load[(size_t)id[j] * K + p0];
          uint32_t bestP = p0;
          int32_t bestGain = 0;
          for (uint32_t p = 0; p < K; ++p) {
            if (p == p0 || !(mask[k] >> p & 1)) continue;
            int32_t there = 0;
            for (uint32_t j = 0; j < cs.arity; ++j) there += load[(size_t)id[j] * K + p];
            const int32_t gain = here - there - (int32_t)cs.arity;   // = -(delta of the sum of squares) / 2
            if (gain > bestGain) { bestGain = gain; bestP = p; }
          }
          if (bestP != p0) {
            for (uint32_t j = 0; j < cs.arity; ++j) { load[(size_t)id[j] * K + p0]--; load[(size_t)id[j] * K + bestP]++; }
            phaseOf[k] = (uint8_t)bestP;
            ++moves;
          }
        }
        if ((uint64_t)moves * 4000u < cs.n) break;   // converged (or nearly: the tail of a sweep moves a handful)
      }
      PBD_PLAN_STAGE("  potential");
      // local search: move a constraint to another admissible phase when that lowers the larger
      // of the two peak loads involved (a few sweeps; deterministic)
      for (int sweep = 0; sweep < 6; ++sweep) {
        uint32_t moves = 0;
        for (uint32_t k = 0; k < cs.n; ++k) {
          if (mask[k] == 0 || (mask[k] & (mask[k] - 1)) == 0) continue;   // residual or forced
          const uint32_t* id = cs.at(k);
          const uint32_t p0 = phaseOf[k];
          uint32_t cur = 0;
          for (uint32_t j = 0; j < cs.arity; ++j) cur = std::max<uint32_t>(cur, load[(size_t)id[j] * K + p0]);
          uint32_t bestP = p0, bestMax = cur;
          for (uint32_t p = 0; p < K; ++p) {
            if (p == p0 || !(mask[k] >> p & 1)) continue;
            uint32_t mxl = 0;
            for (uint32_t j = 0; j < cs.arity; ++j) mxl = std::max<uint32_t>(mxl, load[(size_t)id[j] * K + p] + 1u);
            if (mxl < bestMax) { bestMax = mxl; bestP = p; }
          }
          if (bestP != p0) {
            for (uint32_t j = 0; j < cs.arity; ++j) { load[(size_t)id[j] * K + p0]--; load[(size_t)id[j] * K + bestP]++; }
            phaseOf[k] = (uint8_t)bestP;
            ++moves;
          }
        }
        if ((uint64_t)moves * 4000u < cs.n) break;   // converged (or nearly: the tail of a sweep moves a handful)
      }
      PBD_PLAN_STAGE("  local search");
      // peak repair: the colour count of a tile visit follows its most loaded vertex, and after the
      // passes above only a few vertices per tile sit above the rest.  Lower the ceiling one level
      // at a time: every constraint on an overloaded (vertex, phase) moves to an admissible phase
      // with room at all its vertices, if need be after making room there by moving ONE other
      // constraint away.  Stops at the level that leaves more than one such vertex per tile visit.
      if (knobs().repair) {
        std::vector<uint32_t> incOff((size_t)m.V + 1, 0), inc;
        for (uint32_t k = 0; k < cs.n; ++k)
          if (mask[k]) for (uint32_t j = 0; j < cs.arity; ++j) incOff[cs.at(k)[j] + 1]++;
        for (uint32_t v = 0; v < m.V; ++v) incOff[v + 1] += incOff[v];
        inc.resize(incOff[m.V]);
        {
          std::vector<uint32_t> cur(incOff.begin(), incOff.end() - 1);
          for (uint32_t k = 0; k < cs.n; ++k)
            if (mask[k]) for (uint32_t j = 0; j < cs.arity; ++j) inc[cur[cs.at(k)[j]]++] = k;
        }
        auto peak_at = [&](uint32_t k, uint32_t p) {
          uint32_t mxl = 0;
          for (uint32_t j = 0; j < cs.arity; ++j) mxl = std::max<uint32_t>(mxl, load[(size_t)cs.at(k)[j] * K + p]);
          return mxl;
        };
        auto move_to = [&](uint32_t k, uint32_t p) {
          const uint32_t* id = cs.at(k);
          for (uint32_t j = 0; j < cs.arity; ++j) { load[(size_t)id[j] * K + phaseOf[k]]--; load[(size_t)id[j] * K + p]++; }
          phaseOf[k] = (uint8_t)p;
        };
        auto movable = [&](uint32_t k) { return mask[k] != 0 && (mask[k] & (mask[k] - 1)) != 0; };
        // a phase other than `avoid` where k fits under `target`
        auto room_for = [&](uint32_t k, uint32_t target, uint32_t avoid) {
          for (uint32_t p = 0; p < K; ++p)
            if (p != phaseOf[k] && p != avoid && (mask[k] >> p & 1) && peak_at(k, p) + 1u <= target) return p;
          return NONE;
        };
        uint32_t mxL = 0;
        for (size_t i = 0; i < load.size(); ++i) mxL = std::max<uint32_t>(mxL, load[i]);
        for (uint32_t target = mxL ? mxL - 1 : 0; target >= 1; --target) {
          for (int pass = 0; pass < 3; ++pass) {
            uint32_t moves = 0;
            for (uint32_t k = 0; k < cs.n; ++k) {
              if (!movable(k) || peak_at(k, phaseOf[k]) <= target) continue;
              uint32_t p = room_for(k, target, NONE);
              if (p == NONE) {
                // make room: one phase where exactly one vertex of k is full, and a constraint there that can leave
                for (uint32_t q = 0; q < K && p == NONE; ++q) {
                  if (q == phaseOf[k] || !(mask[k] >> q & 1)) continue;
                  uint32_t full = NONE, nFull = 0;
                  for (uint32_t j = 0; j < cs.arity; ++j)
                    if (load[(size_t)cs.at(k)[j] * K + q] + 1u > target) { full = cs.at(k)[j]; ++nFull; }
                  if (nFull != 1 || load[(size_t)full * K + q] != target) continue;
                  for (uint32_t a = incOff[full]; a < incOff[full + 1]; ++a) {
                    const uint32_t k2 = inc[a];
                    if (k2 == k || phaseOf[k2] != q || !movable(k2)) continue;
                    const uint32_t p2 = room_for(k2, target, phaseOf[k]);
                    if (p2 == NONE) continue;
                    move_to(k2, p2);
                    if (peak_at(k, q) + 1u <= target) { p = q; break; }
                  }
                }
              }
              if (p != NONE) { move_to(k, p); ++moves; }
            }
            if (!moves) break;
          }
          // forced constraints can pin a few vertices above any level; go on while fewer than one
          // (vertex, phase) pair per tile visit is left above this one
          size_t above = 0;
          for (size_t i = 0; i < load.size(); ++i) above += load[i] > target;
          if (above > (size_t)K * nTilesMax) break;
        }
      }
      PBD_PLAN_STAGE("  peak repair");
      if (!mixedThreads) finish_type(ty);
    };
    if (!riding) {
      assign_type(0, false);
      assign_type(1, mixedThreads && knobs().jointLoad);
    } else {
      // tets first; then every edge looks for a host among the tets that contain it (a tet that is
      // admissible in some partition: wherever the planner puts it, all its vertices -- hence the
      // edge's -- lie in one tile there).  A tet takes at most two riders, and two only if they are
      // opposite edges of it (vertex-disjoint), which is what the kernel's register-resident form
      // needs.  Edges without a host stay free and are balanced on top of the tet loads.
      assign_type(1, false);
      riderOf.assign((size_t)m.T * 2, NONE);
      hostOf.assign(m.E, NONE);
      {
        // (vertex pair) -> tets containing it: sorted list of (lo << 32 | hi, tet * 8 + pair code)
        static const uint8_t pr[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};   // code c and 5 - c are opposite
        std::vector<std::pair<uint64_t, uint32_t>> inc;
        inc.reserve((size_t)m.T * 6);
        for (uint32_t t = 0; t < m.T; ++t) {
          if (maskA[1][t] == 0) continue;   // residual tets do not host
          const uint32_t* id = sets[1].at(t);
          for (uint32_t c = 0; c < 6; ++c) {
            const uint32_t a = id[pr[c][0]], b = id[pr[c][1]];
            if (a == b) continue;
            inc.emplace_back(((uint64_t)std::min(a, b) << 32) | std::max(a, b), t * 8u + c);
          }
        }
        std::sort(inc.begin(), inc.end());
        auto range_of = [&](uint32_t e, size_t& lo, size_t& hi) {
          const uint32_t a = sets[0].at(e)[0], b = sets[0].at(e)[1];
          const uint64_t key = ((uint64_t)std::min(a, b) << 32) | std::max(a, b);
          lo = std::lower_bound(inc.begin(), inc.end(), std::make_pair(key, 0u)) - inc.begin();
          hi = lo;
          while (hi < inc.size() && inc[hi].first == key) ++hi;
        };
        // passes 0, 1: a host without a rider yet (edges with <= 2 candidates first, then the rest);
        // pass 2: a host whose rider is the opposite edge
        std::vector<uint8_t> codeOf((size_t)m.T * 2, 0);
        for (int pass = 0; pass < (knobs().riders >= 2 ? 3 : 2); ++pass)
          for (uint32_t e = 0; e < m.E; ++e) {
            if (hostOf[e] != NONE || sets[0].at(e)[0] == sets[0].at(e)[1]) continue;
            size_t lo, hi;
            range_of(e, lo, hi);
            if (pass == 0 && hi - lo > 2) continue;
            for (size_t i = lo; i < hi && hostOf[e] == NONE; ++i) {
              const uint32_t t = inc[i].second >> 3, c = inc[i].second & 7u;
              if (pass < 2) {
                if (riderOf[2 * (size_t)t] == NONE) { riderOf[2 * (size_t)t] = e; codeOf[2 * (size_t)t] = (uint8_t)c; hostOf[e] = t; }
              } else if (riderOf[2 * (size_t)t] != NONE && riderOf[2 * (size_t)t + 1] == NONE && codeOf[2 * (size_t)t] + c == 5u) {
                riderOf[2 * (size_t)t + 1] = e; codeOf[2 * (size_t)t + 1] = (uint8_t)c; hostOf[e] = t;
              }
            }
          }
        uint32_t nRide = 0;
        for (uint32_t e = 0; e < m.E; ++e) nRide += hostOf[e] != NONE;
        plan.riders = nRide;
        if (knobs().debug) fprintf(stderr, "[plan] riding: %u of %u edges ride on a tet, %u stay free\n", nRide, m.E, m.E - nRide);
      }
      PBD_PLAN_STAGE("  riders");
      assign_type(0, true);
    }
    // tagged hand-over (experimental): a shifted tile carries every vertex of its partition cell, touched
    // by a constraint of this phase or not, so that every phase rewrites every vertex
    if (opts.flags & PBD_FLAG_TAGGED_HANDOVER)
      for (uint32_t p = 1; p < K; ++p) {
        for (auto& tb : mainPh[p]) { tb.presetVerts = true; tb.verts.clear(); }
        for (uint32_t sl = 0; sl < m.V; ++sl) mainPh[p][tileOfS[p][sl]].verts.push_back(sl);   // ascending
      }
    if (mixedThreads) {
      joint_repair();
      PBD_PLAN_STAGE("  joint repair");
      finish_type(0);
      finish_type(1);
      if (riding)
        for (uint32_t p = 0; p < K; ++p)
          for (auto& tb : mainPh[p]) {
            tb.nRiders = 0;
            for (uint32_t k : tb.ty[1].cons) tb.nRiders += (riderOf[2 * (size_t)k] != NONE) + (riderOf[2 * (size_t)k + 1] != NONE);
          }
      PBD_PLAN_STAGE("  tile balance");
    }
    PBD_PLAN_STAGE("assignment");
    // shared-memory placement search after the colouring (pbd_placement.cpp): one thread per constraint only (the 2- and
    // 4-lane tet sweeps and the riding order address shared memory differently); bodies beyond ~2.5M tets skip it unless
    // PBD_PLAN_PLACE asks for it -- it costs ~2 s per million tets on 8 host threads for the few percent the bank
    // conflicts are worth, and every rank of a sharded body plans the whole body
    const bool bigBody = (uint64_t)m.T + m.E > 6000000ull && !getenv("PBD_PLAN_PLACE");
    const int placeEffort = (opts.lanes_per_tet <= 1 && !riding && !bigBody) ? knobs().place : 0;
    const bool willPlace = placeEffort > 0;
    // ---- finish the main tiles (independent of each other: spread over host threads; the result
    // does not depend on the thread count)
    bool fits = true;
    for (uint32_t t = 0; t < nTile0 && t < mainPh[0].size(); ++t) {
      TileBuild& tb = mainPh[0][t];
      tb.contiguous = true;
      tb.rangeBegin = tile0Begin[t];
      tb.rangeCount = tile0Begin[t + 1] - tile0Begin[t];
    }
    // Cheap necessary condition first: a tile's vertices and constraint records (with the smallest
    // possible colour-group tables) must fit.  A tile count that fails it is abandoned before any
    // colouring (a failed attempt used to cost as much as the successful one).
    {
      std::vector<uint32_t> stamp(m.V, NONE);
      uint32_t tileId = 0;
      for (uint32_t p = 0; p < K && fits; ++p)
        for (auto& tb : mainPh[p]) {
          uint32_t nv = tb.rangeCount;
          if (!tb.contiguous) {
            nv = 0;
            for (int ty = 0; ty < 2; ++ty)
              for (uint32_t k : tb.ty[ty].cons)
                for (uint32_t j = 0; j < sets[ty].arity; ++j) {
                  const uint32_t sl = sets[ty].at(k)[j];
                  if (stamp[sl] != tileId) { stamp[sl] = tileId; ++nv; }
                }
          }
          ++tileId;
          const uint32_t nE = (uint32_t)tb.ty[0].cons.size() + tb.nRiders, nT = (uint32_t)tb.ty[1].cons.size();
          const uint32_t rec = tile_record_bytes(tb.contiguous ? 0u : nv, nE ? 1u : 0u, nT ? 1u : 0u, nE, nT, riding);
          if (nv > 65535u || 16u * nv + 2u * rec > smemBytes) { fits = false; break; }
        }
    }
    if (fits) {
      std::vector<TileBuild*> work;
      for (uint32_t p = 0; p < K; ++p)
        for (auto& tb : mainPh[p]) work.push_back(&tb);
      std::atomic<size_t> next{0};
      std::atomic<bool> ok{true};
      auto worker = [&](std::vector<uint32_t>& lo, std::vector<uint32_t>& sc) {
        for (size_t i; (i = next.fetch_add(1)) < work.size();) {
          if (!ok.load(std::memory_order_relaxed)) break;   // some tile does not fit: this attempt is void anyway
          TileBuild& tb = *work[i];
          finish_tile(sets, tb, lo, sc, caps, mixedThreads, !willPlace);
          const uint32_t nv = tb.contiguous ? tb.rangeCount : (uint32_t)tb.verts.size();
          if (nv > 65535u || tile_bytes(tb, riding) > smemBytes) ok = false;
        }
      };
      const unsigned nThreads = (m.T + m.E < 200000u) ? 1u : std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
      std::vector<std::thread> pool;
      std::vector<std::vector<uint32_t>> los(nThreads > 1 ? nThreads - 1 : 0, std::vector<uint32_t>(m.V, NONE)), scs(los.size());
      for (size_t i = 0; i < los.size(); ++i) pool.emplace_back(worker, std::ref(los[i]), std::ref(scs[i]));
      worker(localOf, scratch);
      for (auto& th : pool) th.join();
      fits = ok;
    }
    // second colouring pass: a phase lasts as long as its slowest tile, so tiles that ended above
    // what most tiles of their phase reached get a longer search (again spread over host threads)
    if (fits) {
      struct Job { TileBuild* tb; int ty; uint32_t goal; };   // ty == 2: the joint colouring of a mixed tile
      std::vector<Job> jobs;
      for (uint32_t p = 0; p < K; ++p)
        for (int ty = 0; ty < 3; ++ty) {
          auto takes = [&](const TileBuild& tb) {
            return ty == 2 ? tb.mixed : (!tb.mixed && !tb.ty[ty].cons.empty());
          };
          std::vector<uint32_t> ncs;
          for (auto& tb : mainPh[p])
            if (takes(tb)) ncs.push_back(tb.ty[ty & 1].nColours);
          if (ncs.size() < 4) continue;
          std::sort(ncs.begin(), ncs.end());
          const uint32_t goal = ncs[ncs.size() / 4];   // lower quartile
          for (auto& tb : mainPh[p])
            if (takes(tb) && tb.ty[ty & 1].nColours > goal) jobs.push_back({&tb, ty, goal});
        }
      std::atomic<size_t> next{0};
      auto worker = [&](std::vector<uint32_t>& lo, std::vector<uint32_t>& sc) {
        for (size_t i; (i = next.fetch_add(1)) < jobs.size();) {
          TileBuild& tb = *jobs[i].tb;
          const int ty = jobs[i].ty;
          const uint32_t goal = jobs[i].goal;
          TypeList& L = tb.ty[ty & 1];
          uint32_t nLocal;
          if (tb.contiguous) {
            nLocal = tb.rangeCount;
            for (uint32_t q = 0; q < nLocal; ++q) lo[tb.rangeBegin + q] = q;
          } else {
            nLocal = (uint32_t)tb.verts.size();
            for (uint32_t q = 0; q < nLocal; ++q) lo[tb.verts[q]] = q;
          }
          if (ty == 2) {
            for (uint32_t attempt2 = 0; attempt2 < 4 && L.nColours > goal; ++attempt2) {
              TileBuild trial;
              trial.ty[0].cons = tb.ty[0].cons;
              trial.ty[1].cons = tb.ty[1].cons;
              colour_joint(sets, trial, nLocal, lo, sc, mixedThreads, 160, goal, 0x85ebca6bu * (attempt2 + 1));
              if (trial.ty[0].nColours < L.nColours && steps_fit(trial, mixedThreads)) {
                tb.ty[0] = std::move(trial.ty[0]);
                tb.ty[1] = std::move(trial.ty[1]);
              }
            }
            if (!willPlace) { order_groups(sets, tb, lo, 0); order_groups(sets, tb, lo, 1); }
            continue;
          }
          for (uint32_t attempt2 = 0; attempt2 < 4 && L.nColours > goal; ++attempt2) {
            TypeList trial;
            trial.cons = L.cons;
            colour_list(sets[ty], trial, nLocal, lo, sc, caps[ty], 160, goal, 0x85ebca6bu * (attempt2 + 1));
            if (trial.nColours < L.nColours) L = std::move(trial);
          }
          if (!willPlace) order_groups(sets, tb, lo, ty);
        }
      };
      // two jobs may share a tile (its edge list and its tet list): they touch different TypeLists
      // and write the same values into the thread-private numbering, so they are independent
      const unsigned nThreads = jobs.size() < 8 ? 1u : std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
      std::vector<std::thread> pool;
      std::vector<std::vector<uint32_t>> los(nThreads > 1 ? nThreads - 1 : 0, std::vector<uint32_t>(m.V, NONE)), scs(los.size());
      for (size_t i = 0; i < los.size(); ++i) pool.emplace_back(worker, std::ref(los[i]), std::ref(scs[i]));
      worker(localOf, scratch);
      for (auto& th : pool) th.join();
    }
    if (!fits) {
      uint32_t next = K1 + (K1 + 1) / 2;
      if (next > nSMs) next = ((next + nSMs - 1) / nSMs) * nSMs;   // whole waves
      K1 = std::max(K1 + 1, next);
      continue;
    }

    PBD_PLAN_STAGE("tile colouring");
    // ---- shared-memory placement of the main tiles (one thread per constraint only: the 2- and 4-lane tet
    // sweeps and the riding order address shared memory differently).  Again independent per tile.
    // Home tiles first: their vertices ARE a slot range, so their new numbering is a renumbering of the slots
    // inside each range, applied to everything that names a slot; then the shifted tiles, whose vertex lists
    // are re-sorted by the new slots first.  A vertex moves only inside an aligned block of its tile's indices
    // (home: 8 = one 128-byte line of vertex words, shifted: 32 = one warp): measured without that restriction,
    // the scattered global loads / stores of the tile visits cost more than the bank conflicts saved.
    {
      const int effort = placeEffort;
      // fast arithmetic: a tet's vertices may take its four roles in any order (same classes, same rows of edges, same
      // slot numbering as the exact mode's plan -- only the tets' rows and roles differ)
      const bool relabel = (opts.flags & PBD_FLAG_FAST_ARITH) != 0u && knobs().relabel;
      const unsigned nThreads = (m.T + m.E < 200000u) ? 1u : std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
      std::vector<PlaceStats> st(nThreads);
      std::vector<std::vector<uint32_t>> los(nThreads > 1 ? nThreads - 1 : 0, std::vector<uint32_t>(m.V, NONE));
      auto place_phases = [&](uint32_t pBegin, uint32_t pEnd, uint32_t block) {
        std::vector<TileBuild*> work;
        for (uint32_t p = pBegin; p < pEnd; ++p)
          for (auto& tb : mainPh[p])
            if (!tb.ty[0].cons.empty() || !tb.ty[1].cons.empty()) work.push_back(&tb);
        std::atomic<size_t> next{0};
        auto worker = [&](std::vector<uint32_t>& lo, PlaceStats& ps) {
          for (size_t i; (i = next.fetch_add(1)) < work.size();) place_tile(sets, *work[i], lo, effort, block, relabel, ps);
        };
        std::vector<std::thread> pool;
        for (size_t i = 0; i < los.size(); ++i) pool.emplace_back(worker, std::ref(los[i]), std::ref(st[i + 1]));
        worker(localOf, st[0]);
        for (auto& th : pool) th.join();
      };
      place_phases(0, 1, (uint32_t)knobs().placeBlockHome);
      std::vector<uint32_t> slotPerm;
      for (uint32_t t = 0; t < nTile0 && t < mainPh[0].size(); ++t) {
        const TileBuild& tb = mainPh[0][t];
        if (tb.localPerm.empty()) continue;
        if (slotPerm.empty()) { slotPerm.resize(m.V); std::iota(slotPerm.begin(), slotPerm.end(), 0u); }
        for (uint32_t i = 0; i < tb.rangeCount; ++i) slotPerm[tb.rangeBegin + i] = tb.rangeBegin + tb.localPerm[i];
      }
      if (!slotPerm.empty()) {
        std::vector<uint32_t> moved(m.V);
        for (uint32_t sl = 0; sl < m.V; ++sl) moved[slotPerm[sl]] = slotToVertex[sl];
        slotToVertex.swap(moved);
        for (uint32_t sl = 0; sl < m.V; ++sl) vertexToSlot[slotToVertex[sl]] = sl;
        for (auto& sl : eSlots) sl = slotPerm[sl];
        for (auto& sl : tSlots) sl = slotPerm[sl];
        for (uint32_t p = 0; p < K; ++p) {
          for (uint32_t sl = 0; sl < m.V; ++sl) moved[slotPerm[sl]] = tileOfS[p][sl];
          tileOfS[p].swap(moved);
          moved.resize(m.V);
          for (auto& tb : mainPh[p]) {
            for (auto& sl : tb.verts) sl = slotPerm[sl];
            std::sort(tb.verts.begin(), tb.verts.end());   // (placement of the shifted tiles starts from ascending slots)
          }
        }
      }
      if (K > 1) place_phases(1, K, (uint32_t)knobs().placeBlockShifted);
      for (const PlaceStats& ps : st)
        for (int ty = 0; ty < 2; ++ty) { plan.gatherWavefronts[ty] += ps.wavefronts[ty]; plan.gatherIdeal[ty] += ps.ideal[ty]; }
    }
    PBD_PLAN_STAGE("placement");
    // ---- residual phases (constraints interior to no partition)
    const uint32_t avgTile = std::max(64u, (m.V + nTile0 - 1) / std::max(1u, nTile0));
    uint32_t resCap = std::min(65535u, std::min(2u * avgTile, smemBytes / 128u));
    std::vector<std::vector<TileBuild>> resPh[2];
    for (;;) {
      bool ok = true;
      for (int ty = 0; ty < 2; ++ty) {
        resPh[ty].clear();
        build_residual_phases(sets, ty, resid[ty], m.V, resCap, resPh[ty], localOf, scratch, caps);
        for (auto& ph : resPh[ty])
          for (auto& tb : ph)
            if (tile_bytes(tb, riding) > smemBytes) ok = false;
      }
      if (ok) break;
      if (resCap <= 16) { err = "a residual tile does not fit in shared memory"; return false; }
      resCap /= 2;   // high-valence vertices: smaller tiles carry fewer constraints
    }

    PBD_PLAN_STAGE("residual phases");
    // ---- phase list in execution order.  A "view" selects which constraint types of a tile run.
    struct PhaseRef { std::vector<TileBuild>* tiles; bool useE, useT, home; };
    std::vector<PhaseRef> seq;
    if (fused) {
      for (uint32_t p = 0; p < K; ++p) seq.push_back({&mainPh[p], true, true, p == 0});
      for (auto& ph : resPh[0]) seq.push_back({&ph, true, false, false});
      for (auto& ph : resPh[1]) seq.push_back({&ph, false, true, false});
    } else {
      if (m.E) {
        for (uint32_t p = 0; p < K; ++p) seq.push_back({&mainPh[p], true, false, p == 0});
        for (auto& ph : resPh[0]) seq.push_back({&ph, true, false, false});
      }
      if (m.T) {
        for (uint32_t p = 0; p < K; ++p) seq.push_back({&mainPh[p], false, true, p == 0});
        for (auto& ph : resPh[1]) seq.push_back({&ph, false, true, false});
      }
    }
    if (m.E == 0 && m.T == 0) seq.clear();

    // ---- flatten
    plan.slotToVertex = slotToVertex;
    plan.vertexToSlot = vertexToSlot;
    plan.tile0Begin = tile0Begin;
    plan.partitions = K;
    plan.tilesPerPartition = nTile0;
    plan.edgeOrder.reserve(m.E); plan.tetOrder.reserve(m.T);
    plan.edgeLocal.reserve((size_t)m.E * 2); plan.tetLocal.reserve((size_t)m.T * 4);
    plan.edgeDev.reserve(m.E); plan.tetDev.reserve(m.T);
    plan.edgePhase.assign(m.E, 0); plan.edgeTile.assign(m.E, 0); plan.edgeColor.assign(m.E, 0);
    plan.tetPhase.assign(m.T, 0); plan.tetTile.assign(m.T, 0); plan.tetColor.assign(m.T, 0);
    {
      bool anyPerm = false;
      for (uint32_t p = 0; p < K && !anyPerm; ++p)
        for (auto& tb : mainPh[p]) if (!tb.tetPerm.empty()) { anyPerm = true; break; }
      if (anyPerm) plan.tetPerm.assign(m.T, (uint8_t)0xE4);
    }
    for (uint32_t t = 0; t < nTile0; ++t) plan.tileVertexCapacity = std::max(plan.tileVertexCapacity, tile0Begin[t + 1] - tile0Begin[t]);
    uint32_t devCur[2] = {0, 0};
    std::vector<uint32_t> ridePos(riding ? m.E : 0u, NONE);   // rider edge -> its schedule position
    for (size_t pi = 0; pi < seq.size(); ++pi) {
      const PhaseRef& pr = seq[pi];
      Phase P;
      P.tileBegin = (uint32_t)plan.tiles.size();
      uint32_t mxCol[2] = {0, 0};
      bool anyE = false, anyT = false;
      for (size_t ti = 0; ti < pr.tiles->size(); ++ti) {
        TileBuild& tb = (*pr.tiles)[ti];
        const bool hasT = pr.useT && !tb.ty[1].cons.empty();
        const bool hasE = pr.useE && (!tb.ty[0].cons.empty() || (hasT && tb.nRiders != 0));
        const bool isHome = pr.home && pi == 0;   // the very first phase covers every slot (vertex stages are fused into it)
        if (!hasE && !hasT && !(isHome && tb.contiguous)) continue;
        Tile tl;
        tl.contiguous = tb.contiguous ? 1u : 0u;
        tl.mixed = (tb.mixed && hasE && hasT) ? 1u : 0u;
        tl.ride = (riding && hasT) ? 1u : 0u;
        uint32_t nLocal;
        if (tb.contiguous) {
          tl.vertBegin = tb.rangeBegin;
          tl.vertCount = nLocal = tb.rangeCount;
          for (uint32_t i = 0; i < nLocal; ++i) localOf[tb.rangeBegin + i] = i;
        } else {
          tl.vertBegin = (uint32_t)plan.tileVerts.size();
          tl.vertCount = nLocal = (uint32_t)tb.verts.size();
          for (uint32_t i = 0; i < nLocal; ++i) { localOf[tb.verts[i]] = i; plan.tileVerts.push_back(tb.verts[i]); }
        }
        plan.tileVertexCapacity = std::max(plan.tileVertexCapacity, nLocal);
        for (int ty = 0; ty < 2; ++ty) {
          const bool use = ty ? hasT : hasE;
          const CSet& cs = sets[ty];
          std::vector<uint32_t>& order = ty ? plan.tetOrder : plan.edgeOrder;
          std::vector<uint32_t>& dev = ty ? plan.tetDev : plan.edgeDev;
          std::vector<uint16_t>& local = ty ? plan.tetLocal : plan.edgeLocal;
          std::vector<uint32_t>& cPhase = ty ? plan.tetPhase : plan.edgePhase;
          std::vector<uint32_t>& cTile = ty ? plan.tetTile : plan.edgeTile;
          std::vector<uint32_t>& cCol = ty ? plan.tetColor : plan.edgeColor;
          uint32_t& begin = ty ? tl.tetBegin : tl.edgeBegin;
          uint32_t& count = ty ? tl.tetCount : tl.edgeCount;
          uint32_t& gBegin = ty ? tl.tetGroupBegin : tl.edgeGroupBegin;
          uint32_t& gCount = ty ? tl.tetGroupCount : tl.edgeGroupCount;
          uint32_t& devBegin = ty ? tl.tetDevBegin : tl.edgeDevBegin;
          begin = (uint32_t)order.size();
          gBegin = (uint32_t)plan.groups.size();
          devCur[ty] = pad4(devCur[ty]);
          devBegin = devCur[ty];
          if (!use) continue;
          const TypeList& L = tb.ty[ty];
          size_t i = 0;
          if (tl.mixed) {
            // mixed tile: exactly one group per step and type (possibly empty); the planner kept every step within the block
            for (uint32_t sIdx = 0; sIdx < L.nColours; ++sIdx) {
              size_t j = i;
              while (j < L.cons.size() && L.colour[j] == sIdx) ++j;
              Group g;
              g.begin = (uint32_t)order.size();
              g.count = (uint32_t)(j - i);
              plan.groups.push_back(g);
              for (size_t k = i; k < j; ++k) {
                const uint32_t c = L.cons[k];
                cPhase[c] = (uint32_t)plan.phases.size();
                cTile[c] = (uint32_t)plan.tiles.size();
                cCol[c] = sIdx;
                const uint8_t code = (ty == 1 && !tb.tetPerm.empty()) ? tb.tetPerm[k] : (uint8_t)0xE4;
                if (ty == 1 && !plan.tetPerm.empty()) plan.tetPerm[order.size()] = code;
                for (uint32_t a = 0; a < cs.arity; ++a) local.push_back((uint16_t)localOf[cs.at(c)[(code >> (2 * a)) & 3u]]);
                order.push_back(c);
                dev.push_back(devCur[ty]++);
              }
              i = j;
            }
          }
          while (i < L.cons.size()) {
            size_t j = i;
            while (j < L.cons.size() && L.colour[j] == L.colour[i]) ++j;
            // one group per colour, split so that no group exceeds what one block pass can take
            // (the sweep loops then need no second pass and stay half as long)
            const uint32_t lanes = ty ? std::max(1u, opts.lanes_per_tet) : 1u;
            const uint32_t lim = std::max(1u, blockThreads / lanes);
            const uint32_t nGrp = (uint32_t)(j - i), parts = (nGrp + lim - 1) / lim;
            for (uint32_t q = 0; q < parts; ++q) {
              Group g;
              g.begin = (uint32_t)order.size() + (uint32_t)(((uint64_t)nGrp * q) / parts);
              g.count = (uint32_t)(((uint64_t)nGrp * (q + 1)) / parts - ((uint64_t)nGrp * q) / parts);
              plan.groups.push_back(g);
            }
            for (size_t k = i; k < j; ++k) {
              const uint32_t c = L.cons[k];
              cPhase[c] = (uint32_t)plan.phases.size();
              cTile[c] = (uint32_t)plan.tiles.size();
              cCol[c] = L.colour[k];
              const uint8_t code = (ty == 1 && !tb.tetPerm.empty()) ? tb.tetPerm[k] : (uint8_t)0xE4;
              if (ty == 1 && !plan.tetPerm.empty()) plan.tetPerm[order.size()] = code;
              for (uint32_t a = 0; a < cs.arity; ++a) local.push_back((uint16_t)localOf[cs.at(c)[(code >> (2 * a)) & 3u]]);
              order.push_back(c);
              dev.push_back(devCur[ty]++);
            }
            i = j;
          }
          gCount = (uint32_t)plan.groups.size() - gBegin;
          if (ty == 0 && tl.ride && tb.nRiders) {
            // riders: behind the tile's free edges, in the order of their host tets; no colour group
            for (size_t q = 0; q < tb.ty[1].cons.size(); ++q)
              for (uint32_t sl = 0; sl < 2; ++sl) {
                const uint32_t c = riderOf[2 * (size_t)tb.ty[1].cons[q] + sl];
                if (c == NONE) continue;
                cPhase[c] = (uint32_t)plan.phases.size();
                cTile[c] = (uint32_t)plan.tiles.size();
                cCol[c] = tb.ty[1].colour.empty() ? 0u : tb.ty[1].colour[q];
                for (uint32_t a = 0; a < 2; ++a) local.push_back((uint16_t)localOf[cs.at(c)[a]]);
                ridePos[c] = (uint32_t)order.size();
                order.push_back(c);
                dev.push_back(devCur[ty]++);
              }
          }
          if (ty == 1 && riding)
            for (uint32_t q = begin; q < (uint32_t)order.size(); ++q)
              for (uint32_t sl = 0; sl < 2; ++sl) {
                const uint32_t c = tl.ride ? riderOf[2 * (size_t)order[q] + sl] : NONE;
                plan.tetRide.push_back(c == NONE ? NONE : ridePos[c]);
              }
          count = (uint32_t)order.size() - begin;
          mxCol[ty] = std::max(mxCol[ty], gCount);
          (ty ? anyT : anyE) = true;
        }
        plan.tileRecordBytes = std::max(plan.tileRecordBytes,
                                        tile_record_bytes(tl.contiguous ? 0u : tl.vertCount, tl.edgeGroupCount,
                                                          tl.tetGroupCount, tl.edgeCount, tl.tetCount, tl.ride != 0));
        plan.tiles.push_back(tl);
      }
      P.tileCount = (uint32_t)plan.tiles.size() - P.tileBegin;
      if (P.tileCount == 0) continue;
      plan.phases.push_back(P);
      plan.edgeColorSum += mxCol[0];
      plan.tetColorSum += mxCol[1];
      plan.edgePhases += anyE ? 1u : 0u;
      plan.tetPhases += anyT ? 1u : 0u;
    }
    plan.edgeDevCount = pad4(devCur[0]);
    plan.tetDevCount = pad4(devCur[1]);
    break;
  }
  PBD_PLAN_STAGE("flatten");
  plan.planMs = now_ms() - t0;
  if (knobs().debug) {
    uint64_t sum = 0;
    uint32_t mxE = 0, mxT = 0, mxV = 0;
    for (const Tile& t : plan.tiles) {
      sum += tile_record_bytes(t.contiguous ? 0u : t.vertCount, t.edgeGroupCount, t.tetGroupCount, t.edgeCount, t.tetCount, t.ride != 0);
      mxE = std::max(mxE, t.edgeCount); mxT = std::max(mxT, t.tetCount); mxV = std::max(mxV, t.vertCount);
    }
    {
      std::vector<uint32_t> hist(16, 0);
      for (const Tile& t : plan.tiles) hist[std::min<uint32_t>(15u, t.vertCount / 64u)]++;
      fprintf(stderr, "[plan] tile vertices (bins of 64):");
      for (uint32_t b = 0; b < 16; ++b) if (hist[b]) fprintf(stderr, " %u..%u: %u", 64 * b, 64 * b + 63, hist[b]);
      fprintf(stderr, "\n");
    }
    fprintf(stderr, "[plan] record block: max %u bytes, mean %.0f (largest tile: %u edges, %u tets, %u vertices; %zu tiles)\n", plan.tileRecordBytes,
            (double)sum / (double)std::max<size_t>(1, plan.tiles.size()), mxE, mxT, mxV, plan.tiles.size());
    // global side of a tile visit: thread i loads / stores the 16-byte word of the slot at tile position i
    uint64_t warps = 0, sectors = 0, lines = 0;
    for (const Tile& t : plan.tiles)
      for (uint32_t i0 = 0; i0 < t.vertCount; i0 += 32) {
        uint32_t sl[32], n = std::min(32u, t.vertCount - i0);
        for (uint32_t i = 0; i < n; ++i) sl[i] = t.contiguous ? t.vertBegin + i0 + i : plan.tileVerts[t.vertBegin + i0 + i];
        std::sort(sl, sl + n);
        uint32_t ns = 0, nl = 0;
        for (uint32_t i = 0; i < n; ++i) { ns += i == 0 || (sl[i] >> 1) != (sl[i - 1] >> 1); nl += i == 0 || (sl[i] >> 3) != (sl[i - 1] >> 3); }
        ++warps; sectors += ns; lines += nl;
      }
    fprintf(stderr, "[plan] vertex words per warp access: %.2f 32-byte sectors, %.2f 128-byte lines (16 / 4 = contiguous)\n",
            (double)sectors / (double)std::max<uint64_t>(1, warps), (double)lines / (double)std::max<uint64_t>(1, warps));
  }
  if (knobs().debug) {
    // fingerprint of everything the upload reads (FNV-1a over the plan's arrays): planner changes that claim to be
    // result-neutral (threading, data structures) are checked against it on the host, before any GPU run
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const void* d, size_t n) { const unsigned char* c = static_cast<const unsigned char*>(d); for (size_t i = 0; i < n; ++i) { h ^= c[i]; h *= 1099511628211ull; } };
    auto vec = [&](const auto& v) { const uint64_t n = v.size(); mix(&n, sizeof(n)); if (n) mix(v.data(), n * sizeof(v[0])); };
    vec(plan.edgeOrder); vec(plan.tetOrder); vec(plan.edgeColorOff); vec(plan.tetColorOff); vec(plan.slotToVertex); vec(plan.vertexToSlot);
    vec(plan.tileVerts); vec(plan.phases); vec(plan.tiles); vec(plan.groups); vec(plan.edgeLocal); vec(plan.tetLocal);
    vec(plan.edgeDev); vec(plan.tetDev); vec(plan.tile0Begin); vec(plan.tetRide); vec(plan.tetPerm);
    vec(plan.edgePhase); vec(plan.edgeTile); vec(plan.edgeColor); vec(plan.tetPhase); vec(plan.tetTile); vec(plan.tetColor);
    const uint32_t sc[] = {plan.tileVertexCapacity, plan.tileRecordBytes, plan.partitions, plan.tilesPerPartition, plan.edgeDevCount, plan.tetDevCount,
                           plan.blockThreads, plan.tilesPerSm, plan.edgePhases, plan.tetPhases, plan.edgeColorSum, plan.tetColorSum, plan.riders};
    mix(sc, sizeof(sc));
    fprintf(stderr, "[plan] fingerprint %016llx\n", (unsigned long long)h);
  }
  if (knobs().debug)
    fprintf(stderr, "[plan] shared-memory gathers, wavefronts per quarter-warp role (1.0 = conflict-free): edges %.3f, tets %.3f\n",
            (double)plan.gatherWavefronts[0] / (double)std::max<uint64_t>(1, plan.gatherIdeal[0]),
            (double)plan.gatherWavefronts[1] / (double)std::max<uint64_t>(1, plan.gatherIdeal[1]));
  return true;
}

}  // namespace pbd

// pbd_tileplan.cpp -- the "tile" schedule: multi-phase shared-memory tiles.
//
// Why: a globally coloured sweep needs one grid-wide barrier per colour (~46 per iteration on the
// Kuhn grid, 155 on default_Tet), which caps an L2-resident mesh at ~10 % of the HBM roofline
// (SURVEY.md 7 "dependent-phase count vs. bytes").  Here the vertices are partitioned into tiles
// that fit in one SM's shared memory.  A constraint whose vertices all lie in one tile is swept
// inside that tile with LOCAL colours and block barriers only.  Constraints that straddle tiles
// form a residual; the residual's own vertices are re-partitioned (graph Voronoi, so the new cuts
// fall away from the old ones), which makes most of it tile-interior again; and so on until
// nothing is left.  Per constraint type this takes a handful of grid-wide phases (4-6 on the
// Kuhn grid) instead of one per colour.
//
// Order: phase by phase, tile by tile, colour by colour, caller index inside a colour.  Tiles of
// one phase are vertex-disjoint and colour groups are conflict-free, so executing them in
// parallel equals executing them sequentially in that order: still Gauss-Seidel over the same
// set, just permuted (disclosed through pbd_get_schedule_order).  PBD_ORDER_STRICT keeps "all
// edges, then all tets" per iteration like the reference (CProgram/src/Sim.cpp:293-297).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <numeric>
#include <queue>

#include "pbd_plan.h"

namespace pbd {

namespace {

constexpr uint32_t NONE = 0xffffffffu;

double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// ------------------------------------------------------------------ phase 0: RCB vertex tiles

struct Rcb {
  const float* x;                 // 3V
  std::vector<uint32_t> idx;      // permutation being partitioned
  std::vector<uint32_t> tileOf;   // per vertex
  std::vector<uint32_t> tileBegin;  // tile -> first position in idx (tiles are contiguous in idx)
  uint32_t next = 0;

  void split(uint32_t lo, uint32_t hi, uint32_t parts) {
    if (parts <= 1 || hi - lo <= 1) {
      const uint32_t t = next++;
      tileBegin.push_back(lo);
      std::sort(idx.begin() + lo, idx.begin() + hi);   // caller order inside a tile
      for (uint32_t i = lo; i < hi; ++i) tileOf[idx[i]] = t;
      return;
    }
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = lo; i < hi; ++i)
      for (int a = 0; a < 3; ++a) {
        const float c = x[3 * (size_t)idx[i] + a];
        mn[a] = std::min(mn[a], c);
        mx[a] = std::max(mx[a], c);
      }
    int ax = 0;
    for (int a = 1; a < 3; ++a)
      if (mx[a] - mn[a] > mx[ax] - mn[ax]) ax = a;
    const uint32_t pl = parts / 2, pr = parts - pl;
    const uint32_t mid = lo + (uint32_t)(((uint64_t)(hi - lo) * pl) / parts);
    auto cmp = [&](uint32_t a, uint32_t b) {
      const float ca = x[3 * (size_t)a + ax], cb = x[3 * (size_t)b + ax];
      return ca < cb || (ca == cb && a < b);   // total order -> the split is unique
    };
    std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi, cmp);
    split(lo, mid, pl);
    split(mid, hi, pr);
  }
};

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
// avoid the previous phase's cuts.  tileOfSlot (size V, all NONE on entry) receives the result
// and is reset by the caller afterwards.
void grow_tiles(const CSet& cs, const std::vector<uint32_t>& res, uint32_t V, uint32_t cap, uint32_t target,
                std::vector<uint32_t>& compactOf /* size V scratch, all NONE */, GrownTiles& out,
                std::vector<uint32_t>& tileOfSlot) {
  // compact vertex set U and CSR vertex -> residual constraints
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

  std::vector<uint32_t> tileOf(U, NONE);   // compact vertex -> tile
  std::vector<uint32_t> tileSize;
  std::vector<uint32_t> comp(U, NONE), dist(U), queue;
  queue.reserve(U);

  // pending = vertices not yet in a tile; processed as connected components, repeatedly
  std::vector<uint8_t> pending(U, 1);
  uint32_t binTile = NONE;   // tile currently collecting small components
  for (int round = 0; round < 64; ++round) {
    bool any = false;
    std::fill(comp.begin(), comp.end(), NONE);
    for (uint32_t s0 = 0; s0 < U; ++s0) {
      if (!pending[s0] || comp[s0] != NONE) continue;
      any = true;
      // BFS the component of s0 among pending vertices
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
      // large component: k seeds by farthest-point sampling over hop distance, then
      // capacity-limited multi-source BFS (graph Voronoi)
      const uint32_t k = (n + target - 1) / target;
      for (uint32_t u : members) dist[u] = NONE;
      std::vector<uint32_t> seeds;
      // first seed: the vertex farthest from an arbitrary start (a peripheral vertex)
      uint32_t far = members.back();   // last vertex reached by the component BFS
      for (uint32_t si = 0; si < k; ++si) {
        seeds.push_back(far);
        // relax distances from the new seed (bounded: only where it improves)
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
  // any vertex still pending after the rounds (pathological) becomes its own bin tiles
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

// ------------------------------------------------------------------ per-type phase construction

struct TileBuild {
  bool contiguous = false;
  uint32_t rangeBegin = 0, rangeCount = 0;   // contiguous tiles
  std::vector<uint32_t> verts;               // gathered tiles: slots, ascending
  std::vector<uint32_t> cons;                // constraint ids of this type, sorted by (colour, id)
  std::vector<uint32_t> colour;              // parallel to cons
  uint32_t nColours = 0;
};

struct TypeSchedule {
  std::vector<std::vector<TileBuild>> phases;
  uint32_t colourSum = 0;   // sum over phases of the largest local colour count
};

// colour the constraints of one tile locally and sort them by (colour, id)
void colour_tile(const CSet& cs, TileBuild& tb, const std::vector<uint32_t>& localOf /* slot -> local */,
                 std::vector<uint32_t>& scratchIds) {
  const uint32_t n = (uint32_t)tb.cons.size();
  std::sort(tb.cons.begin(), tb.cons.end());
  scratchIds.resize((size_t)n * cs.arity);
  const uint32_t nLocal = tb.contiguous ? tb.rangeCount : (uint32_t)tb.verts.size();
  for (uint32_t i = 0; i < n; ++i)
    for (uint32_t j = 0; j < cs.arity; ++j) scratchIds[(size_t)i * cs.arity + j] = localOf[cs.at(tb.cons[i])[j]];
  std::vector<uint32_t> col;
  tb.nColours = greedy_colour(scratchIds.data(), n, cs.arity, nLocal, col);
  std::vector<uint32_t> order(n);
  std::iota(order.begin(), order.end(), 0u);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return col[a] < col[b]; });
  std::vector<uint32_t> c2(n), k2(n);
  for (uint32_t i = 0; i < n; ++i) { c2[i] = tb.cons[order[i]]; k2[i] = col[order[i]]; }
  tb.cons.swap(c2);
  tb.colour.swap(k2);
}

void build_type_schedule(const CSet& cs, uint32_t V, const std::vector<uint32_t>& tile0Begin /* K1+1 */,
                         const std::vector<uint32_t>& tile0OfSlot, uint32_t cap, uint32_t target,
                         uint32_t maxPhases, TypeSchedule& ts) {
  const uint32_t K1 = (uint32_t)tile0Begin.size() - 1;
  std::vector<uint32_t> localOf(V, NONE), compactOf(V, NONE), tileOfSlot(V, NONE), scratch;

  // phase 0: the RCB tiles (contiguous slot ranges); constraints interior to one tile
  std::vector<TileBuild> p0(K1);
  for (uint32_t t = 0; t < K1; ++t) {
    p0[t].contiguous = true;
    p0[t].rangeBegin = tile0Begin[t];
    p0[t].rangeCount = tile0Begin[t + 1] - tile0Begin[t];
  }
  std::vector<uint32_t> res;
  for (uint32_t k = 0; k < cs.n; ++k) {
    const uint32_t* id = cs.at(k);
    const uint32_t t = tile0OfSlot[id[0]];
    bool same = true;
    for (uint32_t j = 1; j < cs.arity; ++j) same &= tile0OfSlot[id[j]] == t;
    if (same) p0[t].cons.push_back(k); else res.push_back(k);
  }
  for (uint32_t s = 0; s < V; ++s) localOf[s] = s - tile0Begin[tile0OfSlot[s]];
  uint32_t mx = 0;
  for (auto& tb : p0) { colour_tile(cs, tb, localOf, scratch); mx = std::max(mx, tb.nColours); }
  ts.colourSum += mx;
  ts.phases.push_back(std::move(p0));

  // later phases: re-partition the residual
  while (!res.empty()) {
    GrownTiles g;
    const bool last = maxPhases && ts.phases.size() + 1 >= maxPhases;
    (void)last;
    grow_tiles(cs, res, V, cap, target, compactOf, g, tileOfSlot);
    std::vector<TileBuild> ph(g.verts.size());
    std::vector<uint32_t> next;
    for (uint32_t k : res) {
      const uint32_t* id = cs.at(k);
      const uint32_t t = tileOfSlot[id[0]];
      bool same = t != NONE;
      for (uint32_t j = 1; j < cs.arity && same; ++j) same = tileOfSlot[id[j]] == t;
      if (same) ph[t].cons.push_back(k); else next.push_back(k);
    }
    if (next.size() == res.size()) {
      // no progress (cannot happen with hop-distance Voronoi on a connected residual, but stay
      // safe): peel off a vertex-disjoint set of single-constraint tiles
      ph.clear();
      g.verts.clear();
      next.clear();
      for (uint32_t s = 0; s < V; ++s) tileOfSlot[s] = NONE;
      for (uint32_t k : res) {
        const uint32_t* id = cs.at(k);
        bool free = true;
        for (uint32_t j = 0; j < cs.arity; ++j) free &= tileOfSlot[id[j]] == NONE;
        if (!free) { next.push_back(k); continue; }
        std::vector<uint32_t> vs(id, id + cs.arity);
        std::sort(vs.begin(), vs.end());
        vs.erase(std::unique(vs.begin(), vs.end()), vs.end());
        for (uint32_t s : vs) tileOfSlot[s] = (uint32_t)g.verts.size();
        g.verts.push_back(vs);
        ph.emplace_back();
        ph.back().cons.push_back(k);
      }
    }
    // drop tiles that received no constraint; fix local numbering; colour
    std::vector<TileBuild> kept;
    mx = 0;
    for (uint32_t t = 0; t < ph.size(); ++t) {
      if (ph[t].cons.empty()) continue;
      TileBuild& tb = ph[t];
      // keep only vertices that a constraint of this tile touches
      std::vector<uint32_t> used;
      for (uint32_t k : tb.cons)
        for (uint32_t j = 0; j < cs.arity; ++j) used.push_back(cs.at(k)[j]);
      std::sort(used.begin(), used.end());
      used.erase(std::unique(used.begin(), used.end()), used.end());
      tb.verts.swap(used);
      for (uint32_t i = 0; i < tb.verts.size(); ++i) localOf[tb.verts[i]] = i;
      colour_tile(cs, tb, localOf, scratch);
      mx = std::max(mx, tb.nColours);
      kept.push_back(std::move(tb));
    }
    for (auto& vs : g.verts)
      for (uint32_t s : vs) tileOfSlot[s] = NONE;
    ts.colourSum += mx;
    ts.phases.push_back(std::move(kept));
    res.swap(next);
  }
}

}  // namespace

bool build_tile_plan(const MeshView& m, const pbd_options& opts, uint32_t nSMs, uint32_t smemVertexLimit,
                     Plan& plan, std::string& err) {
  const double t0 = now_ms();
  if (nSMs == 0) nSMs = 148;
  if (smemVertexLimit < 64) { err = "shared memory too small for a vertex tile"; return false; }
  if (smemVertexLimit > 65535) smemVertexLimit = 65535;   // tile-local indices are 16 bit
  plan.V = m.V; plan.E = m.E; plan.T = m.T;
  plan.backend = PBD_BACKEND_TILE;
  plan.orderMode = opts.order_mode;
  if (opts.order_mode != PBD_ORDER_STRICT) { err = "interleaved order is not implemented yet"; return false; }
  const uint32_t blockThreads = opts.block_threads ? opts.block_threads : 512;
  if (blockThreads % 32 || blockThreads > 512) { err = "block_threads must be a multiple of 32, <= 512"; return false; }

  // ---- phase-0 tile count: one tile per SM (a whole number of waves) unless the tiles would
  // not fit in shared memory or the body is small enough for fewer tiles
  uint32_t tv = opts.tile_vertices;
  uint32_t K1;
  if (tv) {
    tv = std::min(tv, smemVertexLimit);
    K1 = std::max(1u, (m.V + tv - 1) / tv);
  } else {
    const uint32_t minTile = 1024;   // below this a tile is all interface: use fewer SMs instead
    K1 = std::max(1u, std::min(nSMs, m.V / minTile));
    if ((uint64_t)K1 * smemVertexLimit < m.V) {
      K1 = (m.V + smemVertexLimit - 1) / smemVertexLimit;
      K1 = ((K1 + nSMs - 1) / nSMs) * nSMs;
    }
    tv = (m.V + K1 - 1) / K1;
  }
  if (m.V == 0) K1 = 1;

  // ---- RCB, slot numbering (tile-major, caller order inside a tile)
  Rcb rcb;
  rcb.x = m.x0;
  rcb.idx.resize(m.V);
  std::iota(rcb.idx.begin(), rcb.idx.end(), 0u);
  rcb.tileOf.assign(m.V, 0);
  if (m.V) rcb.split(0, m.V, K1); else rcb.tileBegin.push_back(0);
  K1 = (uint32_t)rcb.tileBegin.size();
  std::vector<uint32_t> tile0Begin = rcb.tileBegin;
  tile0Begin.push_back(m.V);
  plan.slotToVertex = rcb.idx;
  plan.vertexToSlot.resize(m.V);
  for (uint32_t s = 0; s < m.V; ++s) plan.vertexToSlot[rcb.idx[s]] = s;
  std::vector<uint32_t> tile0OfSlot(m.V);
  uint32_t cap0 = 0;
  for (uint32_t t = 0; t < K1; ++t) {
    for (uint32_t s = tile0Begin[t]; s < tile0Begin[t + 1]; ++s) tile0OfSlot[s] = t;
    cap0 = std::max(cap0, tile0Begin[t + 1] - tile0Begin[t]);
  }
  if (cap0 > smemVertexLimit) { err = "tile does not fit in shared memory"; return false; }

  // constraints in slot numbering
  std::vector<uint32_t> eSlots((size_t)m.E * 2), tSlots((size_t)m.T * 4);
  for (size_t i = 0; i < eSlots.size(); ++i) eSlots[i] = plan.vertexToSlot[m.edges[i]];
  for (size_t i = 0; i < tSlots.size(); ++i) tSlots[i] = plan.vertexToSlot[m.tets[i]];

  // later phases may use larger tiles: their sweeps are latency-bound, so fewer, larger tiles
  // cost nothing and leave fewer straddling constraints
  const uint32_t cap = std::min(smemVertexLimit, std::max(2 * tv, 2048u));
  const uint32_t target = std::max(64u, (uint32_t)(cap * 0.75));

  TypeSchedule sched[2];
  CSet sets[2] = {{eSlots.data(), m.E, 2}, {tSlots.data(), m.T, 4}};
  for (int ty = 0; ty < 2; ++ty) build_type_schedule(sets[ty], m.V, tile0Begin, tile0OfSlot, cap, target, opts.max_phases, sched[ty]);

  // ---- flatten
  plan.tile0Begin = tile0Begin;
  plan.edgeOrder.clear(); plan.tetOrder.clear();
  plan.edgeOrder.reserve(m.E); plan.tetOrder.reserve(m.T);
  plan.edgeLocal.clear(); plan.tetLocal.clear();
  plan.edgeLocal.reserve((size_t)m.E * 2); plan.tetLocal.reserve((size_t)m.T * 4);
  plan.edgePhase.assign(m.E, 0); plan.edgeTile.assign(m.E, 0); plan.edgeColor.assign(m.E, 0);
  plan.tetPhase.assign(m.T, 0); plan.tetTile.assign(m.T, 0); plan.tetColor.assign(m.T, 0);
  plan.tileVertexCapacity = cap0;
  std::vector<uint32_t> localOf(m.V, NONE);
  for (int ty = 0; ty < 2; ++ty) {
    const CSet& cs = sets[ty];
    std::vector<uint32_t>& order = ty ? plan.tetOrder : plan.edgeOrder;
    std::vector<uint16_t>& local = ty ? plan.tetLocal : plan.edgeLocal;
    std::vector<uint32_t>& cPhase = ty ? plan.tetPhase : plan.edgePhase;
    std::vector<uint32_t>& cTile = ty ? plan.tetTile : plan.edgeTile;
    std::vector<uint32_t>& cCol = ty ? plan.tetColor : plan.edgeColor;
    for (auto& ph : sched[ty].phases) {
      if (cs.n == 0) break;   // a type without constraints contributes no phase
      Phase P;
      P.tileBegin = (uint32_t)plan.tiles.size();
      P.isTet = (uint32_t)ty;
      for (auto& tb : ph) {
        Tile tl;
        tl.isTet = (uint32_t)ty;
        tl.contiguous = tb.contiguous ? 1u : 0u;
        if (tb.contiguous) {
          tl.vertBegin = tb.rangeBegin;
          tl.vertCount = tb.rangeCount;
          for (uint32_t i = 0; i < tb.rangeCount; ++i) localOf[tb.rangeBegin + i] = i;
        } else {
          tl.vertBegin = (uint32_t)plan.tileVerts.size();
          tl.vertCount = (uint32_t)tb.verts.size();
          for (uint32_t i = 0; i < tb.verts.size(); ++i) { localOf[tb.verts[i]] = i; plan.tileVerts.push_back(tb.verts[i]); }
        }
        plan.tileVertexCapacity = std::max(plan.tileVertexCapacity, tl.vertCount);
        tl.groupBegin = (uint32_t)plan.groups.size();
        // one group per colour, split so that no group exceeds the block size
        size_t i = 0;
        while (i < tb.cons.size()) {
          size_t j = i;
          while (j < tb.cons.size() && tb.colour[j] == tb.colour[i]) ++j;
          const uint32_t lim = ty ? blockThreads / 4 : blockThreads;   // a tet is swept by 4 lanes
          const uint32_t n = (uint32_t)(j - i), parts = (n + lim - 1) / lim;
          for (uint32_t q = 0; q < parts; ++q) {
            Group g;
            g.begin = (uint32_t)order.size() + (uint32_t)(((uint64_t)n * q) / parts);
            g.count = (uint32_t)(((uint64_t)n * (q + 1)) / parts - ((uint64_t)n * q) / parts);
            plan.groups.push_back(g);
          }
          for (size_t k = i; k < j; ++k) {
            const uint32_t c = tb.cons[k];
            cPhase[c] = (uint32_t)plan.phases.size();
            cTile[c] = (uint32_t)plan.tiles.size();
            cCol[c] = tb.colour[k];
            for (uint32_t a = 0; a < cs.arity; ++a) local.push_back((uint16_t)localOf[cs.at(c)[a]]);
          }
          for (size_t k = i; k < j; ++k) order.push_back(tb.cons[k]);
          i = j;
        }
        tl.groupCount = (uint32_t)plan.groups.size() - tl.groupBegin;
        plan.tiles.push_back(tl);
      }
      P.tileCount = (uint32_t)plan.tiles.size() - P.tileBegin;
      plan.phases.push_back(P);
    }
  }
  plan.edgePhases = m.E ? (uint32_t)sched[0].phases.size() : 0;
  plan.tetPhases = m.T ? (uint32_t)sched[1].phases.size() : 0;
  plan.edgeColorSum = sched[0].colourSum;
  plan.tetColorSum = sched[1].colourSum;
  plan.blockThreads = blockThreads;
  plan.planMs = now_ms() - t0;
  return true;
}

}  // namespace pbd

// pbd_placement.cpp -- where a tile's vertices sit in shared memory, and which eight constraints
// of a colour group share a quarter-warp (host, pure C++17, deterministic).
//
// The colour sweeps (pbd_sweep.cuh) gather and scatter 16-byte vertices; the hardware serves a
// 128-bit shared-memory access one quarter-warp (8 lanes) at a time, in ONE wavefront when the 8
// addresses fall into 8 different 16-byte bank groups, i.e. when the 8 tile-local vertex indices
// differ modulo 8 -- otherwise in as many wavefronts as the fullest bank group holds.  With
// vertices in partition order and a greedy row packing the sweeps measured 1.4 (edges) / 1.67
// (tets) wavefronts per quarter-warp access, 35 % of all shared-memory wavefronts of the frame
// kernel (ncu l1tex__data_bank_conflicts_pipe_lsu_mem_shared), on the pipe that bounds it.
//
// Two freedoms cost nothing at run time and change no result (a colour group's constraints share no
// vertex; a tile's vertices may sit anywhere in its shared array):
//   * the CLASS (index mod 8) of every vertex of the tile  -> a permutation of the tile's local indices
//   * the ROW (quarter-warp) of every constraint in its group -> the order inside the group
// optimise_placement() chooses both:
//   A. classes by local search (class swaps) on the per-(group, role) class histograms: no class
//      above ceil(n/8) -- for edges that is also sufficient (B);
//   B. rows: edges by bipartite edge colouring (Koenig: a bipartite multigraph of maximum degree D
//      splits into D matchings; nodes = classes of the a / b endpoints, a matching = a conflict-free
//      row), tets by randomised greedy; both followed by pairwise row exchanges;
// (Class swaps on the exact wavefront count of the finished rows were tried as a third stage and never found an
// improving move: a vertex sits in ~9 rows, moving it repairs one and breaks the others.  What limits the tets is
// B: with four roles a collision-free row is a 4-dimensional matching; ~1.5 wavefronts per access remain.)
// Replaces nothing in the reference (CProgram sweeps sequentially): this is schedule layout only.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <cstring>
#include <numeric>
#include <vector>

#include "pbd_plan.h"

namespace pbd {

namespace {

constexpr uint32_t NONE = 0xffffffffu;

struct Lcg {
  uint32_t s;
  uint32_t next() { s = s * 1664525u + 1013904223u; return s >> 8; }
  uint32_t below(uint32_t n) { return n ? next() % n : 0u; }
};

inline uint8_t max8(const uint8_t* c) {
  uint8_t m = c[0];
  for (int k = 1; k < 8; ++k) m = std::max(m, c[k]);
  return m;
}

}  // namespace

// Randomised-greedy packing of n constraints with given residues (one byte per role, 4 per
// constraint) into rows of 8: fill one row at a time with constraints whose residues are still free
// in every role (several pseudo-random scan orders, keep the fullest), complete a short row with
// the constraints that add the fewest collisions.  out = the constraints' indices in row order.
void pack_rows_greedy(const uint8_t* res, uint32_t ar, uint32_t n, uint32_t seed, int tries, std::vector<uint32_t>& out) {
  out.clear();
  out.reserve(n);
  std::vector<uint32_t> alive(n);
  std::iota(alive.begin(), alive.end(), 0u);
  Lcg rng{seed};
  std::vector<uint32_t> row, bestRow;
  std::vector<uint8_t> inRow(n, 0);
  while (!alive.empty()) {
    const uint32_t m = (uint32_t)alive.size();
    if (m <= 8) {
      for (uint32_t a : alive) out.push_back(a);
      break;
    }
    bestRow.clear();
    for (int tr = 0; tr < tries && bestRow.size() < 8; ++tr) {
      const uint32_t start = rng.below(m);
      row.clear();
      uint8_t used[4] = {0, 0, 0, 0};
      for (uint32_t q = 0; q < m && row.size() < 8; ++q) {
        const uint32_t a = alive[(start + q) % m];
        const uint8_t* rs = &res[(size_t)a * 4];
        bool ok = true;
        for (uint32_t r = 0; r < ar && ok; ++r) ok = !(used[r] >> rs[r] & 1u);
        if (!ok) continue;
        for (uint32_t r = 0; r < ar; ++r) used[r] |= (uint8_t)(1u << rs[r]);
        row.push_back(a);
      }
      if (row.size() > bestRow.size()) bestRow = row;
    }
    if (bestRow.size() < 8) {
      uint8_t cnt[4][8] = {};
      for (uint32_t a : bestRow) { inRow[a] = 1; for (uint32_t r = 0; r < ar; ++r) cnt[r][res[(size_t)a * 4 + r]]++; }
      while (bestRow.size() < 8) {
        uint32_t pick = NONE, pickCost = NONE;
        for (uint32_t a : alive) {
          if (inRow[a]) continue;
          uint32_t cost = 0;
          for (uint32_t r = 0; r < ar; ++r) cost += cnt[r][res[(size_t)a * 4 + r]];
          if (cost < pickCost) { pickCost = cost; pick = a; }
        }
        if (pick == NONE) break;
        inRow[pick] = 1;
        for (uint32_t r = 0; r < ar; ++r) cnt[r][res[(size_t)pick * 4 + r]]++;
        bestRow.push_back(pick);
      }
    }
    for (uint32_t a : bestRow) { out.push_back(a); inRow[a] = 2; }
    alive.erase(std::remove_if(alive.begin(), alive.end(), [&](uint32_t a) { return inRow[a] == 2; }), alive.end());
  }
}

namespace {

// Proper arc colouring of a bipartite multigraph (nodes 0..7 on either side, arc e = (ax[e], ay[e])) with D = its
// maximum degree colours (Koenig): colOf[e] in [0, D).  Arcs are coloured one by one; when the colour free at x is
// taken at y, the two-coloured alternating path starting at y is flipped (it cannot end at x).  Returns D.
uint32_t bipartite_arc_colouring(const uint8_t* ax, const uint8_t* ay, uint32_t stride, uint32_t n, std::vector<uint32_t>& colOf) {
  uint32_t degA[8] = {}, degB[8] = {};
  for (uint32_t i = 0; i < n; ++i) { degA[ax[(size_t)i * stride]]++; degB[ay[(size_t)i * stride]]++; }
  uint32_t D = 1;
  for (int k = 0; k < 8; ++k) D = std::max({D, degA[k], degB[k]});
  // atA[x * D + c] = arc of colour c at left node x (NONE: colour free there)
  std::vector<uint32_t> atA((size_t)8 * D, NONE), atB((size_t)8 * D, NONE), path;
  colOf.assign(n, NONE);
  auto X = [&](uint32_t e) { return (uint32_t)ax[(size_t)e * stride]; };
  auto Y = [&](uint32_t e) { return (uint32_t)ay[(size_t)e * stride]; };
  for (uint32_t e = 0; e < n; ++e) {
    const uint32_t x = X(e), y = Y(e);
    uint32_t ca = 0, cb = 0;
    while (atA[x * D + ca] != NONE) ++ca;   // a free colour exists at both ends: fewer than D arcs are coloured there
    while (atB[y * D + cb] != NONE) ++cb;
    if (atB[y * D + ca] != NONE) {
      path.clear();
      uint32_t node = y, want = ca;
      bool sideB = true;
      for (;;) {
        const uint32_t f = sideB ? atB[node * D + want] : atA[node * D + want];
        if (f == NONE) break;
        path.push_back(f);
        node = sideB ? X(f) : Y(f);
        sideB = !sideB;
        want = want == ca ? cb : ca;
      }
      for (uint32_t f : path) { atA[X(f) * D + colOf[f]] = NONE; atB[Y(f) * D + colOf[f]] = NONE; }
      for (uint32_t f : path) {
        colOf[f] = colOf[f] == ca ? cb : ca;
        atA[X(f) * D + colOf[f]] = f;
        atB[Y(f) * D + colOf[f]] = f;
      }
    }
    colOf[e] = ca;
    atA[x * D + ca] = e;
    atB[y * D + ca] = e;
  }
  return D;
}

// Edges: rows by bipartite edge colouring.  Node sets = the 8 classes of the a-endpoints and the 8
// classes of the b-endpoints; every edge constraint is an arc (class a, class b); every colour class of a
// proper arc colouring is a matching = a row without a collision in either role.  The D matchings are then
// compacted into ceil(n/8) rows, all full but the last (the kernel addresses a group densely): arcs of the
// smallest matchings fill the holes of the larger ones where they add the fewest collisions.
void pack_rows_koenig(const uint8_t* res, uint32_t n, std::vector<uint32_t>& out) {
  out.clear();
  if (n == 0) return;
  std::vector<uint32_t> colOf;
  const uint32_t D = bipartite_arc_colouring(res, res + 1, 4, n, colOf);
  std::vector<std::vector<uint32_t>> rows(D);
  for (uint32_t e = 0; e < n; ++e) rows[colOf[e]].push_back(e);
  std::stable_sort(rows.begin(), rows.end(), [](const std::vector<uint32_t>& p, const std::vector<uint32_t>& q) { return p.size() > q.size(); });
  const uint32_t R = (n + 7) / 8;
  std::vector<uint32_t> pool;
  for (uint32_t r = R; r < D; ++r) pool.insert(pool.end(), rows[r].begin(), rows[r].end());
  rows.resize(R);
  auto added = [&](const std::vector<uint32_t>& row, uint32_t e) {
    uint32_t c = 0;
    for (uint32_t f : row) c += (res[4 * (size_t)f] == res[4 * (size_t)e]) + (res[4 * (size_t)f + 1] == res[4 * (size_t)e + 1]);
    return c;
  };
  for (uint32_t r = 0; r + 1 < R || (r < R && !pool.empty()); ++r) {
    const uint32_t want = r + 1 < R ? 8u : (n - 8u * (R - 1));
    while (rows[r].size() < want) {
      std::vector<uint32_t>& src = !pool.empty() ? pool : rows[R - 1];
      if (&src == &rows[r] || src.empty()) break;
      uint32_t best = 0, bestCost = NONE;
      for (uint32_t i = 0; i < src.size(); ++i) {
        const uint32_t c = added(rows[r], src[i]);
        if (c < bestCost) { bestCost = c; best = i; }
      }
      rows[r].push_back(src[best]);
      src.erase(src.begin() + best);
    }
  }
  for (auto& row : rows) out.insert(out.end(), row.begin(), row.end());
}

// Tets whose four vertices may take the four roles in ANY order (fast arithmetic: a permutation of a tet's vertices
// changes its signed volume by the permutation's sign only; the caller negates the rest volume of odd ones).  A row of
// eight tets is then collision-free in all four roles as soon as no class occurs more than four times among its 32
// vertices: the multigraph tet -- class has maximum degree 4 and its arc colouring with 4 colours hands every tet one
// vertex per role and every role eight different classes.  So rows are chosen for flat class histograms (greedy, then
// pairwise exchanges) -- a one-dimensional condition instead of the four-dimensional matching of fixed roles.
// cls4: the classes of every tet's vertices; out: tets in row order; roleOf[4 t + i] = role of tet t's i-th vertex.
void pack_rows_relabel(const uint8_t* cls4, uint32_t n, std::vector<uint32_t>& out, std::vector<uint8_t>& roleOf) {
  out.clear();
  roleOf.assign((size_t)n * 4, 0);
  for (uint32_t t = 0; t < n; ++t) for (uint8_t i = 0; i < 4; ++i) roleOf[(size_t)t * 4 + i] = i;
  if (n == 0) return;
  const uint32_t R = (n + 7) / 8;
  std::vector<uint32_t> rowOfT(n, NONE), fill(R, 0);
  std::vector<uint8_t> hist((size_t)R * 8, 0);
  auto over_if_added = [&](uint32_t r, uint32_t t) {
    uint8_t add[8] = {};
    for (int i = 0; i < 4; ++i) add[cls4[(size_t)t * 4 + i]]++;
    int o = 0;
    for (int k = 0; k < 8; ++k) {
      const int before = hist[(size_t)r * 8 + k], after = before + add[k];
      o += std::max(0, after - 4) - std::max(0, before - 4);
    }
    return o;
  };
  auto put = [&](uint32_t r, uint32_t t, int sign) {
    for (int i = 0; i < 4; ++i) hist[(size_t)r * 8 + cls4[(size_t)t * 4 + i]] += (uint8_t)sign;
  };
  // greedy: every row takes, eight times, the free tet that overfills its histogram least (ties: the lowest index)
  std::vector<uint32_t> freeT(n);
  std::iota(freeT.begin(), freeT.end(), 0u);
  for (uint32_t r = 0; r < R; ++r) {
    const uint32_t want = r + 1 < R ? 8u : n - 8u * (R - 1);
    while (fill[r] < want) {
      uint32_t best = 0;
      int bestO = 1 << 30;
      for (uint32_t q = 0; q < freeT.size() && bestO > 0; ++q) {
        const int o = over_if_added(r, freeT[q]);
        if (o < bestO) { bestO = o; best = q; }
      }
      const uint32_t t = freeT[best];
      freeT.erase(freeT.begin() + best);
      rowOfT[t] = r; put(r, t, 1); fill[r]++;
    }
  }
  // exchanges between rows while they lower the total overfill
  auto row_over = [&](uint32_t r) { int o = 0; for (int k = 0; k < 8; ++k) o += std::max(0, (int)hist[(size_t)r * 8 + k] - 4); return o; };
  for (int pass = 0; pass < 3; ++pass) {
    bool any = false;
    for (uint32_t t = 0; t < n; ++t) {
      const uint32_t r = rowOfT[t];
      if (row_over(r) == 0) continue;
      for (uint32_t u = 0; u < n; ++u) {
        const uint32_t q = rowOfT[u];
        if (q == r) continue;
        const int before = row_over(r) + row_over(q);
        put(r, t, -1); put(q, u, -1); put(r, u, 1); put(q, t, 1);
        if (row_over(r) + row_over(q) < before) { rowOfT[t] = q; rowOfT[u] = r; any = true; break; }
        put(r, u, -1); put(q, t, -1); put(r, t, 1); put(q, u, 1);
      }
    }
    if (!any) break;
  }
  // roles inside every row: arc colouring of tet -- class; colours beyond the four roles (a class more than four times
  // in the row) fall back to a role the tet has not used yet
  std::vector<std::vector<uint32_t>> rows(R);
  for (uint32_t t = 0; t < n; ++t) rows[rowOfT[t]].push_back(t);
  std::vector<uint8_t> ax, ay;
  std::vector<uint32_t> colOf;
  for (uint32_t r = 0; r < R; ++r) {
    const std::vector<uint32_t>& row = rows[r];
    ax.clear(); ay.clear();
    for (uint32_t q = 0; q < row.size(); ++q)
      for (int i = 0; i < 4; ++i) { ax.push_back((uint8_t)q); ay.push_back(cls4[(size_t)row[q] * 4 + i]); }
    bipartite_arc_colouring(ax.data(), ay.data(), 1, (uint32_t)ax.size(), colOf);
    for (uint32_t q = 0; q < row.size(); ++q) {
      uint32_t used = 0;
      for (int i = 0; i < 4; ++i) if (colOf[4 * q + i] < 4) used |= 1u << colOf[4 * q + i];
      for (int i = 0; i < 4; ++i) {
        uint32_t c = colOf[4 * q + i];
        if (c >= 4) { c = 0; while (used >> c & 1u) ++c; used |= 1u << c; }
        roleOf[(size_t)row[q] * 4 + i] = (uint8_t)c;
      }
      out.push_back(row[q]);
    }
  }
}

struct Optimiser {
  uint32_t nLocal, nCons;
  uint32_t block = 0;   // class swaps stay inside aligned blocks of this many indices (a multiple of 8; 0: anywhere)
  const PlaceGroup* groups;
  uint32_t nGroups;
  uint32_t* loc;        // 4 per constraint, tile-local vertex indices (NONE beyond the arity); reordered in place
  uint32_t* payload;    // 1 per constraint; reordered with loc
  uint8_t* perm = nullptr;   // 1 per constraint or null: relabelling allowed; out: bits 2r..2r+1 = which of the constraint's
                             // ORIGINAL vertices (0..3) sits in role r now
  std::vector<uint8_t> cls;          // class (0..7) of every local vertex
  std::vector<uint32_t> incOff, inc; // vertex -> constraint * 4 + role
  std::vector<uint32_t> groupOf, rowOf;   // per constraint
  std::vector<uint32_t> rowBase;          // per group: its first row
  std::vector<uint32_t> groupOfRow;
  uint32_t nRows = 0;
  std::vector<uint8_t> rcnt;   // [(row * 4 + role) * 8 + class]
  std::vector<uint8_t> rmax;   // [row * 4 + role]
  std::vector<uint16_t> gcnt;  // [(group * 4 + role) * 8 + class]
  std::vector<uint16_t> gcap;  // per group: ceil(count / 8)
  Lcg rng{0x9e3779b9u};

  bool merged(uint32_t g) const { return perm != nullptr && groups[g].arity == 4; }

  void build_incidence() {
    incOff.assign((size_t)nLocal + 1, 0);
    for (uint32_t c = 0; c < nCons; ++c)
      for (uint32_t r = 0; r < 4; ++r)
        if (loc[4 * (size_t)c + r] != NONE) incOff[loc[4 * (size_t)c + r] + 1]++;
    for (uint32_t v = 0; v < nLocal; ++v) incOff[v + 1] += incOff[v];
    inc.resize(incOff[nLocal]);
    std::vector<uint32_t> cur(incOff.begin(), incOff.end() - 1);
    for (uint32_t c = 0; c < nCons; ++c)
      for (uint32_t r = 0; r < 4; ++r)
        if (loc[4 * (size_t)c + r] != NONE) inc[cur[loc[4 * (size_t)c + r]]++] = 4 * c + r;
  }

  // ---- A: class histograms per (group, role)
  // cost = 64 * (members above the group's cap) + sum of squared counts (the second term keeps the
  // histograms flat below the cap too, which is what the tet rows need)
  int64_t move_hist(uint32_t v, uint8_t from, uint8_t to) {
    int64_t d = 0;
    for (uint32_t a = incOff[v]; a < incOff[v + 1]; ++a) {
      const uint32_t c = inc[a] >> 2, g = groupOf[c], r = merged(g) ? 0u : (inc[a] & 3u);
      uint16_t* h = &gcnt[((size_t)g * 4 + r) * 8];
      const uint16_t cap = gcap[g];
      d -= 2 * (int64_t)h[from] - 1; if (h[from] > cap) d -= 64; h[from]--;
      d += 2 * (int64_t)h[to] + 1; if (h[to] >= cap) d += 64; h[to]++;
    }
    return d;
  }
  int64_t swap_hist(uint32_t u, uint32_t v) {
    const uint8_t cu = cls[u], cv = cls[v];
    const int64_t d = move_hist(u, cu, cv) + move_hist(v, cv, cu);
    cls[u] = cv; cls[v] = cu;
    return d;
  }
  void balance_classes(int passes, uint32_t candidates) {
    gcnt.assign((size_t)nGroups * 32, 0);
    gcap.resize(nGroups);
    // (relabelled tets: the roles are dealt out per row afterwards, so a group's four roles share ONE histogram, four times as deep)
    for (uint32_t g = 0; g < nGroups; ++g) gcap[g] = (uint16_t)((merged(g) ? 4u : 1u) * ((groups[g].count + 7) / 8));
    for (uint32_t c = 0; c < nCons; ++c)
      for (uint32_t r = 0; r < 4; ++r)
        if (loc[4 * (size_t)c + r] != NONE) gcnt[((size_t)groupOf[c] * 4 + (merged(groupOf[c]) ? 0u : r)) * 8 + cls[loc[4 * (size_t)c + r]]]++;
    std::vector<uint32_t> hot;
    for (int pass = 0; pass < passes; ++pass) {
      hot.clear();
      for (uint32_t v = 0; v < nLocal; ++v)
        for (uint32_t a = incOff[v]; a < incOff[v + 1]; ++a) {
          const uint32_t c = inc[a] >> 2, g = groupOf[c], r = merged(g) ? 0u : (inc[a] & 3u);
          if (gcnt[((size_t)g * 4 + r) * 8 + cls[v]] > gcap[g]) { hot.push_back(v); break; }
        }
      if (hot.empty()) break;
      uint32_t moves = 0;
      for (uint32_t i = (uint32_t)hot.size(); i > 1; --i) std::swap(hot[i - 1], hot[rng.below(i)]);
      for (uint32_t u : hot) {
        uint32_t bestV = NONE;
        int64_t bestD = 0;
        for (uint32_t t = 0; t < candidates; ++t) {
          // partners come from u's own block of `block` consecutive indices: the renumbering then moves a vertex by
          // less than one block, so a warp still loads and stores the same few global-memory lines as before
          const uint32_t b0 = block ? u / block * block : 0u;
          const uint32_t v = b0 + rng.below(std::min(block ? block : nLocal, nLocal - b0));
          if (cls[v] == cls[u]) continue;
          const int64_t d = swap_hist(u, v);
          if (d < bestD) { bestD = d; bestV = v; }
          swap_hist(u, v);   // back
        }
        if (bestV != NONE) { swap_hist(u, bestV); ++moves; }
      }
      if (!moves) break;
    }
  }

  // ---- B/C: exact wavefront bookkeeping of the rows
  void rebuild_rows() {
    rowBase.resize(nGroups);
    nRows = 0;
    for (uint32_t g = 0; g < nGroups; ++g) { rowBase[g] = nRows; nRows += (groups[g].count + 7) / 8; }
    groupOfRow.clear();
    for (uint32_t g = 0; g < nGroups; ++g) groupOfRow.insert(groupOfRow.end(), (groups[g].count + 7) / 8, g);
    rowOf.resize(nCons);
    rcnt.assign((size_t)nRows * 32, 0);
    rmax.assign((size_t)nRows * 4, 0);
    for (uint32_t g = 0; g < nGroups; ++g)
      for (uint32_t i = 0; i < groups[g].count; ++i) {
        const uint32_t c = groups[g].begin + i;
        rowOf[c] = rowBase[g] + i / 8;
        for (uint32_t r = 0; r < groups[g].arity; ++r) rcnt[((size_t)rowOf[c] * 4 + r) * 8 + cls[loc[4 * (size_t)c + r]]]++;
      }
    for (size_t e = 0; e < rmax.size(); ++e) rmax[e] = max8(&rcnt[e * 8]);
  }
  uint64_t total_wavefronts() const {
    uint64_t s = 0;
    for (uint8_t m : rmax) s += m;
    return s;
  }
  // change one entry (row, role): class `from` leaves, `to` arrives; returns the change of its wavefront count
  int retag(size_t e, uint8_t from, uint8_t to) {
    uint8_t* h = &rcnt[e * 8];
    h[from]--; h[to]++;
    const uint8_t m = max8(h);
    const int d = (int)m - (int)rmax[e];
    rmax[e] = m;
    return d;
  }
  // constraint c, counted in row rowC, goes to row rowD and d the other way (bookkeeping only; the caller swaps the
  // positions once it keeps the exchange).  Undo: the same call with the two rows exchanged.
  int exchange(uint32_t c, uint32_t rowC, uint32_t d, uint32_t rowD, uint32_t ar) {
    int delta = 0;
    for (uint32_t r = 0; r < ar; ++r) {
      const uint8_t kc = cls[loc[4 * (size_t)c + r]], kd = cls[loc[4 * (size_t)d + r]];
      if (kc == kd) continue;
      delta += retag((size_t)rowC * 4 + r, kc, kd);
      delta += retag((size_t)rowD * 4 + r, kd, kc);
    }
    return delta;
  }
  void swap_positions(uint32_t c, uint32_t d) {
    for (uint32_t r = 0; r < 4; ++r) std::swap(loc[4 * (size_t)c + r], loc[4 * (size_t)d + r]);
    std::swap(payload[c], payload[d]);
    if (perm) std::swap(perm[c], perm[d]);
    // rowOf stays with the POSITION; the incidence lists name positions: rebuild lazily (caller)
  }

  // rows of every group from scratch, for the current classes
  void pack_all() {
    std::vector<uint8_t> res, tmpPerm;
    std::vector<uint32_t> order, tmpLoc, tmpPay;
    for (uint32_t g = 0; g < nGroups; ++g) {
      const PlaceGroup& G = groups[g];
      if (G.count <= 1) continue;
      res.assign((size_t)G.count * 4, 0);
      for (uint32_t i = 0; i < G.count; ++i)
        for (uint32_t r = 0; r < G.arity; ++r) res[(size_t)i * 4 + r] = cls[loc[4 * (size_t)(G.begin + i) + r]];
      std::vector<uint8_t> roleOf;
      const bool relabel = perm != nullptr && G.arity == 4;
      if (G.arity == 2) pack_rows_koenig(res.data(), G.count, order);
      else if (relabel) pack_rows_relabel(res.data(), G.count, order, roleOf);
      else pack_rows_greedy(res.data(), G.arity, G.count, 0x2545f491u + g, 32, order);
      if (relabel) {
        // move every tet's vertices into their roles (on top of whatever relabelling it carries already)
        for (uint32_t i = 0; i < G.count; ++i) {
          uint32_t v[4];
          uint8_t src[4];
          const uint32_t c = G.begin + i;
          for (uint32_t j = 0; j < 4; ++j) { v[roleOf[(size_t)i * 4 + j]] = loc[4 * (size_t)c + j]; src[roleOf[(size_t)i * 4 + j]] = (uint8_t)(perm[c] >> (2 * j) & 3u); }
          uint8_t code = 0;
          for (uint32_t r = 0; r < 4; ++r) { loc[4 * (size_t)c + r] = v[r]; code |= (uint8_t)(src[r] << (2 * r)); }
          perm[c] = code;
        }
      }
      tmpLoc.assign(loc + 4 * (size_t)G.begin, loc + 4 * (size_t)(G.begin + G.count));
      tmpPay.assign(payload + G.begin, payload + G.begin + G.count);
      if (perm) tmpPerm.assign(perm + G.begin, perm + G.begin + G.count);
      for (uint32_t i = 0; i < G.count; ++i) {
        std::memcpy(loc + 4 * (size_t)(G.begin + i), &tmpLoc[4 * (size_t)order[i]], 16);
        if (perm) perm[G.begin + i] = tmpPerm[order[i]];
        payload[G.begin + i] = tmpPay[order[i]];
      }
    }
  }

  // pairwise exchanges between the rows of a group while they lower its wavefront count
  bool exchange_rows(int passes) {
    bool any = false;
    for (uint32_t g = 0; g < nGroups; ++g) {
      const PlaceGroup& G = groups[g];
      if (G.count <= 8) continue;
      for (int pass = 0; pass < passes; ++pass) {
        uint32_t moves = 0;
        for (uint32_t i = 0; i < G.count; ++i) {
          const uint32_t c = G.begin + i;
          bool hot = false;   // c collides with another constraint of its row in some role
          for (uint32_t r = 0; r < G.arity && !hot; ++r) {
            const size_t e = (size_t)rowOf[c] * 4 + r;
            hot = rmax[e] > 1 && rcnt[e * 8 + cls[loc[4 * (size_t)c + r]]] == rmax[e];
          }
          if (!hot) continue;
          uint32_t bestD = NONE;
          int best = 0;
          for (uint32_t j = 0; j < G.count; ++j) {
            const uint32_t d = G.begin + j;
            if (rowOf[d] == rowOf[c]) continue;
            const int delta = exchange(c, rowOf[c], d, rowOf[d], G.arity);
            if (delta < best) { best = delta; bestD = d; }
            exchange(c, rowOf[d], d, rowOf[c], G.arity);   // undo
          }
          if (bestD != NONE) {
            exchange(c, rowOf[c], bestD, rowOf[bestD], G.arity);
            swap_positions(c, bestD);   // rowOf belongs to the position
            ++moves;
            any = true;
          }
        }
        if (!moves) break;
      }
    }
    if (any) build_incidence();   // positions moved
    return any;
  }
};

}  // namespace

void optimise_placement(uint32_t nLocal, const PlaceGroup* groups, uint32_t nGroups, uint32_t* loc, uint32_t* payload,
                        uint32_t nCons, int effort, uint32_t block, std::vector<uint32_t>& newLocal, PlaceStats* stats, uint8_t* perm) {
  // PBD_PLACE_DUMP=<path>[:k]: write the k-th problem this process sees to <path> (input of tools/place_bench.cpp)
  if (const char* dump = getenv("PBD_PLACE_DUMP")) {
    static std::atomic<int> seen{0};
    const char* colon = strrchr(dump, ':');
    const int want = colon ? atoi(colon + 1) : 0;
    if (seen.fetch_add(1) == want) {
      const std::string path = colon ? std::string(dump, colon) : std::string(dump);
      if (FILE* f = fopen(path.c_str(), "wb")) {
        const uint32_t hdr[3] = {nLocal, nGroups, nCons};
        fwrite(hdr, 4, 3, f);
        fwrite(groups, sizeof(PlaceGroup), nGroups, f);
        fwrite(loc, 16, nCons, f);
        fclose(f);
      }
    }
  }
  newLocal.resize(nLocal);
  std::iota(newLocal.begin(), newLocal.end(), 0u);
  Optimiser o;
  o.nLocal = nLocal; o.nCons = nCons; o.block = block & ~7u; o.perm = perm;
  o.groups = groups; o.nGroups = nGroups; o.loc = loc; o.payload = payload;
  if (perm) std::memset(perm, 0xE4, nCons);   // identity: role r holds the constraint's r-th vertex
  o.cls.resize(nLocal);
  for (uint32_t v = 0; v < nLocal; ++v) o.cls[v] = (uint8_t)(v & 7u);
  o.groupOf.resize(nCons);
  for (uint32_t g = 0; g < nGroups; ++g)
    for (uint32_t i = 0; i < groups[g].count; ++i) o.groupOf[groups[g].begin + i] = g;
  if (nLocal >= 16 && nCons > 0 && effort > 0) {
    o.build_incidence();
    const bool dbg = getenv("PBD_PLACE_DEBUG") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
      if (!dbg) return;
      const auto t1 = std::chrono::steady_clock::now();
      uint64_t wf[2] = {0, 0};
      for (size_t e = 0; e < o.rmax.size(); ++e) wf[groups[o.groupOfRow[e / 4]].arity == 2 ? 0 : 1] += o.rmax[e];
      fprintf(stderr, "[place] %-14s %7.2f ms  edges %llu tets %llu\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count(),
              (unsigned long long)wf[0], (unsigned long long)wf[1]);
      t0 = t1;
    };
    o.balance_classes(2 * effort, 24);
    if (dbg) { o.rebuild_rows(); lap("balance"); }
    o.pack_all();
    o.build_incidence();
    o.rebuild_rows();
    lap("pack");
    o.exchange_rows(3 * effort);
    lap("exchange");
    // classes -> a permutation of the local indices: the j-th vertex of class k sits at 8 j + k (the swaps kept
    // every class as large as the residue class of its index)
    const uint32_t blk = o.block ? o.block : nLocal;
    for (uint32_t b0 = 0; b0 < nLocal; b0 += blk) {
      uint32_t nextOf[8];
      for (uint32_t k = 0; k < 8; ++k) nextOf[k] = b0 + k;   // (b0 is a multiple of 8)
      for (uint32_t v = b0; v < std::min(nLocal, b0 + blk); ++v) { newLocal[v] = nextOf[o.cls[v]]; nextOf[o.cls[v]] += 8; }
    }
    for (size_t i = 0; i < (size_t)nCons * 4; ++i)
      if (loc[i] != NONE) loc[i] = newLocal[loc[i]];
  }
  if (stats) {
    for (uint32_t g = 0; g < nGroups; ++g) {
      const int ty = groups[g].arity == 2 ? 0 : 1;
      for (uint32_t q0 = 0; q0 < groups[g].count; q0 += 8)
        for (uint32_t r = 0; r < groups[g].arity; ++r) {
          uint8_t cnt[8] = {};
          for (uint32_t q = q0; q < std::min(groups[g].count, q0 + 8); ++q) cnt[loc[4 * (size_t)(groups[g].begin + q) + r] & 7u]++;
          stats->wavefronts[ty] += max8(cnt);
          stats->ideal[ty] += 1;
        }
    }
  }
}

}  // namespace pbd

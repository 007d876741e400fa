// pbd_plan.h -- host-side schedule builder (pure C++17, no CUDA).
//
// Turns the caller's constraint arrays into a conflict-free parallel schedule that is still a
// Gauss-Seidel sweep over the same constraint set as the reference's sequential loops
// (CProgram/src/Sim.cpp:100-173): within one "group" no two constraints share a vertex, groups
// run one after another.  The schedule is a permutation of the caller's arrays
// (edgeOrder / tetOrder); running the reference on the permuted arrays reproduces what the GPU
// computes, which is how parity is checked.
//
// Two schedule shapes:
//   stream : one global greedy colouring per constraint type; one kernel launch per colour.
//   tile   : vertices are partitioned into shared-memory tiles; several grid-wide phases per
//            constraint type; inside a (phase, tile) constraints are coloured locally and swept
//            with block barriers only (see pbd_tileplan.cpp).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/pbd_b200.h"

namespace pbd {

struct MeshView {
  uint32_t V = 0, E = 0, T = 0;
  const float* x0 = nullptr;        // 3V
  const uint32_t* edges = nullptr;  // 2E
  const uint32_t* tets = nullptr;   // 4T
};

// One colour group of one tile: `count` constraints of one type starting at `begin` in the
// schedule-ordered constraint arrays.
struct Group {
  uint32_t begin = 0;
  uint32_t count = 0;
};

// One (phase, tile): a set of vertices that fits in one SM's shared memory together with the
// constraints projected on it in this phase.  A tile may carry edges, tets or both; its edge colour
// groups run first, then its tet colour groups -- or, `mixed`, step s runs edge group s and tet
// group s together.
struct Tile {
  uint32_t vertBegin = 0;   // into Plan::tileVerts (device vertex slots), or a contiguous range
  uint32_t vertCount = 0;
  uint32_t contiguous = 0;  // 1: the tile's vertices are the slot range [vertBegin, vertBegin+vertCount)
  uint32_t edgeBegin = 0, edgeCount = 0;            // schedule positions of this tile's edges
  uint32_t tetBegin = 0, tetCount = 0;
  uint32_t edgeGroupBegin = 0, edgeGroupCount = 0;  // into Plan::groups (one group per local colour)
  uint32_t tetGroupBegin = 0, tetGroupCount = 0;
  uint32_t edgeDevBegin = 0, tetDevBegin = 0;       // first device index (16-byte aligned ranges)
  uint32_t mixed = 0;       // 1: edge group s and tet group s form colour step s of the visit (same group count, vertex-disjoint)
  uint32_t ride = 0;        // 1: PBD_ORDER_RIDING -- the tile's edge list ends with RIDERS (edges projected by the thread of
                            //    their host tet right after it, no colour group of their own); its record block carries a
                            //    u32 per tet = tile-local edge positions of the (at most two) riders, 0xffff = none
};

struct Phase {
  uint32_t tileBegin = 0, tileCount = 0;
};

struct Plan {
  uint32_t V = 0, E = 0, T = 0;
  uint32_t backend = PBD_BACKEND_STREAM;
  uint32_t orderMode = PBD_ORDER_STRICT;

  // schedule order -> caller index
  std::vector<uint32_t> edgeOrder, tetOrder;

  // stream backend: colour c owns schedule positions [off[c], off[c+1])
  std::vector<uint32_t> edgeColorOff, tetColorOff;

  // tile backend
  std::vector<uint32_t> slotToVertex;  // device vertex slot -> caller vertex (a permutation)
  std::vector<uint32_t> vertexToSlot;
  std::vector<uint32_t> tileVerts;     // gathered vertex slots of non-contiguous tiles
  std::vector<Phase> phases;           // per iteration, in execution order
  std::vector<Tile> tiles;
  std::vector<Group> groups;
  // local (within-tile) vertex indices of every constraint, schedule order
  std::vector<uint16_t> edgeLocal;     // 2E
  std::vector<uint16_t> tetLocal;      // 4T
  uint32_t tileVertexCapacity = 0;     // max vertCount over tiles
  uint32_t tileRecordBytes = 0;        // max tile_record_bytes over tiles (one shared-memory record buffer)
  uint32_t partitions = 0;             // number of shifted vertex partitions (main phases per sweep)
  uint32_t tilesPerPartition = 0;
  // device index of schedule position k (tile backend pads every tile's range to 16 bytes so it
  // can be moved with bulk copies; identity for the stream backend)
  std::vector<uint32_t> edgeDev, tetDev;
  uint32_t edgeDevCount = 0, tetDevCount = 0;
  std::vector<uint32_t> tile0Begin;    // K1+1 slot offsets of the phase-0 (RCB) tiles: a partition of all slots
  uint32_t blockThreads = 0;
  uint32_t tilesPerSm = 1;             // resident tiles (CTAs) per SM the plan was sized for
  uint32_t edgePhases = 0, tetPhases = 0;
  uint32_t edgeColorSum = 0, tetColorSum = 0;  // sum over phases of the max local colour count
  // PBD_ORDER_RIDING: per tet schedule position, the schedule positions of its riders (2 entries,
  // 0xffffffff = none); empty otherwise.  riders = number of edges that ride.
  std::vector<uint32_t> tetRide;
  uint32_t riders = 0;
  // fast arithmetic: per tet schedule position, the role permutation its record carries (see optimise_placement; empty:
  // every tet as the caller gave it).  tetLocal is already in role order; an odd code means the record holds the negated
  // rest volume and the tet's multiplier lives with the opposite sign.
  std::vector<uint8_t> tetPerm;

  // introspection, caller indexing
  std::vector<uint32_t> edgePhase, edgeTile, edgeColor;
  std::vector<uint32_t> tetPhase, tetTile, tetColor;

  // shared-memory wavefronts of the sweeps' vertex gathers under this plan's placement (main tiles; edges, tets)
  uint64_t gatherWavefronts[2] = {0, 0}, gatherIdeal[2] = {0, 0};

  double planMs = 0.0;
};

// Greedy first-fit colouring in the given visiting order: colour[k] = smallest colour unused at
// all of constraint k's vertices.  `arity` = 2 (edges) or 4 (tets); ids are vertex indices
// < nVerts.  Deterministic.  Returns the number of colours.
uint32_t greedy_colour(const uint32_t* ids, uint32_t n, uint32_t arity, uint32_t nVerts,
                       std::vector<uint32_t>& colour);

// Colour the constraints of one whole body (largest-degree-first greedy + recolouring pass) and
// return them sorted by colour, bank-aware inside a colour (pbd_tileplan.cpp): order[k] = constraint
// projected k-th, counts[g] = size of group g (a colour, split into several groups when it exceeds
// maxGroup = what one block pass of the sweep takes).  Used by the batch backend (one body = one tile).
void colour_and_order(const uint32_t* ids, uint32_t n, uint32_t arity, uint32_t nVerts, uint32_t maxGroup,
                      std::vector<uint32_t>& order, std::vector<uint32_t>& counts);

// Shared-memory placement of one tile (pbd_placement.cpp).  A colour group = `count` constraints of one
// arity starting at `begin` in the tile's constraint arrays; the kernel gives thread i of a group its i-th
// constraint, so eight consecutive constraints share a quarter-warp.
struct PlaceGroup {
  uint32_t begin = 0, count = 0, arity = 0;
};
struct PlaceStats {
  uint64_t wavefronts[2] = {0, 0};   // shared-memory wavefronts of the 16-byte vertex gathers (edges, tets) ...
  uint64_t ideal[2] = {0, 0};        // ... and what they would be without bank conflicts (one per quarter-warp and role)
};
// loc: 4 tile-local vertex indices per constraint (0xffffffff beyond the arity), payload: one word per constraint
// (its id).  Reorders both inside every group and renumbers the tile's vertices so that the eight vertices a
// quarter-warp gathers in one role fall into different 16-byte bank groups wherever it finds a way:
// newLocal[old index] = new index (a permutation of 0..nLocal-1); loc is rewritten in the new numbering.
// effort 0: keep numbering and order (statistics only).  block (a multiple of 8, 0 = the whole tile): a vertex
// moves only inside its aligned block of that many indices, which keeps the tile's global loads / stores (thread i
// <-> the slot at index i) as coalesced as they were.  Deterministic.
// perm (one byte per constraint, or null): the four vertices of a TET may take its four roles in any order
// (PBD_FLAG_FAST_ARITH); out: bits 2r..2r+1 = which of the tet's original vertices sits in role r (0xE4 = unchanged).
// An odd permutation flips the sign of the tet's volume: the record builder negates its rest volume (and the sign
// of its multiplier follows).
void optimise_placement(uint32_t nLocal, const PlaceGroup* groups, uint32_t nGroups, uint32_t* loc, uint32_t* payload,
                        uint32_t nCons, int effort, uint32_t block, std::vector<uint32_t>& newLocal, PlaceStats* stats,
                        uint8_t* perm = nullptr);
// sign of a role permutation coded as above: true = odd
inline bool tet_perm_is_odd(uint8_t code) {
  int inv = 0;
  for (int a = 0; a < 4; ++a)
    for (int b = a + 1; b < 4; ++b) inv += ((code >> (2 * a)) & 3) > ((code >> (2 * b)) & 3);
  return (inv & 1) != 0;
}
// rows of 8 by randomised greedy for given residues (4 bytes per constraint); out = constraint indices in row order
void pack_rows_greedy(const uint8_t* res, uint32_t arity, uint32_t n, uint32_t seed, int tries, std::vector<uint32_t>& out);

// Returns false and fills err on invalid input (index >= V).
bool validate_mesh(const MeshView& m, std::string& err);

void build_stream_plan(const MeshView& m, Plan& plan);
bool build_tile_plan(const MeshView& m, const pbd_options& opts, uint32_t nSMs, uint32_t smemBytes,
                     Plan& plan, std::string& err);

// Shared-memory record block of one tile (built by pbd_tile.cu, sized here so the planner can keep
// every tile within the SM's capacity).  Sections, each padded to 16 bytes:
//   header 64 B | predecessor tiles u32[kMaxPreds] | gathered vertex slots u32[nVertGather]
//   | edge groups uint2[] | tet groups uint2[]
//   | edge idx u32[nE] | edge rest f32[nE] | tet idx uint2[nT] | tet rest f32[nT]
//   | edge lambda f32[nE] | tet lambda f32[nT]            (the last two come from the lambda arrays)
inline uint32_t pad4(uint32_t n) { return (n + 3u) & ~3u; }
constexpr uint32_t kMaxPreds = 32;   // tiles of the previous phase a tile can depend on (point-to-point sync)
// (`ride`: one more u32 per tet, right after the tet rest values)
inline uint32_t tile_static_bytes(uint32_t nVertGather, uint32_t nEdgeGroups, uint32_t nTetGroups, uint32_t nE,
                                  uint32_t nT, bool ride = false) {
  return 64u + 4u * kMaxPreds + 4u * pad4(nVertGather) + 8u * (pad4(nEdgeGroups * 2) / 2) + 8u * (pad4(nTetGroups * 2) / 2) +
         8u * pad4(nE) + 8u * (pad4(nT * 2) / 2) + 4u * pad4(nT) + (ride ? 4u * pad4(nT) : 0u);
}
inline uint32_t tile_record_bytes(uint32_t nVertGather, uint32_t nEdgeGroups, uint32_t nTetGroups, uint32_t nE,
                                  uint32_t nT, bool ride = false) {
  return tile_static_bytes(nVertGather, nEdgeGroups, nTetGroups, nE, nT, ride) + 4u * pad4(nE) + 4u * pad4(nT);
}

// Reference init helpers restated for the host (bit-exact, caller's constraint order):
//   inverse masses  CProgram/src/Sim.cpp:63-79 ; rest state  CProgram/src/Sim.cpp:81-95
void host_inverse_mass(const MeshView& m, const uint32_t* pinned, uint32_t nPinned, std::vector<float>& w);
void host_rest_state(const MeshView& m, std::vector<float>& edgeRest, std::vector<float>& tetRest);

uint64_t algorithmic_bytes_per_substep(uint32_t V, uint32_t E, uint32_t T, uint32_t iterations);

}  // namespace pbd

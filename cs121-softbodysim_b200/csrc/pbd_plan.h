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

struct Tile {
  uint32_t vertBegin = 0;   // into Plan::tileVerts (device vertex slots), or a contiguous range
  uint32_t vertCount = 0;
  uint32_t contiguous = 0;  // 1: the tile's vertices are the slot range [vertBegin, vertBegin+vertCount)
  uint32_t groupBegin = 0;  // into Plan::groups
  uint32_t groupCount = 0;
  uint32_t isTet = 0;       // strict order: a tile carries one constraint type
};

struct Phase {
  uint32_t tileBegin = 0, tileCount = 0;
  uint32_t isTet = 0;
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
  std::vector<uint32_t> tile0Begin;    // K1+1 slot offsets of the phase-0 (RCB) tiles: a partition of all slots
  uint32_t blockThreads = 0;           // groups are split so that none exceeds this
  uint32_t edgePhases = 0, tetPhases = 0;
  uint32_t edgeColorSum = 0, tetColorSum = 0;  // sum over phases of the max local colour count

  // introspection, caller indexing
  std::vector<uint32_t> edgePhase, edgeTile, edgeColor;
  std::vector<uint32_t> tetPhase, tetTile, tetColor;

  double planMs = 0.0;
};

// Greedy first-fit colouring in the given visiting order: colour[k] = smallest colour unused at
// all of constraint k's vertices.  `arity` = 2 (edges) or 4 (tets); ids are vertex indices
// < nVerts.  Deterministic.  Returns the number of colours.
uint32_t greedy_colour(const uint32_t* ids, uint32_t n, uint32_t arity, uint32_t nVerts,
                       std::vector<uint32_t>& colour);

// Returns false and fills err on invalid input (index >= V).
bool validate_mesh(const MeshView& m, std::string& err);

void build_stream_plan(const MeshView& m, Plan& plan);
bool build_tile_plan(const MeshView& m, const pbd_options& opts, uint32_t nSMs, uint32_t smemVertexLimit,
                     Plan& plan, std::string& err);

// Reference init helpers restated for the host (bit-exact, caller's constraint order):
//   inverse masses  CProgram/src/Sim.cpp:63-79 ; rest state  CProgram/src/Sim.cpp:81-95
void host_inverse_mass(const MeshView& m, const uint32_t* pinned, uint32_t nPinned, std::vector<float>& w);
void host_rest_state(const MeshView& m, std::vector<float>& edgeRest, std::vector<float>& tetRest);

uint64_t algorithmic_bytes_per_substep(uint32_t V, uint32_t E, uint32_t T, uint32_t iterations);

}  // namespace pbd

// CudaStepper.cpp -- see CudaStepper.h.  Compiled against the reference's PBDServer.h in place
// (integration/Makefile); zero warnings under -Wall -Wextra -Wpedantic is a test.
#include "CudaStepper.h"

namespace {

// quiet NaN whose payload is the generation (never 0): distinguishable from the +0.0f of build_rest
float stamp_value(uint32_t generation) {
  const uint32_t bits = 0x7fc00000u | (generation & 0x003fffffu);
  float f;
  std::memcpy(&f, &bits, sizeof f);
  return f;
}
uint32_t bits_of(float f) {
  uint32_t b;
  std::memcpy(&b, &f, sizeof b);
  return b;
}

}  // namespace

CudaStepper::CudaStepper(int device, const pbd_options* opts) : device_(device) {
  if (opts) { opts_ = *opts; haveOpts_ = true; }
}

CudaStepper::~CudaStepper() { pbd_destroy(h_); }

const char* CudaStepper::name() const { return h_ ? pbd_backend_name(h_) : "b200"; }

bool CudaStepper::bound_to(const PBDState& s) const {
  if (!h_ || s.V != V_ || s.E != E_ || s.T != T_) return false;
#ifdef PBD_STATE_HAS_GENERATION
  return s.generation == generation_;
#else
  if (!stamped_) return false;                       // nothing to carry the stamp: rebuild (no constraints, cheap)
  const float slot = !s.edgeLambda.empty() ? s.edgeLambda[0] : s.tetLambda[0];
  return bits_of(slot) == bits_of(stamp_value(generation_));
#endif
}

void CudaStepper::stamp(PBDState& s) {
#ifdef PBD_STATE_HAS_GENERATION
  generation_ = (uint32_t)s.generation;
  stamped_ = true;
#else
  stamped_ = false;
  if (!h_) return;
  if (!s.edgeLambda.empty()) { s.edgeLambda[0] = stamp_value(generation_); stamped_ = true; }
  else if (!s.tetLambda.empty()) { s.tetLambda[0] = stamp_value(generation_); stamped_ = true; }
#endif
}

void CudaStepper::bind(const PBDState& s) {
  pbd_destroy(h_);
  h_ = nullptr;
  V_ = s.V; E_ = s.E; T_ = s.T;
  generation_ = (generation_ % 0x003ffffeu) + 1u;    // 1 .. 2^22-2, never 0
  ++binds_;

  std::vector<uint32_t> edges(2 * (size_t)E_), tets(4 * (size_t)T_), pinned;
  for (uint32_t k = 0; k < E_; ++k) { edges[2 * (size_t)k] = s.edgeI0[k]; edges[2 * (size_t)k + 1] = s.edgeI1[k]; }
  for (uint32_t k = 0; k < T_; ++k) {
    tets[4 * (size_t)k] = s.tetA[k]; tets[4 * (size_t)k + 1] = s.tetB[k];
    tets[4 * (size_t)k + 2] = s.tetC[k]; tets[4 * (size_t)k + 3] = s.tetD[k];
  }
  // comm_loop has already run compute_inv_mass (Server.cpp:103): w == 0 <=> pinned or in no tet, and a
  // vertex in no tet gets w = 0 from pbd_create as well, so "pinned := every w == 0 vertex" reproduces
  // the reference's inverse masses exactly (the pinned list itself is not kept in PBDState)
  for (uint32_t i = 0; i < V_; ++i)
    if (s.w[i] == 0.0f) pinned.push_back(i);

  pbd_params p{};                                    // SolverParams (PBDServer.h:147-161) -> wire order
  p.substeps = s.params.substeps; p.iterations = s.params.iterations;
  p.dtHint = s.params.dtHint; p.omega = s.params.omega;
  p.edgeCompliance = s.params.edgeCompliance; p.volumeCompliance = s.params.volumeCompliance;
  p.gx = s.params.gravity.x; p.gy = s.params.gravity.y; p.gz = s.params.gravity.z;
  p.groundEnabled = s.params.groundEnabled; p.groundY = s.params.groundY; p.friction = s.params.friction;

  static_assert(sizeof(Vec3) == 3 * sizeof(float), "Vec3 is three packed floats (PBDServer.h:123-127)");
  const float* x0 = V_ ? &s.x[0].x : nullptr;        // V == 0: no element to take the address of
  int st = PBD_OK;
  h_ = pbd_create(&p, V_, E_, T_, x0, edges.data(), tets.data(), pinned.data(), (uint32_t)pinned.size(), device_,
                  haveOpts_ ? &opts_ : nullptr, &st);
  ok_ = h_ != nullptr;
  if (!ok_) {
    err_ = pbd_last_error();
    std::fprintf(stderr, "[PBDServer] gpu stepper: pbd_create failed (%d): %s\n", st, err_.c_str());
  }
}

void CudaStepper::step(PBDState& s, float dt, perf::StepStats& out) {
  if (!bound_to(s)) { bind(s); stamp(s); }
  if (!h_) return;
  pbd_step_stats g{};
  const int rc = pbd_step(h_, dt, &g);               // ADDS, like the reference steppers (Sim.cpp:283-304)
  if (rc != PBD_OK) {
    ok_ = false;
    err_ = pbd_last_error();
    std::fprintf(stderr, "[PBDServer] gpu stepper: pbd_step failed (%d): %s\n", rc, err_.c_str());
    return;
  }
  out.predictMs += g.predictMs; out.solveMs += g.solveMs; out.commitMs += g.commitMs; out.totalMs += g.totalMs;
}

void CudaStepper::pack_positions(const PBDState& s, std::vector<float>& outPos, double& outPackMs) {
  if (outPos.size() != 3 * (size_t)s.V) outPos.resize(3 * (size_t)s.V);
  if (!bound_to(s)) {
    // pack before any step of this INIT (the reference would return x0): the state is const here,
    // so it cannot be stamped -- hand back the host positions, the next step() binds
    for (uint32_t i = 0; i < s.V; ++i) { outPos[3 * (size_t)i] = s.x[i].x; outPos[3 * (size_t)i + 1] = s.x[i].y; outPos[3 * (size_t)i + 2] = s.x[i].z; }
    return;
  }
  if (s.V && pbd_read_positions(h_, outPos.data(), &outPackMs) != PBD_OK) {
    ok_ = false;
    err_ = pbd_last_error();
    std::fprintf(stderr, "[PBDServer] gpu stepper: pbd_read_positions failed: %s\n", err_.c_str());
  }
}

// ref_harness.cpp -- ORACLE (test infrastructure, NOT product code).
//
// A thin extern "C" face over the UNMODIFIED reference solver so Python can drive it
// in-process.  It is compiled together with /root/reference/CProgram/src/Sim.cpp *where that
// file lies* (see oracle/Makefile); no reference source is copied into this repository and
// only the resulting oracle/_ref/libpbdref.so travels to the GPU box.
//
// What comes from the reference (called, not restated): compute_inv_mass, build_rest,
// SerialStepper::step / ParallelStepper::step, *::pack_positions
// (CProgram/include/PBDServer.h:182-280, CProgram/src/Sim.cpp:63-361).
// What is ours: filling a PBDState from flat arrays the way comm_loop does for MSG_INIT
// (CProgram/src/Server.cpp:72-104), and permuting the constraint arrays.
//
// The exported names mirror oracle/pbd_oracle.c (prefix pbdr_ instead of pbdo_) so one
// ctypes wrapper (oracle/pyoracle.py) serves both.
#include "PBDServer.h"

#include <memory>

namespace {

struct Params {  // == pbdo_params, MSG_INIT wire order (Server.cpp:38-50)
  uint32_t substeps, iterations;
  float dtHint, omega;
  float edgeCompliance, volumeCompliance;
  float gx, gy, gz;
  uint32_t groundEnabled;
  float groundY, friction;
};

struct Box {
  PBDState st;
  perf::StepStats acc{};
  SerialStepper serial;
  std::unique_ptr<ParallelStepper> parallel;
  IStepper* stepper = &serial;
};

template <class T>
void gather(std::vector<T>& a, const uint32_t* order) {
  std::vector<T> t(a.size());
  for (size_t k = 0; k < a.size(); ++k) t[k] = a[order[k]];
  a.swap(t);
}

}  // namespace

extern "C" {

Box* pbdr_create(const Params* p, uint32_t V, uint32_t E, uint32_t T, const float* x0,
                 const uint32_t* edgeIds, const uint32_t* tetIds, const uint32_t* pinned,
                 uint32_t nPinned) {
  Box* b = new Box();
  PBDState& s = b->st;
  s.V = V; s.E = E; s.T = T;
  s.params.substeps = p->substeps;
  s.params.iterations = p->iterations;
  s.params.dtHint = p->dtHint;
  s.params.omega = p->omega;
  s.params.edgeCompliance = p->edgeCompliance;
  s.params.volumeCompliance = p->volumeCompliance;
  s.params.gravity = Vec3(p->gx, p->gy, p->gz);
  s.params.groundEnabled = p->groundEnabled;
  s.params.groundY = p->groundY;
  s.params.friction = p->friction;

  s.x.resize(V);
  s.v.assign(V, Vec3(0, 0, 0));
  s.xStar.resize(V);
  for (uint32_t i = 0; i < V; ++i) {
    s.x[i] = Vec3(x0[3 * i], x0[3 * i + 1], x0[3 * i + 2]);
    s.xStar[i] = s.x[i];
  }
  s.edgeI0.resize(E); s.edgeI1.resize(E);
  for (uint32_t e = 0; e < E; ++e) { s.edgeI0[e] = edgeIds[2 * e]; s.edgeI1[e] = edgeIds[2 * e + 1]; }
  s.tetA.resize(T); s.tetB.resize(T); s.tetC.resize(T); s.tetD.resize(T);
  for (uint32_t t = 0; t < T; ++t) {
    s.tetA[t] = tetIds[4 * t]; s.tetB[t] = tetIds[4 * t + 1];
    s.tetC[t] = tetIds[4 * t + 2]; s.tetD[t] = tetIds[4 * t + 3];
  }
  std::vector<uint32_t> pin(pinned, pinned + nPinned);
  compute_inv_mass(s, pin);   // reference code
  build_rest(s);              // reference code
  return b;
}

void pbdr_destroy(Box* b) { delete b; }

// threads == 0 -> SerialStepper ; threads >= 1 -> ParallelStepper(threads)
void pbdr_use_parallel(Box* b, uint32_t threads) {
  if (threads == 0) { b->parallel.reset(); b->stepper = &b->serial; return; }
  b->parallel.reset(new ParallelStepper(threads));
  b->stepper = b->parallel.get();
}

void pbdr_permute_constraints(Box* b, const uint32_t* edgeOrder, const uint32_t* tetOrder) {
  PBDState& s = b->st;
  if (edgeOrder) {
    gather(s.edgeI0, edgeOrder); gather(s.edgeI1, edgeOrder);
    gather(s.edgeRest, edgeOrder); gather(s.edgeLambda, edgeOrder);
  }
  if (tetOrder) {
    gather(s.tetA, tetOrder); gather(s.tetB, tetOrder); gather(s.tetC, tetOrder); gather(s.tetD, tetOrder);
    gather(s.tetRestVol, tetOrder); gather(s.tetLambda, tetOrder);
  }
}

void pbdr_step(Box* b, float dt) { b->stepper->step(b->st, dt, b->acc); }  // reference code

void pbdr_pack(Box* b, float* out) {
  std::vector<float> tmp;
  b->stepper->pack_positions(b->st, tmp, b->acc.packMs);  // reference code
  std::memcpy(out, tmp.data(), tmp.size() * sizeof(float));
}

void pbdr_get(const Box* b, int what, void* out) {
  const PBDState& s = b->st;
  auto v3 = [&](const std::vector<Vec3>& a) {
    float* o = static_cast<float*>(out);
    for (uint32_t i = 0; i < s.V; ++i) { o[3 * i] = a[i].x; o[3 * i + 1] = a[i].y; o[3 * i + 2] = a[i].z; }
  };
  switch (what) {
    case 0: std::memcpy(out, s.w.data(), sizeof(float) * s.V); break;
    case 1: std::memcpy(out, s.edgeRest.data(), sizeof(float) * s.E); break;
    case 2: std::memcpy(out, s.tetRestVol.data(), sizeof(float) * s.T); break;
    case 3: std::memcpy(out, s.edgeLambda.data(), sizeof(float) * s.E); break;
    case 4: std::memcpy(out, s.tetLambda.data(), sizeof(float) * s.T); break;
    case 5: v3(s.v); break;
    case 6: v3(s.xStar); break;
    default: break;
  }
}

void pbdr_set_inv_mass(Box* b, const float* w) { std::memcpy(b->st.w.data(), w, sizeof(float) * b->st.V); }

void pbdr_stats(Box* b, double* out5, int reset) {
  out5[0] = b->acc.predictMs; out5[1] = b->acc.solveMs; out5[2] = b->acc.commitMs;
  out5[3] = b->acc.packMs; out5[4] = b->acc.totalMs;
  if (reset) b->acc = perf::StepStats{};
}

const char* pbdr_name(void) { return "reference"; }

}  // extern "C"

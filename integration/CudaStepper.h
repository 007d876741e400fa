// CudaStepper.h -- the reference's IStepper seam (CProgram/include/PBDServer.h:261-266) on the GPU.
//
// This file is the adapter a maintainer of Captain-Noble/CS121-softbodysim adds next to
// SerialStepper / ParallelStepper (CProgram/include/PBDServer.h:268-280).  It includes the
// REFERENCE's own header (found on the include path at build time -- nothing of the reference is
// copied into this repository) and this repository's C ABI, include/pbd_b200.h.
//
//   name()            -> pbd_backend_name                        PBDServer.h:263
//   step()            -> pbd_step          (ADDS into StepStats) PBDServer.h:264, Sim.cpp:280-305
//   pack_positions()  -> pbd_read_positions (3V floats, x)       PBDServer.h:265, Sim.cpp:307-316
//
// PBDState stays the host-side description of the body exactly as comm_loop fills it
// (CProgram/src/Server.cpp:72-104); all solver state lives in HBM inside the pbd_handle.
//
// Re-INIT.  comm_loop move-assigns every new MSG_INIT into the SAME Shared::state object
// (Server.cpp:106-110), so neither the address of the PBDState nor its V/E/T tell a second INIT from
// the first (a client restarting the same scene sends identical counts).  The adapter therefore
// STAMPS the state it has bound: build_rest leaves every lambda at +0.0f (Sim.cpp:83,90) and nothing
// on the GPU path ever reads the host lambdas, so bind() writes a quiet NaN carrying its generation
// number into the first lambda slot; a fresh INIT brings +0.0f back and the next step()/pack rebuilds
// the device state.  (With PBD_STATE_HAS_GENERATION defined the adapter uses a `uint64_t generation`
// member instead -- the one-line patch to PBDState + `++generation` in comm_loop a maintainer may prefer.)
#pragma once
#include "PBDServer.h"
#include "pbd_b200.h"

struct CudaStepper final : IStepper {
  explicit CudaStepper(int device = 0, const pbd_options* opts = nullptr);
  ~CudaStepper() override;
  CudaStepper(const CudaStepper&) = delete;
  CudaStepper& operator=(const CudaStepper&) = delete;

  const char* name() const override;
  void step(PBDState& s, float dt, perf::StepStats& out) override;
  void pack_positions(const PBDState& s, std::vector<float>& outPos, double& outPackMs) override;

  // number of times the device state was (re)built -- one per MSG_INIT that was followed by a step
  unsigned binds() const { return binds_; }
  // false after a failed pbd_create / pbd_step: message in last_error(); the reference's steppers
  // cannot fail, so the adapter prints the message once and leaves the positions unchanged
  bool ok() const { return ok_; }
  const char* last_error() const { return err_.c_str(); }

 private:
  bool bound_to(const PBDState& s) const;
  void bind(const PBDState& s);
  void stamp(PBDState& s);

  pbd_handle* h_ = nullptr;
  pbd_options opts_{};
  bool haveOpts_ = false;
  int device_ = 0;
  uint32_t V_ = 0, E_ = 0, T_ = 0;
  uint32_t generation_ = 0;     // stamped into the bound state
  bool stamped_ = false;        // the bound state carries generation_ (it has a lambda slot to carry it)
  unsigned binds_ = 0;
  bool ok_ = true;
  std::string err_;
};

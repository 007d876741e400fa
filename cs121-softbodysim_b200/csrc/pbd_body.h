// pbd_body.h -- device-resident state of one (or a batch of) soft bodies + backend interface.
//
// HBM layout (SoA of float4, SURVEY.md 7 "float4 SoA"):
//   pos [V] float4  (xStar.x, xStar.y, xStar.z, invMass)   -- the array every projection gathers/scatters
//   prev[V] float4  (x.x, x.y, x.z, unused)                -- committed positions
//   vel [V] float4  (v.x, v.y, v.z, unused)
//   edgeRest/edgeLam [E] f32, tetRest/tetLam [T] f32        -- schedule order
//   constraint indices: backend specific (stream: uint2/uint4 vertex slots; tile: u16 tile-local)
// "slot" = device vertex index; slotOf[callerVertex] (null = identity) maps the caller's order.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "pbd_math.cuh"
#include "pbd_plan.h"

namespace pbd {

struct DeviceArrays {
  uint32_t V = 0, E = 0, T = 0;
  float4* pos = nullptr;
  float4* prev = nullptr;
  float4* vel = nullptr;
  float* edgeRest = nullptr;
  float* edgeLam = nullptr;
  float* tetRest = nullptr;
  float* tetLam = nullptr;
  uint32_t* slotOf = nullptr;   // [V] caller vertex -> slot (nullptr: identity)
  float* packed = nullptr;      // [3V] readback staging, caller order
  StepConsts* consts = nullptr; // device copy of the per-frame scalars
  ColliderSet* colliders = nullptr;   // device copy of the primitive colliders (pbd_set_colliders)
  uint32_t nColliders = 0;
  uint64_t bytes = 0;
};

struct FrameShape {
  uint32_t substeps = 1;    // already clamped to >= 1
  uint32_t iterations = 0;
  int groundEnabled = 0;
  uint32_t nColliders = 0;  // primitive colliders in the clamp stage (pbd_set_colliders)
  bool tetInert = false;    // alphaTet == 0 this frame: tet multipliers never enter a correction (fast mode may drop them)
};

class Backend {
 public:
  virtual ~Backend() {}
  virtual const char* name() const = 0;
  // upload the backend-specific constraint tables; returns cudaSuccess or the failing error
  virtual cudaError_t upload(const Plan& plan, const MeshView& mesh, DeviceArrays& d) = 0;
  // enqueue one frame (all substeps) on `s`; the per-frame scalars are already in d.consts
  virtual cudaError_t enqueue_frame(const DeviceArrays& d, const FrameShape& f, cudaStream_t s) = 0;
  virtual uint32_t launches_per_frame(const FrameShape& f) const = 0;
  virtual void invalidate() {}
  virtual void set_omega(float) {}   // SOR factor (Jacobi comparison backend: pbd_params.omega)
  virtual void debug_dump() {}  // PBD_TILE_TRACE: per-phase timing of the last frame to stderr  // frame shape changed (set_params)
  virtual uint64_t device_bytes() const = 0;
  virtual void fill_info(pbd_info& info) const {}
  // true (once) if a kernel gave up waiting for another CTA / GPU since the last call (bounded spins)
  virtual bool take_abort() { return false; }
  // bring d.pos up to date before the host reads it (backends that keep positions in another layout)
  virtual cudaError_t export_pos(const DeviceArrays&, cudaStream_t) { return cudaSuccess; }
  // one body across several GPUs (tile backend only)
  virtual uint32_t shard_world() const { return 1; }
  virtual uint32_t shard_rank() const { return 0; }
  virtual void shard_slot_ranges(std::vector<uint32_t>& begin) const { begin.clear(); }
  virtual cudaError_t shard_export(void*) { return cudaErrorNotSupported; }
  virtual cudaError_t shard_attach_ipc(const void*) { return cudaErrorNotSupported; }
  virtual void shard_local_pointers(void** pos, void** done) { *pos = nullptr; *done = nullptr; }
  virtual cudaError_t shard_attach_pointers(uint32_t, void*, void*, int) { return cudaErrorNotSupported; }
  // optional per-stage timing (stream backend); returns false if unsupported
  virtual bool stage_ms(double& predict, double& solve, double& commit) { return false; }
  // fused backends: the predict / commit stages as shares of the last frame's device time; false if unsupported
  virtual bool stage_share(double& predict, double& commit) { return false; }
};

Backend* make_stream_backend(uint32_t flags, uint32_t blockThreads);
Backend* make_tile_backend(const pbd_options& opts, int device);
Backend* make_jacobi_backend(const pbd_options& opts);

// shared vertex-stage kernels (pbd_stream.cu)
cudaError_t launch_pack(const DeviceArrays& d, cudaStream_t s);
// per-vertex normals of the surface triangles into d.packed (caller vertex order)
cudaError_t launch_normals(const DeviceArrays& d, const uint32_t* tris, const uint32_t* adjOff, const uint32_t* adjTri, cudaStream_t s);

StepConsts make_consts(const pbd_params& p, float dt);

// Every entry point runs on its handle's device and leaves the caller's current device as it found it.
struct DeviceScope {
  int prev = -1;
  bool switched = false;
  cudaError_t err = cudaSuccess;
  explicit DeviceScope(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
    if (prev != dev) { err = cudaSetDevice(dev); switched = err == cudaSuccess; }
  }
  ~DeviceScope() { if (switched && prev >= 0) cudaSetDevice(prev); }
  DeviceScope(const DeviceScope&) = delete;
  DeviceScope& operator=(const DeviceScope&) = delete;
};

// message returned by pbd_last_error() on this thread (pbd_capi.cu)
void set_last_error(const std::string& msg);

}  // namespace pbd

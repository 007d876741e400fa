// pbd_capi.cu -- the extern "C" boundary declared in include/pbd_b200.h.
//
// Host side of the drop-in: what comm_loop does for MSG_INIT (CProgram/src/Server.cpp:30-114)
// becomes pbd_create (validate, derive w / rest values on the host exactly like
// compute_inv_mass / build_rest, build the parallel schedule, upload once); IStepper::step
// becomes pbd_step; IStepper::pack_positions becomes pbd_read_positions.  No CPU fallback:
// without a CUDA device every computing entry fails with PBD_ERR_NO_DEVICE.
#include <chrono>
#include <cmath>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "pbd_body.h"

using namespace pbd;

namespace {

thread_local std::string g_err;

// shared memory the tile kernel keeps for itself (static variables, alignment slack)
constexpr uint32_t kSmemReserve = 8192;

double wall_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

int fail(int code, const std::string& msg, int* status = nullptr) {
  g_err = msg;
  if (status) *status = code;
  return code;
}

int cuda_fail(cudaError_t e, const char* what, int* status = nullptr) {
  const int code = (e == cudaErrorMemoryAllocation) ? PBD_ERR_OOM
                   : (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? PBD_ERR_NO_DEVICE
                                                                                   : PBD_ERR_CUDA;
  return fail(code, std::string(what) + ": " + cudaGetErrorString(e), status);
}

#define CU(call)                                                   \
  do {                                                             \
    cudaError_t e__ = (call);                                      \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call);          \
  } while (0)

pbd_options resolve_options(const pbd_options* in) {
  pbd_options o;
  std::memset(&o, 0, sizeof(o));
  if (in) {
    size_t n = in->struct_size ? in->struct_size : sizeof(pbd_options);
    if (n > sizeof(pbd_options)) n = sizeof(pbd_options);
    std::memcpy(&o, in, n);
  }
  o.struct_size = sizeof(pbd_options);
  return o;
}

template <class T>
cudaError_t dev_alloc(T** p, size_t n, uint64_t& bytes) {
  cudaError_t e = cudaMalloc((void**)p, sizeof(T) * (n + 1));
  if (e == cudaSuccess) bytes += sizeof(T) * n;
  return e;
}

void free_arrays(DeviceArrays& d) {
  cudaFree(d.pos); cudaFree(d.prev); cudaFree(d.vel);
  cudaFree(d.edgeRest); cudaFree(d.edgeLam); cudaFree(d.tetRest); cudaFree(d.tetLam);
  cudaFree(d.slotOf); cudaFree(d.packed); cudaFree(d.consts); cudaFree(d.colliders);
  d = DeviceArrays{};
}

}  // namespace

namespace pbd {
void set_last_error(const std::string& msg) { g_err = msg; }
}  // namespace pbd

struct pbd_plan {
  Plan plan;
  pbd_options opts;
};

struct pbd_handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  pbd_params params{};
  pbd_options opts{};
  Plan plan;
  DeviceArrays d;
  std::unique_ptr<Backend> be;
  double uploadMs = 0.0;
  bool pending = false;
  uint32_t* surfTris = nullptr;   // pbd_set_surface: triangles + per-vertex adjacency (CSR), caller vertex order
  uint32_t* surfAdjOff = nullptr;
  uint32_t* surfAdjTri = nullptr;
  uint32_t nSurfTris = 0;

  ~pbd_handle() {
    be.reset();
    free_arrays(d);
    cudaFree(surfTris); cudaFree(surfAdjOff); cudaFree(surfAdjTri);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (stream) cudaStreamDestroy(stream);
  }
  FrameShape shape() const {
    FrameShape f;
    f.substeps = params.substeps > 1u ? params.substeps : 1u;
    f.iterations = params.iterations;
    f.groundEnabled = params.groundEnabled ? 1 : 0;
    f.nColliders = d.nColliders;
    return f;
  }
};

namespace {

bool build_plan(const MeshView& m, const pbd_options& o, int nSMs, uint32_t smemBytes, Plan& plan, std::string& err) {
  uint32_t backend = o.backend;
  // auto: the tile backend (one persistent kernel per frame); the stream backend (one launch per
  // global colour) remains as the simple cross-check
  if (backend == PBD_BACKEND_AUTO) backend = PBD_BACKEND_TILE;
  if (backend == PBD_BACKEND_STREAM) {
    if (o.order_mode != PBD_ORDER_STRICT) { err = "stream backend supports PBD_ORDER_STRICT only"; return false; }
    build_stream_plan(m, plan);
    return true;
  }
  if (backend == PBD_BACKEND_TILE) return build_tile_plan(m, o, (uint32_t)nSMs, smemBytes, plan, err);
  if (backend == PBD_BACKEND_JACOBI) {
    // no schedule: every vertex gathers its own constraints; caller's order throughout
    plan = Plan();
    plan.V = m.V; plan.E = m.E; plan.T = m.T;
    plan.backend = PBD_BACKEND_JACOBI;
    plan.edgeOrder.resize(m.E); plan.tetOrder.resize(m.T); plan.edgeDev.resize(m.E); plan.tetDev.resize(m.T);
    for (uint32_t k = 0; k < m.E; ++k) plan.edgeOrder[k] = plan.edgeDev[k] = k;
    for (uint32_t k = 0; k < m.T; ++k) plan.tetOrder[k] = plan.tetDev[k] = k;
    plan.edgeDevCount = m.E; plan.tetDevCount = m.T;
    plan.slotToVertex.resize(m.V); plan.vertexToSlot.resize(m.V);
    for (uint32_t i = 0; i < m.V; ++i) plan.slotToVertex[i] = plan.vertexToSlot[i] = i;
    plan.edgePhase.assign(m.E, 0); plan.edgeTile.assign(m.E, 0); plan.edgeColor.assign(m.E, 0);
    plan.tetPhase.assign(m.T, 0); plan.tetTile.assign(m.T, 0); plan.tetColor.assign(m.T, 0);
    return true;
  }
  err = "unknown backend";
  return false;
}

void base_info(const Plan& p, const pbd_params* prm, pbd_info& out) {
  std::memset(&out, 0, sizeof(out));
  out.V = p.V; out.E = p.E; out.T = p.T;
  out.backend = p.backend;
  out.edge_colors = p.edgeColorSum; out.tet_colors = p.tetColorSum;
  out.edge_phases = p.edgePhases; out.tet_phases = p.tetPhases;
  out.tiles = (uint32_t)p.tiles.size();
  out.partitions = p.partitions;
  out.plan_ms = p.planMs;
  for (int ty = 0; ty < 2; ++ty)
    out.gather_wavefronts_permille[ty] = p.gatherIdeal[ty] ? (uint32_t)((1000ull * p.gatherWavefronts[ty] + p.gatherIdeal[ty] / 2) / p.gatherIdeal[ty]) : 0u;
  out.algorithmic_bytes_per_substep = algorithmic_bytes_per_substep(p.V, p.E, p.T, prm ? prm->iterations : 6);
}

}  // namespace

extern "C" {

int pbd_abi_version(void) { return PBD_ABI_VERSION; }
const char* pbd_last_error(void) { return g_err.c_str(); }

int pbd_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

pbd_handle* pbd_create(const pbd_params* params, uint32_t V, uint32_t E, uint32_t T, const float* x0,
                       const uint32_t* edgeIds, const uint32_t* tetIds, const uint32_t* pinned,
                       uint32_t nPinned, int device, const pbd_options* opts, int* status) {
  if (status) *status = PBD_OK;
  if (!params) { fail(PBD_ERR_INVALID, "params is null", status); return nullptr; }
  if (nPinned && !pinned) { fail(PBD_ERR_INVALID, "pinned is null", status); return nullptr; }
  MeshView m{V, E, T, x0, edgeIds, tetIds};
  std::string err;
  if (!validate_mesh(m, err)) {
    fail(err.find("out of range") != std::string::npos ? PBD_ERR_INDEX : PBD_ERR_INVALID, err, status);
    return nullptr;
  }
  int nDev = 0;
  cudaError_t ce = cudaGetDeviceCount(&nDev);
  if (ce != cudaSuccess || nDev == 0) {
    cudaGetLastError();
    fail(PBD_ERR_NO_DEVICE, "no CUDA device (this library has no CPU fallback)", status);
    return nullptr;
  }
  if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
  if (device >= nDev) { fail(PBD_ERR_INVALID, "device ordinal out of range", status); return nullptr; }
  DeviceScope onDevice(device);
  if ((ce = onDevice.err) != cudaSuccess) { cuda_fail(ce, "cudaSetDevice", status); return nullptr; }

  std::unique_ptr<pbd_handle> h(new pbd_handle());
  h->device = device;
  h->params = *params;
  h->opts = resolve_options(opts);

  cudaDeviceProp prop{};
  if ((ce = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { cuda_fail(ce, "cudaGetDeviceProperties", status); return nullptr; }
  const uint32_t smemBytes = (uint32_t)(prop.sharedMemPerBlockOptin > kSmemReserve ? prop.sharedMemPerBlockOptin - kSmemReserve : 0);
  const uint32_t world = h->opts.shard_world > 1 ? h->opts.shard_world : 1u;
  if (world > 1 && (h->opts.shard_rank >= world || world > 8)) { fail(PBD_ERR_INVALID, "shard_rank/shard_world out of range", status); return nullptr; }
  if (world > 1 && h->opts.backend == PBD_BACKEND_STREAM) { fail(PBD_ERR_UNSUPPORTED, "a sharded body needs the tile backend", status); return nullptr; }
  // one tile per SM of every GPU the body is spread over (or what the caller asks to plan for)
  const int planSMs = h->opts.plan_sms ? (int)h->opts.plan_sms : prop.multiProcessorCount * (int)world;
  if (!build_plan(m, h->opts, planSMs, smemBytes, h->plan, err)) {
    fail(PBD_ERR_INVALID, err, status);
    return nullptr;
  }
  const Plan& plan = h->plan;

  // reference init helpers on the host, caller's order (bit-exact): Sim.cpp:63-95
  std::vector<float> w, eRest, tRest;
  host_inverse_mass(m, pinned, nPinned, w);
  host_rest_state(m, eRest, tRest);

  const double tUp = wall_ms();
  DeviceArrays& d = h->d;
  d.V = V; d.E = E; d.T = T;
  auto bail = [&](cudaError_t e, const char* what) { cuda_fail(e, what, status); return nullptr; };
  if ((ce = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(ce, "cudaStreamCreate");
  if ((ce = cudaEventCreate(&h->ev0)) != cudaSuccess) return bail(ce, "cudaEventCreate");
  if ((ce = cudaEventCreate(&h->ev1)) != cudaSuccess) return bail(ce, "cudaEventCreate");
  if ((ce = dev_alloc(&d.pos, V, d.bytes)) != cudaSuccess) return bail(ce, "cudaMalloc pos");
  if ((ce = dev_alloc(&d.prev, V, d.bytes)) != cudaSuccess) return bail(ce, "cudaMalloc prev");
  if ((ce = dev_alloc(&d.vel, V, d.bytes)) != cudaSuccess) return bail(ce, "cudaMalloc vel");
  // constraint arrays live at the plan's device indices (schedule order; the tile backend pads
  // every tile's range to 16 bytes)
  const size_t nE = plan.edgeDevCount, nT = plan.tetDevCount;
  if ((ce = dev_alloc(&d.edgeRest, nE, d.bytes)) != cudaSuccess) return bail(ce, "cudaMalloc edgeRest");
  if ((ce = dev_alloc(&d.edgeLam, nE, d.bytes)) != cudaSuccess) return bail(ce, "cudaMalloc edgeLam");
  if ((ce = dev_alloc(&d.tetRest, nT, d.bytes)) != cudaSuccess) return bail(ce, "cudaMalloc tetRest");
  if ((ce = dev_alloc(&d.tetLam, nT, d.bytes)) != cudaSuccess) return bail(ce, "cudaMalloc tetLam");
  if ((ce = dev_alloc(&d.packed, (size_t)V * 3, d.bytes)) != cudaSuccess) return bail(ce, "cudaMalloc packed");
  if ((ce = dev_alloc(&d.consts, 1, d.bytes)) != cudaSuccess) return bail(ce, "cudaMalloc consts");
  if ((ce = dev_alloc(&d.colliders, 1, d.bytes)) != cudaSuccess) return bail(ce, "cudaMalloc colliders");
  if ((ce = cudaMemset(d.colliders, 0, sizeof(ColliderSet))) != cudaSuccess) return bail(ce, "memset colliders");

  bool identity = true;
  for (uint32_t i = 0; i < V && identity; ++i) identity = plan.vertexToSlot[i] == i;
  if (!identity) {
    if ((ce = dev_alloc(&d.slotOf, V, d.bytes)) != cudaSuccess) return bail(ce, "cudaMalloc slotOf");
    if ((ce = cudaMemcpy(d.slotOf, plan.vertexToSlot.data(), sizeof(uint32_t) * V, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(ce, "upload slotOf");
  }
  {
    std::vector<float4> pos(V), prev(V);
    for (uint32_t v = 0; v < V; ++v) {
      const uint32_t s = plan.vertexToSlot[v];
      pos[s] = make_float4(x0[3 * (size_t)v], x0[3 * (size_t)v + 1], x0[3 * (size_t)v + 2], w[v]);
      prev[s] = make_float4(x0[3 * (size_t)v], x0[3 * (size_t)v + 1], x0[3 * (size_t)v + 2], 0.0f);
    }
    if (V && (ce = cudaMemcpy(d.pos, pos.data(), sizeof(float4) * V, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(ce, "upload pos");
    if (V && (ce = cudaMemcpy(d.prev, prev.data(), sizeof(float4) * V, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(ce, "upload prev");
    if ((ce = cudaMemset(d.vel, 0, sizeof(float4) * ((size_t)V + 1))) != cudaSuccess) return bail(ce, "memset vel");
    std::vector<float> tmp(std::max(nE, nT), 0.0f);
    for (uint32_t k = 0; k < E; ++k) tmp[plan.edgeDev[k]] = eRest[plan.edgeOrder[k]];
    if (nE && (ce = cudaMemcpy(d.edgeRest, tmp.data(), sizeof(float) * nE, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(ce, "upload edgeRest");
    std::fill(tmp.begin(), tmp.end(), 0.0f);
    for (uint32_t k = 0; k < T; ++k) tmp[plan.tetDev[k]] = tRest[plan.tetOrder[k]];
    if (nT && (ce = cudaMemcpy(d.tetRest, tmp.data(), sizeof(float) * nT, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(ce, "upload tetRest");
    if ((ce = cudaMemset(d.edgeLam, 0, sizeof(float) * (nE + 1))) != cudaSuccess) return bail(ce, "memset edgeLam");
    if ((ce = cudaMemset(d.tetLam, 0, sizeof(float) * (nT + 1))) != cudaSuccess) return bail(ce, "memset tetLam");
  }
  if (plan.backend == PBD_BACKEND_STREAM) h->be.reset(make_stream_backend(h->opts.flags, h->opts.block_threads));
  else if (plan.backend == PBD_BACKEND_JACOBI) h->be.reset(make_jacobi_backend(h->opts));
  else h->be.reset(make_tile_backend(h->opts, device));
  h->be->set_omega(h->params.omega);
  if (!h->be) { fail(PBD_ERR_UNSUPPORTED, "backend not available", status); return nullptr; }
  if ((ce = h->be->upload(plan, m, d)) != cudaSuccess) return bail(ce, "backend upload");
  if ((ce = cudaDeviceSynchronize()) != cudaSuccess) return bail(ce, "cudaDeviceSynchronize");
  h->uploadMs = wall_ms() - tUp;
  return h.release();
}

void pbd_destroy(pbd_handle* h) {
  if (!h) return;
  DeviceScope onDevice(h->device);
  cudaStreamSynchronize(h->stream);
  delete h;
}

const char* pbd_backend_name(const pbd_handle* h) { return h && h->be ? h->be->name() : "none"; }

static int enqueue_frames(pbd_handle* h, float dt, uint32_t frames) {
  DeviceScope onDevice(h->device);
  CU(onDevice.err);
  const StepConsts k = make_consts(h->params, dt);
  CU(cudaMemcpyAsync(h->d.consts, &k, sizeof(k), cudaMemcpyHostToDevice, h->stream));
  FrameShape f = h->shape();
  f.tetInert = k.alphaTet == 0.0f;
  CU(cudaEventRecord(h->ev0, h->stream));
  for (uint32_t i = 0; i < frames; ++i) CU(h->be->enqueue_frame(h->d, f, h->stream));
  CU(cudaEventRecord(h->ev1, h->stream));
  h->pending = true;
  return PBD_OK;
}

int pbd_step_async(pbd_handle* h, float dt, uint32_t frames) {
  if (!h) return fail(PBD_ERR_INVALID, "handle is null");
  return enqueue_frames(h, dt, frames);
}

int pbd_sync(pbd_handle* h, double* device_ms) {
  if (!h) return fail(PBD_ERR_INVALID, "handle is null");
  DeviceScope onDevice(h->device);
  CU(onDevice.err);
  CU(cudaStreamSynchronize(h->stream));
  if (device_ms) {
    float ms = 0.f;
    if (h->pending) CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    *device_ms = ms;
  }
  h->pending = false;
  static const bool trace = getenv("PBD_TILE_TRACE") != nullptr;
  if (trace) h->be->debug_dump();
  if (h->be->take_abort())
    return fail(PBD_ERR_CUDA, "a tile waited longer than PBD_SPIN_LIMIT_MS for another tile (was every rank of a sharded body "
                              "launched?); the frame ran to its end but the state is invalid");
  return PBD_OK;
}

uint64_t pbd_init_payload_size(uint32_t V, uint32_t E, uint32_t T, uint32_t pinnedCount) {
  return 64ull + 4ull * pinnedCount + 12ull * V + 8ull * E + 16ull * T;
}

// comm_loop's MSG_INIT decode (Server.cpp:30-70) with the bounds check the reference leaves out
pbd_handle* pbd_create_from_init(const void* payload, uint64_t size, int device, const pbd_options* opts, int* status) {
  if (status) *status = PBD_OK;
  if (!payload) { fail(PBD_ERR_INVALID, "payload is null", status); return nullptr; }
  if (size < 64) { fail(PBD_ERR_INVALID, "MSG_INIT payload shorter than its 64-byte fixed part", status); return nullptr; }
  const unsigned char* p = static_cast<const unsigned char*>(payload);
  uint32_t head[3], nPinned;
  pbd_params params;
  static_assert(sizeof(pbd_params) == 48, "pbd_params is the 48-byte wire block");
  std::memcpy(head, p, 12);
  std::memcpy(&params, p + 12, 48);
  std::memcpy(&nPinned, p + 60, 4);
  const uint32_t V = head[0], E = head[1], T = head[2];
  const uint64_t need = pbd_init_payload_size(V, E, T, nPinned);
  if (size < need) {
    fail(PBD_ERR_INVALID, "MSG_INIT payload of " + std::to_string(size) + " bytes, but V/E/T/pinnedCount need " + std::to_string(need), status);
    return nullptr;
  }
  // the arrays follow unaligned in general: copy them out, as the reference does
  std::vector<uint32_t> pinned(nPinned), edges((size_t)E * 2), tets((size_t)T * 4);
  std::vector<float> x0((size_t)V * 3);
  p += 64;
  if (nPinned) std::memcpy(pinned.data(), p, 4ull * nPinned);
  p += 4ull * nPinned;
  if (V) std::memcpy(x0.data(), p, 12ull * V);
  p += 12ull * V;
  if (E) std::memcpy(edges.data(), p, 8ull * E);
  p += 8ull * E;
  if (T) std::memcpy(tets.data(), p, 16ull * T);
  return pbd_create(&params, V, E, T, x0.data(), edges.data(), tets.data(), pinned.data(), nPinned, device, opts, status);
}

int pbd_step(pbd_handle* h, float dt, pbd_step_stats* stats) {
  if (!h) return fail(PBD_ERR_INVALID, "handle is null");
  const double t0 = wall_ms();
  int rc = enqueue_frames(h, dt, 1);
  if (rc != PBD_OK) return rc;
  double devMs = 0.0;
  rc = pbd_sync(h, &devMs);
  if (rc != PBD_OK) return rc;
  if (stats) {
    double p = 0, s = 0, c = 0;
    if (h->be->stage_ms(p, s, c)) { stats->predictMs += p; stats->solveMs += s; stats->commitMs += c; }
    else if (h->be->stage_share(p, c)) { stats->predictMs += devMs * p; stats->commitMs += devMs * c; stats->solveMs += devMs * (1.0 - p - c); }
    else stats->solveMs += devMs;
    stats->totalMs += wall_ms() - t0;
  }
  return PBD_OK;
}

int pbd_read_positions(pbd_handle* h, float* out, double* packMs) {
  if (!h || !out) return fail(PBD_ERR_INVALID, "null argument");
  const double t0 = wall_ms();
  DeviceScope onDevice(h->device);
  CU(onDevice.err);
  CU(launch_pack(h->d, h->stream));
  if (h->d.V) CU(cudaMemcpyAsync(out, h->d.packed, sizeof(float) * 3 * (size_t)h->d.V, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (packMs) *packMs += wall_ms() - t0;
  return PBD_OK;
}

int pbd_set_surface(pbd_handle* h, const uint32_t* tris, uint32_t nTris) {
  if (!h || (nTris && !tris)) return fail(PBD_ERR_INVALID, "null argument");
  const uint32_t V = h->d.V;
  for (size_t i = 0; i < (size_t)nTris * 3; ++i)
    if (tris[i] >= V) return fail(PBD_ERR_INDEX, "surface triangle index out of range (triangle " + std::to_string(i / 3) + ")");
  // BuildTriAdjacency, SoftBodySolver.cs:1173-1213: per vertex the incident triangles in ascending triangle order
  std::vector<uint32_t> off((size_t)V + 1, 0), adj((size_t)nTris * 3);
  for (size_t i = 0; i < (size_t)nTris * 3; ++i) off[tris[i] + 1]++;
  for (uint32_t v = 0; v < V; ++v) off[v + 1] += off[v];
  {
    std::vector<uint32_t> cur(off.begin(), off.end() - 1);
    for (uint32_t t = 0; t < nTris; ++t)
      for (int j = 0; j < 3; ++j) adj[cur[tris[3 * (size_t)t + j]]++] = t;
  }
  DeviceScope onDevice(h->device);
  CU(onDevice.err);
  CU(cudaStreamSynchronize(h->stream));
  cudaFree(h->surfTris); cudaFree(h->surfAdjOff); cudaFree(h->surfAdjTri);
  h->surfTris = h->surfAdjOff = h->surfAdjTri = nullptr;
  h->nSurfTris = 0;
  CU(cudaMalloc((void**)&h->surfTris, sizeof(uint32_t) * (adj.size() + 1)));
  CU(cudaMalloc((void**)&h->surfAdjOff, sizeof(uint32_t) * off.size()));
  CU(cudaMalloc((void**)&h->surfAdjTri, sizeof(uint32_t) * (adj.size() + 1)));
  if (nTris) CU(cudaMemcpy(h->surfTris, tris, sizeof(uint32_t) * 3 * (size_t)nTris, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(h->surfAdjOff, off.data(), sizeof(uint32_t) * off.size(), cudaMemcpyHostToDevice));
  if (nTris) CU(cudaMemcpy(h->surfAdjTri, adj.data(), sizeof(uint32_t) * adj.size(), cudaMemcpyHostToDevice));
  h->nSurfTris = nTris;
  return PBD_OK;
}

int pbd_read_normals(pbd_handle* h, float* out) {
  if (!h || !out) return fail(PBD_ERR_INVALID, "null argument");
  if (!h->surfAdjOff) return fail(PBD_ERR_INVALID, "pbd_set_surface first");
  DeviceScope onDevice(h->device);
  CU(onDevice.err);
  CU(launch_normals(h->d, h->surfTris, h->surfAdjOff, h->surfAdjTri, h->stream));
  if (h->d.V) CU(cudaMemcpyAsync(out, h->d.packed, sizeof(float) * 3 * (size_t)h->d.V, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return PBD_OK;
}

int pbd_set_params(pbd_handle* h, const pbd_params* p) {
  if (!h || !p) return fail(PBD_ERR_INVALID, "null argument");
  DeviceScope onDevice(h->device);
  CU(onDevice.err);
  CU(cudaStreamSynchronize(h->stream));
  h->params = *p;
  h->be->set_omega(p->omega);
  h->be->invalidate();
  return PBD_OK;
}

int pbd_set_colliders(pbd_handle* h, const pbd_collider* cols, uint32_t n, float particleRadius) {
  if (!h || (n && !cols)) return fail(PBD_ERR_INVALID, "null argument");
  if (n > PBD_MAX_COLLIDERS) return fail(PBD_ERR_INVALID, "at most PBD_MAX_COLLIDERS colliders");
  static_assert(sizeof(pbd_collider) == sizeof(Collider) && sizeof(pbd_collider) == 44, "pbd_collider layout");
  ColliderSet cs{};
  cs.n = n;
  cs.particleRadius = particleRadius;
  for (uint32_t i = 0; i < n; ++i) {
    if (cols[i].type > PBD_COLLIDER_CAPSULE) return fail(PBD_ERR_INVALID, "unknown collider type");
    const float* f = &cols[i].px;
    for (int j = 0; j < 10; ++j)
      if (!std::isfinite(f[j])) return fail(PBD_ERR_INVALID, "collider " + std::to_string(i) + " is not finite");
    std::memcpy(&cs.c[i], &cols[i], sizeof(Collider));
  }
  if (!std::isfinite(particleRadius)) return fail(PBD_ERR_INVALID, "particleRadius is not finite");
  DeviceScope onDevice(h->device);
  CU(onDevice.err);
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaMemcpy(h->d.colliders, &cs, sizeof(cs), cudaMemcpyHostToDevice));
  h->d.nColliders = n;
  h->be->invalidate();
  return PBD_OK;
}

int pbd_get_info(const pbd_handle* h, pbd_info* out) {
  if (!h || !out) return fail(PBD_ERR_INVALID, "null argument");
  base_info(h->plan, &h->params, *out);
  h->be->fill_info(*out);
  out->launches_per_frame = h->be->launches_per_frame(h->shape());
  out->device_bytes = h->d.bytes + h->be->device_bytes();
  out->upload_ms = h->uploadMs;
  return PBD_OK;
}

int pbd_get_schedule_order(const pbd_handle* h, uint32_t* eo, uint32_t* to) {
  if (!h) return fail(PBD_ERR_INVALID, "handle is null");
  if (eo && h->plan.E) std::memcpy(eo, h->plan.edgeOrder.data(), sizeof(uint32_t) * h->plan.E);
  if (to && h->plan.T) std::memcpy(to, h->plan.tetOrder.data(), sizeof(uint32_t) * h->plan.T);
  return PBD_OK;
}

static int plan_sequence(const Plan& p, uint32_t* items) {
  if (!items) return fail(PBD_ERR_INVALID, "items is null");
  if (p.backend != PBD_BACKEND_TILE) {
    // stream backend: all edges in schedule order, then all tets
    for (uint32_t k = 0; k < p.E; ++k) items[k] = k;
    for (uint32_t k = 0; k < p.T; ++k) items[p.E + k] = 0x80000000u | k;
    return PBD_OK;
  }
  // tile backend: phases -> tiles -> [edge groups, tet groups] (strict order yields the same list as above)
  size_t n = 0;
  for (const Phase& ph : p.phases)
    for (uint32_t t = ph.tileBegin; t < ph.tileBegin + ph.tileCount; ++t) {
      const Tile& tl = p.tiles[t];
      // a tet, then its riders (PBD_ORDER_RIDING): the tet's thread projects them right after it
      auto tet_unit = [&](uint32_t q) {
        items[n++] = 0x80000000u | q;
        if (tl.ride)
          for (uint32_t sl = 0; sl < 2; ++sl)
            if (p.tetRide[2 * (size_t)q + sl] != 0xffffffffu) items[n++] = p.tetRide[2 * (size_t)q + sl];
      };
      if (tl.mixed) {   // colour step s = edge group s, then tet group s (vertex-disjoint: any order inside a step is the same)
        for (uint32_t g = 0; g < tl.edgeGroupCount; ++g) {
          const Group& ge = p.groups[tl.edgeGroupBegin + g];
          const Group& gt = p.groups[tl.tetGroupBegin + g];
          for (uint32_t j = 0; j < ge.count; ++j) items[n++] = ge.begin + j;
          for (uint32_t j = 0; j < gt.count; ++j) tet_unit(gt.begin + j);
        }
        continue;
      }
      // the tile's free edges (its colour groups cover exactly those), then its tets
      uint32_t nFree = 0;
      for (uint32_t g = 0; g < tl.edgeGroupCount; ++g) nFree += p.groups[tl.edgeGroupBegin + g].count;
      if (!tl.ride) nFree = tl.edgeCount;
      for (uint32_t j = 0; j < nFree; ++j) items[n++] = tl.edgeBegin + j;
      for (uint32_t j = 0; j < tl.tetCount; ++j) tet_unit(tl.tetBegin + j);
    }
  return PBD_OK;
}

int pbd_get_schedule_sequence(const pbd_handle* h, uint32_t* items) {
  if (!h) return fail(PBD_ERR_INVALID, "handle is null");
  return plan_sequence(h->plan, items);
}

int pbd_get_array(pbd_handle* h, int what, float* out) {
  if (!h || !out) return fail(PBD_ERR_INVALID, "null argument");
  DeviceScope onDevice(h->device);
  CU(onDevice.err);
  CU(cudaStreamSynchronize(h->stream));
  const Plan& p = h->plan;
  const DeviceArrays& d = h->d;
  if (what == PBD_ARRAY_INV_MASS || what == PBD_ARRAY_XSTAR) {
    CU(h->be->export_pos(d, h->stream));
    CU(cudaStreamSynchronize(h->stream));
  }
  auto vec4 = [&](const float4* src, int comps, bool wOnly) -> int {
    std::vector<float4> tmp(d.V);
    if (d.V) CU(cudaMemcpy(tmp.data(), src, sizeof(float4) * d.V, cudaMemcpyDeviceToHost));
    for (uint32_t v = 0; v < d.V; ++v) {
      const float4 q = tmp[p.vertexToSlot[v]];
      if (wOnly) out[v] = q.w;
      else { out[(size_t)comps * v] = q.x; out[(size_t)comps * v + 1] = q.y; out[(size_t)comps * v + 2] = q.z; }
    }
    return PBD_OK;
  };
  auto scal = [&](const float* src, uint32_t n, uint32_t nDev, const std::vector<uint32_t>& order,
                  const std::vector<uint32_t>& dev) -> int {
    std::vector<float> tmp(nDev);
    if (nDev) CU(cudaMemcpy(tmp.data(), src, sizeof(float) * nDev, cudaMemcpyDeviceToHost));
    for (uint32_t k = 0; k < n; ++k) out[order[k]] = tmp[dev[k]];
    return PBD_OK;
  };
  switch (what) {
    case PBD_ARRAY_INV_MASS: return vec4(d.pos, 1, true);
    case PBD_ARRAY_XSTAR: return vec4(d.pos, 3, false);
    case PBD_ARRAY_VELOCITY: return vec4(d.vel, 3, false);
    case PBD_ARRAY_EDGE_REST: return scal(d.edgeRest, d.E, p.edgeDevCount, p.edgeOrder, p.edgeDev);
    case PBD_ARRAY_EDGE_LAMBDA: return scal(d.edgeLam, d.E, p.edgeDevCount, p.edgeOrder, p.edgeDev);
    case PBD_ARRAY_TET_REST: return scal(d.tetRest, d.T, p.tetDevCount, p.tetOrder, p.tetDev);
    case PBD_ARRAY_TET_LAMBDA: {
      const int rc = scal(d.tetLam, d.T, p.tetDevCount, p.tetOrder, p.tetDev);
      // relabelled tets (fast arithmetic, odd role permutation) carry their multiplier with the opposite sign
      if (rc == PBD_OK && !p.tetPerm.empty())
        for (uint32_t k = 0; k < d.T; ++k) if (tet_perm_is_odd(p.tetPerm[k])) out[p.tetOrder[k]] = -out[p.tetOrder[k]];
      return rc;
    }
    default: return fail(PBD_ERR_INVALID, "unknown array id");
  }
}

/* ---- one body across several GPUs ---- */

int pbd_shard_export(pbd_handle* h, void* out) {
  if (!h || !out) return fail(PBD_ERR_INVALID, "null argument");
  DeviceScope onDevice(h->device);
  CU(onDevice.err);
  CU(h->be->shard_export(out));
  return PBD_OK;
}

int pbd_shard_attach_ipc(pbd_handle* h, const void* all) {
  if (!h || !all) return fail(PBD_ERR_INVALID, "null argument");
  DeviceScope onDevice(h->device);
  CU(onDevice.err);
  CU(h->be->shard_attach_ipc(all));
  return PBD_OK;
}

int pbd_shard_attach_local(pbd_handle* const* hs, uint32_t world) {
  if (!hs || world == 0) return fail(PBD_ERR_INVALID, "null argument");
  for (uint32_t a = 0; a < world; ++a) {
    if (!hs[a] || hs[a]->be->shard_world() != world || hs[a]->be->shard_rank() != a)
      return fail(PBD_ERR_INVALID, "handles must be given in rank order, all created with shard_world = world");
  }
  for (uint32_t a = 0; a < world; ++a) {
    DeviceScope onDevice(hs[a]->device);
    CU(onDevice.err);
    for (uint32_t b = 0; b < world; ++b) {
      void *pos = nullptr, *done = nullptr;
      hs[b]->be->shard_local_pointers(&pos, &done);
      CU(hs[a]->be->shard_attach_pointers(b, pos, done, hs[b]->device));
    }
  }
  return PBD_OK;
}

int pbd_shard_owner(const pbd_handle* h, uint8_t* owner) {
  if (!h || !owner) return fail(PBD_ERR_INVALID, "null argument");
  std::vector<uint32_t> begin;
  h->be->shard_slot_ranges(begin);
  for (uint32_t v = 0; v < h->plan.V; ++v) {
    uint32_t r = 0;
    const uint32_t s = h->plan.vertexToSlot[v];
    while (r + 2 < begin.size() && s >= begin[r + 1]) ++r;
    owner[v] = (uint8_t)r;
  }
  return PBD_OK;
}

/* ---- schedule only (host) ---- */

pbd_plan* pbd_plan_create(uint32_t V, uint32_t E, uint32_t T, const float* x0, const uint32_t* edgeIds,
                          const uint32_t* tetIds, const pbd_options* opts, int* status) {
  if (status) *status = PBD_OK;
  MeshView m{V, E, T, x0, edgeIds, tetIds};
  std::string err;
  if (!validate_mesh(m, err)) {
    fail(err.find("out of range") != std::string::npos ? PBD_ERR_INDEX : PBD_ERR_INVALID, err, status);
    return nullptr;
  }
  std::unique_ptr<pbd_plan> p(new pbd_plan());
  p->opts = resolve_options(opts);
  // B200 defaults when no device is consulted: 148 SMs, 227 KB opt-in shared memory per CTA
  if (!build_plan(m, p->opts, 148, 227u * 1024u - kSmemReserve, p->plan, err)) {
    fail(PBD_ERR_INVALID, err, status);
    return nullptr;
  }
  return p.release();
}

int pbd_plan_get_info(const pbd_plan* p, pbd_info* out) {
  if (!p || !out) return fail(PBD_ERR_INVALID, "null argument");
  base_info(p->plan, nullptr, *out);
  return PBD_OK;
}

int pbd_plan_get_order(const pbd_plan* p, uint32_t* eo, uint32_t* to) {
  if (!p) return fail(PBD_ERR_INVALID, "plan is null");
  if (eo && p->plan.E) std::memcpy(eo, p->plan.edgeOrder.data(), sizeof(uint32_t) * p->plan.E);
  if (to && p->plan.T) std::memcpy(to, p->plan.tetOrder.data(), sizeof(uint32_t) * p->plan.T);
  return PBD_OK;
}

int pbd_plan_get_sequence(const pbd_plan* p, uint32_t* items) {
  if (!p) return fail(PBD_ERR_INVALID, "plan is null");
  return plan_sequence(p->plan, items);
}

static void copy_slots(const std::vector<uint32_t>& src, uint32_t* dst) {
  if (dst && !src.empty()) std::memcpy(dst, src.data(), sizeof(uint32_t) * src.size());
}

int pbd_plan_get_edge_slots(const pbd_plan* p, uint32_t* phase, uint32_t* tile, uint32_t* colour) {
  if (!p) return fail(PBD_ERR_INVALID, "plan is null");
  copy_slots(p->plan.edgePhase, phase); copy_slots(p->plan.edgeTile, tile); copy_slots(p->plan.edgeColor, colour);
  return PBD_OK;
}

int pbd_plan_get_tet_slots(const pbd_plan* p, uint32_t* phase, uint32_t* tile, uint32_t* colour) {
  if (!p) return fail(PBD_ERR_INVALID, "plan is null");
  copy_slots(p->plan.tetPhase, phase); copy_slots(p->plan.tetTile, tile); copy_slots(p->plan.tetColor, colour);
  return PBD_OK;
}

void pbd_plan_destroy(pbd_plan* p) { delete p; }

}  // extern "C"

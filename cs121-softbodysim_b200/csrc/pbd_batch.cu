// pbd_batch.cu -- many independent soft bodies (BASELINE.json config 4: 4096 x 5k-tet bodies).
//
// The reference runs one body per PBDServer process (CProgram/src/main.cpp:69-98); a batch is
// that path N times.  Bodies never interact, so nothing here synchronises across CTAs: ONE kernel
// per frame, each CTA takes whole bodies.  A body's record block (tile-local u16 indices, rest
// values, colour-group table, lambdas -- the same block format the tile backend uses, one "tile"
// = the whole body) is fetched with TMA bulk copies and its vertices are loaded ONCE; all
// substeps x iterations of the frame then run out of shared memory (pbd_sweep.cuh) with block
// barriers only, and positions / velocities / lambdas go back to HBM once per frame.
//
//   for body in my bodies:                                   SerialStepper::step, Sim.cpp:280-305
//     bulk-load record block; load xStar|w, x, v; predict     Sim.cpp:178-185
//     for substep: for iteration: [ground clamp of the previous iteration] edges; tets
//                  ground + commit (+ predict of the next substep) in shared memory      :187-222
//     store x, v, xStar; bulk-store lambdas
//
// Bodies must fit one SM's shared memory (16 B/vertex + 20 B/edge + 28 B/tet, ~213 KB for the
// 6k-tet body of config 4); larger bodies belong to pbd_create.  Order of projection per body:
// all edges colour by colour, then all tets colour by colour (PBD_ORDER_STRICT), disclosed by
// pbd_batch_get_schedule_order -- the unmodified reference run on the permuted arrays matches bit
// for bit (tests/test_parity_gpu.py).
#include <chrono>
#include <cstddef>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "pbd_body.h"
#include "pbd_device.cuh"
#include "pbd_sweep.cuh"

using namespace pbd;

namespace {

double wall_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

struct BodyDesc {
  unsigned long long blobOff;
  uint32_t staticBytes;
  uint32_t edgeDevBegin, edgeLamBytes;
  uint32_t tetDevBegin, tetLamBytes;
  uint32_t pad;
};

struct BatchParams {
  float4* pos;
  float4* prev;
  float4* vel;
  const unsigned char* blob;
  const BodyDesc* bodies;
  float* edgeLam;
  float* tetLam;
  const StepConsts* consts;
  const ColliderSet* colliders;   // (batches carry no colliders: always null / 0; the shared vertex stages read them)
  uint32_t nColliders;
  uint32_t nBodies, substeps, iterations;
  uint32_t recStride;
  // work list (see BatchPiece): CTA c runs pieces [pieceBegin[c], pieceBegin[c + 1])
  const struct BatchPiece* pieces;
  const uint32_t* pieceBegin;
  unsigned* pieceDone;      // per body: frame stamp set when the HEAD piece of a split body has been stored
  uint32_t frameStamp;      // this frame's stamp (counts frames, wrap-safe compare)
  long long spinLimit;      // bounded wait (cycles) of a tail piece for its head piece
  unsigned* abortHost;      // mapped host word set when that wait gives up
  uint32_t* stageHost;      // mapped host memory, 4 words per CTA: cycles of thread 0 in (predict loads, fused
                            // commit+predict stages, final commit stages, the whole frame) -> pbd_step_stats
};

// One unit of work of a CTA: substeps [subBegin, subEnd) of one body's frame.  Most pieces are whole
// frames.  A frame is a sequential chain of `substeps` units, so dealing WHOLE bodies to the CTAs
// quantises the kernel time to ceil(bodies / CTAs) body-frames (512 bodies on 148 SMs: 4 instead of
// 3.46, i.e. 13.5 % idle -- what capped the strong scaling at 8 GPUs).  Instead the body-major
// sequence of (body, substep) units is cut into one equal share per CTA; a body that straddles a cut
// is stepped by two CTAs: the first runs its head substeps FIRST, stores the state (positions,
// velocities, lambdas: exactly what a frame end stores) and stamps pieceDone[body]; the second runs its
// tail substeps LAST, after polling the stamp -- by then the head has long been done, so nobody waits.
// Results are bit-identical to stepping the frame in one piece (the state at a substep boundary is the
// state the next frame would start from).
struct BatchPiece {
  uint32_t body, subBegin, subEnd;
  uint32_t flags;   // bit 0: wait for pieceDone[body] before loading, bit 1: set it after storing
};

// LANES: threads cooperating on one tet (1, 2, 4: bit-identical results); FAST: PBD_FLAG_FAST_ARITH forms
template <int LANES, bool FAST>
__global__ void __launch_bounds__(512, 1) batch_frame_kernel(const BatchParams P) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long mbar;
  const uint32_t svOff = P.recStride;
  float4* const sv = reinterpret_cast<float4*>(smem + svOff);
  const StepConsts k = *P.consts;
  const uint32_t tid = threadIdx.x, nth = blockDim.x;
  const bool clamp = P.iterations > 0;   // the reference clamps once per iteration (Sim.cpp:296)
  if (tid == 0) {
    mbar_init(&mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
  }
  __syncthreads();
  uint32_t parity = 0;
  __shared__ uint32_t stageS[4];   // per-stage cycle accounting (thread 0), see BatchParams::stageHost
  if (tid == 0) { stageS[0] = stageS[1] = stageS[2] = 0u; stageS[3] = (uint32_t)clock64(); }
  for (uint32_t pi = P.pieceBegin[blockIdx.x]; pi < P.pieceBegin[blockIdx.x + 1]; ++pi) {
    const BatchPiece piece = P.pieces[pi];
    const uint32_t b = piece.body;
    const BodyDesc c = P.bodies[b];
    if (piece.flags & 1u) {   // tail piece: the head piece (another CTA) must have stored the body's state
      if (tid == 0) {
        uint32_t polls = 0;
        long long t0 = 0;
        while ((int)(ld_acquire(P.pieceDone + b) - P.frameStamp) < 0) {
          if ((++polls & 1023u) == 0u) {
            if (polls == 1024u) t0 = clock64();
            else if (clock64() - t0 > P.spinLimit) { *reinterpret_cast<volatile unsigned*>(P.abortHost) = 1u; __threadfence_system(); break; }
          }
        }
      }
      __syncthreads();
    }
    if (tid == 0) {
      mbar_expect_tx(&mbar, c.staticBytes + c.edgeLamBytes + c.tetLamBytes);
      bulk_load(smem, P.blob + c.blobOff, c.staticBytes, &mbar);
      if (c.edgeLamBytes) bulk_load(smem + c.staticBytes, P.edgeLam + c.edgeDevBegin, c.edgeLamBytes, &mbar);
      if (c.tetLamBytes) bulk_load(smem + c.staticBytes + c.edgeLamBytes, P.tetLam + c.tetDevBegin, c.tetLamBytes, &mbar);
    }
    while (!mbar_try_wait(&mbar, parity)) {}
    parity ^= 1u;
    const TileHdr h = *reinterpret_cast<const TileHdr*>(smem);
    // predict of the piece's first substep while the vertices are loaded
    if (tid == 0) stageS[0] -= (uint32_t)clock64();
    for (uint32_t i = tid; i < h.vertCount; i += nth) sv[i] = load_transform(P, k, h.vertBegin + i, LOAD_PREDICT, false);
    __syncthreads();
    if (tid == 0) stageS[0] += (uint32_t)clock64();
    for (uint32_t sub = piece.subBegin; sub < piece.subEnd; ++sub) {
      for (uint32_t it = 0; it < P.iterations; ++it) {
        if (it != 0) {   // ground clamp that closes the previous iteration
          for (uint32_t i = tid; i < h.vertCount; i += nth) { float4 p = sv[i]; ground_vertex(p, k); sv[i] = p; }
          __syncthreads();
        }
        sweep_edges<FAST>(h, 0, svOff, k.alphaEdge, nullptr);
        // (fast arithmetic, alpha == 0: the tet multipliers never enter a correction and are not accumulated, as in the tile kernel)
        sweep_tets<LANES, FAST>(h, 0, svOff, k.alphaTet, nullptr, 0.0f, !(FAST && k.alphaTet == 0.0f));
      }
      const bool last = sub + 1 == piece.subEnd;
      if (tid == 0) stageS[last ? 2 : 1] -= (uint32_t)clock64();
      for (uint32_t i = tid; i < h.vertCount; i += nth) {
        const uint32_t s = h.vertBegin + i;
        float4 p = sv[i], x = __ldcg(P.prev + s), v;
        if (clamp) ground_vertex(p, k);
        commit_vertex(p, x, v, k);
        v.w = 0.0f;
        if (last) {
          __stcg(P.pos + s, p);
        } else {
          p = predict_vertex(x, v, p.w, k);
          sv[i] = p;
        }
        __stcg(P.prev + s, x);
        __stcg(P.vel + s, v);
      }
      __syncthreads();
      if (tid == 0) stageS[last ? 2 : 1] += (uint32_t)clock64();
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      if (c.edgeLamBytes) bulk_store(P.edgeLam + c.edgeDevBegin, smem + h.offEdgeLam, c.edgeLamBytes);
      if (c.tetLamBytes) bulk_store(P.tetLam + c.tetDevBegin, smem + h.offTetLam, c.tetLamBytes);
      bulk_commit();
      if (piece.flags & 2u) {   // head piece of a split body: everything it stored is visible before the stamp
        bulk_wait_all();
        asm volatile("fence.proxy.async;" ::: "memory");
        st_release(P.pieceDone + b, P.frameStamp);
      } else {
        bulk_wait_read();       // the buffer is overwritten by the next body
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    bulk_wait_all();
    if (P.stageHost) {
      uint32_t* o = P.stageHost + 4u * blockIdx.x;
      o[0] = stageS[0]; o[1] = stageS[1]; o[2] = stageS[2]; o[3] = (uint32_t)clock64() - stageS[3];
    }
  }
}

template <class T>
cudaError_t upload_vec(T** dst, const std::vector<T>& src, uint64_t& bytes) {
  cudaError_t e = cudaMalloc((void**)dst, sizeof(T) * src.size() + 256);
  if (e != cudaSuccess) return e;
  bytes += sizeof(T) * src.size();
  if (!src.empty()) e = cudaMemcpy(*dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice);
  return e;
}

}  // namespace

struct pbd_batch {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  pbd_params params{};
  uint32_t nBodies = 0;
  uint64_t Vtot = 0, Etot = 0, Ttot = 0;
  DeviceArrays d;                     // pos/prev/vel/packed/consts over the concatenated vertices
  unsigned char* blob = nullptr;
  BodyDesc* bodies = nullptr;
  uint32_t recStride = 128, grid = 1, maxVerts = 0;
  size_t smemBytes = 0;
  uint32_t edgeColorsMax = 0, tetColorsMax = 0;
  uint64_t gatherWavefronts[2] = {0, 0}, gatherIdeal[2] = {0, 0};
  std::vector<uint32_t> edgeOrder, tetOrder;   // per body, body-local constraint indices, concatenated
  uint64_t bytes = 0;
  double planMs = 0.0, uploadMs = 0.0;
  bool pending = false;
  uint32_t lanes = 1;
  bool fast = false;
  BatchPiece* pieces = nullptr;        // the CTAs' work lists (see BatchPiece)
  uint32_t* pieceBegin = nullptr;
  unsigned* pieceDone = nullptr;
  uint32_t frameStamp = 0, nSplit = 0;
  unsigned* abortHost = nullptr;       // mapped host word + its device alias (bounded wait of a tail piece)
  unsigned* abortHostDev = nullptr;
  uint32_t* stageHost = nullptr;       // mapped host memory, 4 words per CTA (per-stage cycle accounting)
  uint32_t* stageHostDev = nullptr;
  long long spinLimit = 0;
  const void* kernel() const {
    if (fast) return (const void*)batch_frame_kernel<1, true>;
    return lanes == 2 ? (const void*)batch_frame_kernel<2, false>
           : lanes == 4 ? (const void*)batch_frame_kernel<4, false> : (const void*)batch_frame_kernel<1, false>;
  }

  ~pbd_batch() {
    cudaFree(d.pos); cudaFree(d.prev); cudaFree(d.vel); cudaFree(d.edgeLam); cudaFree(d.tetLam);
    cudaFree(d.packed); cudaFree(d.consts); cudaFree(d.slotOf); cudaFree(blob); cudaFree(bodies);
    cudaFree(pieces); cudaFree(pieceBegin); cudaFree(pieceDone);
    if (abortHost) cudaFreeHost(abortHost);
    if (stageHost) cudaFreeHost(stageHost);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (stream) cudaStreamDestroy(stream);
  }
};

namespace {

int bfail(int code, const std::string& msg, int* status = nullptr) {
  set_last_error(msg);
  if (status) *status = code;
  return code;
}

#define BCU(call)                                                                             \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess) return bfail(PBD_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
  } while (0)

struct Coloured {
  std::vector<uint32_t> eOrder, eCounts, tOrder, tCounts;
  std::vector<uint32_t> newLocal;   // body vertex -> its index in the body's shared-memory array / slot range (pbd_placement.cpp)
  std::vector<uint8_t> tetPerm;     // fast arithmetic: role permutation of every tet, parallel to tOrder (empty: none)
  PlaceStats gathers;
};

}  // namespace

extern "C" {

pbd_batch* pbd_batch_create(const pbd_params* params, uint32_t nBodies, const uint64_t* vOff, const uint64_t* eOff,
                            const uint64_t* tOff, const float* x0, const uint32_t* edgeIds, const uint32_t* tetIds,
                            int device, const pbd_options* opts, int* status) {
  if (status) *status = PBD_OK;
  if (!params || !vOff || !eOff || !tOff) { bfail(PBD_ERR_INVALID, "null argument", status); return nullptr; }
  const uint64_t Vtot = vOff[nBodies], Etot = eOff[nBodies], Ttot = tOff[nBodies];
  if (Vtot > 0xfffffff0ull || Etot > 0xfffffff0ull || Ttot > 0xfffffff0ull) { bfail(PBD_ERR_INVALID, "batch too large", status); return nullptr; }
  if ((Vtot && !x0) || (Etot && !edgeIds) || (Ttot && !tetIds)) { bfail(PBD_ERR_INVALID, "null array", status); return nullptr; }
  for (uint32_t b = 0; b < nBodies; ++b) {
    if (vOff[b + 1] < vOff[b] || eOff[b + 1] < eOff[b] || tOff[b + 1] < tOff[b]) { bfail(PBD_ERR_INVALID, "offsets must be non-decreasing", status); return nullptr; }
    const uint64_t Vb = vOff[b + 1] - vOff[b];
    if (Vb > 65535) { bfail(PBD_ERR_UNSUPPORTED, "a batch body has more than 65535 vertices: use pbd_create for it", status); return nullptr; }
    for (uint64_t i = 3 * vOff[b]; i < 3 * vOff[b + 1]; ++i)
      if (!std::isfinite(x0[i])) { bfail(PBD_ERR_INVALID, "x0 is not finite in body " + std::to_string(b), status); return nullptr; }
    for (uint64_t i = 2 * eOff[b]; i < 2 * eOff[b + 1]; ++i)
      if (edgeIds[i] >= Vb) { bfail(PBD_ERR_INDEX, "edge index out of range in body " + std::to_string(b), status); return nullptr; }
    for (uint64_t i = 4 * tOff[b]; i < 4 * tOff[b + 1]; ++i)
      if (tetIds[i] >= Vb) { bfail(PBD_ERR_INDEX, "tet index out of range in body " + std::to_string(b), status); return nullptr; }
  }
  int nDev = 0;
  if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev == 0) {
    cudaGetLastError();
    bfail(PBD_ERR_NO_DEVICE, "no CUDA device (this library has no CPU fallback)", status);
    return nullptr;
  }
  if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
  if (device >= nDev) { bfail(PBD_ERR_INVALID, "device ordinal out of range", status); return nullptr; }
  cudaError_t ce;
  auto cbail = [&](cudaError_t e, const char* what) {
    bfail(e == cudaErrorMemoryAllocation ? PBD_ERR_OOM : PBD_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e), status);
    return nullptr;
  };
  DeviceScope onDevice(device);
  if ((ce = onDevice.err) != cudaSuccess) return cbail(ce, "cudaSetDevice");
  cudaDeviceProp prop{};
  if ((ce = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return cbail(ce, "cudaGetDeviceProperties");
  const size_t smemLimit = prop.sharedMemPerBlockOptin > 1024 ? prop.sharedMemPerBlockOptin - 1024 : 0;

  std::unique_ptr<pbd_batch> B(new pbd_batch());
  B->device = device;
  B->params = *params;
  B->nBodies = nBodies;
  B->Vtot = Vtot; B->Etot = Etot; B->Ttot = Ttot;
  // options honoured by the batch backend: PBD_FLAG_FAST_ARITH and lanes_per_tet (struct_size-guarded)
  if (opts) {
    const uint32_t sz = opts->struct_size ? opts->struct_size : (uint32_t)sizeof(pbd_options);
    if (sz >= offsetof(pbd_options, flags) + 4u) B->fast = (opts->flags & PBD_FLAG_FAST_ARITH) != 0u;
    if (sz >= offsetof(pbd_options, lanes_per_tet) + 4u && !B->fast && (opts->lanes_per_tet == 2 || opts->lanes_per_tet == 4)) B->lanes = opts->lanes_per_tet;
  }

  // ---- host: per-body derived state (reference init helpers, caller's order) + colouring
  const double tPlan = wall_ms();
  std::vector<float4> pos(Vtot), prev(Vtot);
  std::vector<uint32_t> slotOf(Vtot);   // caller vertex (concatenated) -> slot: a body's vertices are renumbered inside its range
  std::vector<unsigned char> blob;
  std::vector<BodyDesc> descs(nBodies);
  B->edgeOrder.resize(Etot);
  B->tetOrder.resize(Ttot);
  std::map<std::string, std::shared_ptr<Coloured>> cache;   // bodies that share a topology share a colouring
  uint32_t eDev = 0, tDev = 0, recMax = 64;
  for (uint32_t b = 0; b < nBodies; ++b) {
    const uint32_t Vb = (uint32_t)(vOff[b + 1] - vOff[b]), Eb = (uint32_t)(eOff[b + 1] - eOff[b]), Tb = (uint32_t)(tOff[b + 1] - tOff[b]);
    MeshView m{Vb, Eb, Tb, x0 + 3 * vOff[b], edgeIds + 2 * eOff[b], tetIds + 4 * tOff[b]};
    std::vector<float> w, eRest, tRest;
    host_inverse_mass(m, nullptr, 0, w);
    host_rest_state(m, eRest, tRest);
    std::string key(reinterpret_cast<const char*>(&Vb), 4);
    key.append(reinterpret_cast<const char*>(m.edges), sizeof(uint32_t) * 2 * (size_t)Eb);
    key.append(reinterpret_cast<const char*>(m.tets), sizeof(uint32_t) * 4 * (size_t)Tb);
    std::shared_ptr<Coloured>& col = cache[key];
    if (!col) {
      col = std::make_shared<Coloured>();
      colour_and_order(m.edges, Eb, 2, Vb, 512, col->eOrder, col->eCounts);   // 512 = the kernel's block size
      colour_and_order(m.tets, Tb, 4, Vb, 512 / B->lanes, col->tOrder, col->tCounts);
      // where the body's vertices sit in shared memory and which constraints of a colour group share a quarter-warp:
      // the placement search of the tile backend on the whole body (one thread per constraint only).  A body lives in
      // shared memory for the whole frame, so every sweep of the frame gathers through these bank groups.
      {
        std::vector<PlaceGroup> groups;
        std::vector<uint32_t> loc, payload;
        for (int ty = 0; ty < 2; ++ty) {
          const std::vector<uint32_t>& ord = ty ? col->tOrder : col->eOrder;
          const std::vector<uint32_t>& cnt = ty ? col->tCounts : col->eCounts;
          const uint32_t ar = ty ? 4u : 2u;
          const uint32_t* ids = ty ? m.tets : m.edges;
          uint32_t at = 0;
          for (uint32_t c : cnt) {
            groups.push_back({(uint32_t)payload.size(), c, ar});
            for (uint32_t q = 0; q < c; ++q, ++at) {
              payload.push_back(ord[at]);
              for (uint32_t r = 0; r < 4; ++r) loc.push_back(r < ar ? ids[(size_t)ar * ord[at] + r] : 0xffffffffu);
            }
          }
        }
        const int effort = (B->lanes == 1 && !getenv("PBD_BATCH_NOPLACE")) ? 1 : 0;
        std::vector<uint8_t> perm(B->fast && effort && !getenv("PBD_PLAN_NORELABEL") ? payload.size() : 0);
        optimise_placement(Vb, groups.data(), (uint32_t)groups.size(), loc.data(), payload.data(), (uint32_t)payload.size(), effort, 0u,
                           col->newLocal, &col->gathers, perm.empty() ? nullptr : perm.data());
        if (!perm.empty()) col->tetPerm.assign(perm.begin() + Eb, perm.end());
        std::copy(payload.begin(), payload.begin() + Eb, col->eOrder.begin());
        std::copy(payload.begin() + Eb, payload.end(), col->tOrder.begin());
      }
    }
    const std::vector<uint32_t>& nl = col->newLocal;
    for (int ty = 0; ty < 2; ++ty) { B->gatherWavefronts[ty] += col->gathers.wavefronts[ty]; B->gatherIdeal[ty] += col->gathers.ideal[ty]; }
    std::copy(col->eOrder.begin(), col->eOrder.end(), B->edgeOrder.begin() + eOff[b]);
    std::copy(col->tOrder.begin(), col->tOrder.end(), B->tetOrder.begin() + tOff[b]);
    for (uint32_t v = 0; v < Vb; ++v) {
      const float* q = m.x0 + 3 * (size_t)v;
      pos[vOff[b] + nl[v]] = make_float4(q[0], q[1], q[2], w[v]);
      prev[vOff[b] + nl[v]] = make_float4(q[0], q[1], q[2], 0.0f);
      slotOf[vOff[b] + v] = (uint32_t)(vOff[b] + nl[v]);
    }
    const uint32_t nEG = (uint32_t)col->eCounts.size(), nTG = (uint32_t)col->tCounts.size();
    // the sweeps project a colour group in ONE pass of the 512-thread block: a larger group would lose constraints
    for (uint32_t cnt : col->eCounts) if (cnt > 512u) { bfail(PBD_ERR_INVALID, "internal: edge colour group exceeds the block", status); return nullptr; }
    for (uint32_t cnt : col->tCounts) if (cnt * B->lanes > 512u) { bfail(PBD_ERR_INVALID, "internal: tet colour group exceeds the block", status); return nullptr; }
    B->edgeColorsMax = std::max(B->edgeColorsMax, nEG);
    B->tetColorsMax = std::max(B->tetColorsMax, nTG);
    B->maxVerts = std::max(B->maxVerts, Vb);

    TileHdr h{};
    h.vertCount = Vb; h.flags = 1; h.vertBegin = (uint32_t)vOff[b];
    h.nEdgeGroups = nEG; h.nTetGroups = nTG; h.nEdges = Eb; h.nTets = Tb;
    uint32_t off = 64 + 4u * kMaxPreds;
    h.offVertIdx = off;
    h.offEdgeGroups = off; off += 8u * (pad4(nEG * 2) / 2);
    h.offTetGroups = off; off += 8u * (pad4(nTG * 2) / 2);
    h.offEdgeIdx = off; off += 4u * pad4(Eb);
    h.offEdgeRest = off; off += 4u * pad4(Eb);
    h.offTetIdx = off; off += 8u * (pad4(Tb * 2) / 2);
    h.offTetRest = off; off += 4u * pad4(Tb);
    const uint32_t staticBytes = off;
    h.offEdgeLam = off; off += 4u * pad4(Eb);
    h.offTetLam = off; off += 4u * pad4(Tb);
    recMax = std::max(recMax, off);
    if ((size_t)((off + 127u) & ~127u) + 16ull * Vb > smemLimit) {
      bfail(PBD_ERR_UNSUPPORTED, "body " + std::to_string(b) + " does not fit one SM's shared memory: use pbd_create for it", status);
      return nullptr;
    }
    const size_t base = blob.size();
    blob.resize(base + staticBytes, 0);
    unsigned char* p = blob.data() + base;
    memcpy(p, &h, sizeof(h));
    uint32_t* eg = reinterpret_cast<uint32_t*>(p + h.offEdgeGroups);
    for (uint32_t g = 0, at = 0; g < nEG; ++g) { eg[2 * g] = at; eg[2 * g + 1] = col->eCounts[g]; at += col->eCounts[g]; }
    uint32_t* tg = reinterpret_cast<uint32_t*>(p + h.offTetGroups);
    for (uint32_t g = 0, at = 0; g < nTG; ++g) { tg[2 * g] = at; tg[2 * g + 1] = col->tCounts[g]; at += col->tCounts[g]; }
    uint32_t* ei = reinterpret_cast<uint32_t*>(p + h.offEdgeIdx);
    float* er = reinterpret_cast<float*>(p + h.offEdgeRest);
    for (uint32_t q = 0; q < Eb; ++q) {
      const uint32_t e = col->eOrder[q];
      ei[q] = nl[m.edges[2 * (size_t)e]] | (nl[m.edges[2 * (size_t)e + 1]] << 16);
      er[q] = eRest[e];
    }
    uint32_t* ti = reinterpret_cast<uint32_t*>(p + h.offTetIdx);
    float* tr = reinterpret_cast<float*>(p + h.offTetRest);
    for (uint32_t q = 0; q < Tb; ++q) {
      const uint32_t* id = m.tets + 4 * (size_t)col->tOrder[q];
      // (fast arithmetic: the placement search may have given the tet's vertices other roles; an odd permutation negates
      // the signed volume, so the record carries the negated rest volume)
      const uint8_t code = col->tetPerm.empty() ? (uint8_t)0xE4 : col->tetPerm[q];
      ti[2 * q] = nl[id[code & 3u]] | (nl[id[(code >> 2) & 3u]] << 16);
      ti[2 * q + 1] = nl[id[(code >> 4) & 3u]] | (nl[id[(code >> 6) & 3u]] << 16);
      tr[q] = tet_perm_is_odd(code) ? -tRest[col->tOrder[q]] : tRest[col->tOrder[q]];
    }
    BodyDesc& c = descs[b];
    c.blobOff = base; c.staticBytes = staticBytes;
    c.edgeDevBegin = eDev; c.edgeLamBytes = 4u * pad4(Eb); eDev += pad4(Eb);
    c.tetDevBegin = tDev; c.tetLamBytes = 4u * pad4(Tb); tDev += pad4(Tb);
    c.pad = 0;
  }
  B->planMs = wall_ms() - tPlan;
  B->recStride = (recMax + 127u) & ~127u;
  B->smemBytes = (size_t)B->recStride + 16ull * std::max(B->maxVerts, 1u);

  // ---- device
  const double tUp = wall_ms();
  DeviceArrays& d = B->d;
  d.V = (uint32_t)Vtot; d.E = (uint32_t)Etot; d.T = (uint32_t)Ttot;
  if ((ce = cudaStreamCreateWithFlags(&B->stream, cudaStreamNonBlocking)) != cudaSuccess) return cbail(ce, "cudaStreamCreate");
  if ((ce = cudaEventCreate(&B->ev0)) != cudaSuccess || (ce = cudaEventCreate(&B->ev1)) != cudaSuccess) return cbail(ce, "cudaEventCreate");
  if ((ce = upload_vec(&d.pos, pos, B->bytes)) != cudaSuccess) return cbail(ce, "upload pos");
  if ((ce = upload_vec(&d.prev, prev, B->bytes)) != cudaSuccess) return cbail(ce, "upload prev");
  if ((ce = upload_vec(&d.slotOf, slotOf, B->bytes)) != cudaSuccess) return cbail(ce, "upload slotOf");
  if ((ce = cudaMalloc((void**)&d.vel, sizeof(float4) * (Vtot + 1))) != cudaSuccess) return cbail(ce, "cudaMalloc vel");
  if ((ce = cudaMemset(d.vel, 0, sizeof(float4) * (Vtot + 1))) != cudaSuccess) return cbail(ce, "memset vel");
  if ((ce = cudaMalloc((void**)&d.edgeLam, sizeof(float) * ((size_t)eDev + 4))) != cudaSuccess) return cbail(ce, "cudaMalloc edgeLam");
  if ((ce = cudaMalloc((void**)&d.tetLam, sizeof(float) * ((size_t)tDev + 4))) != cudaSuccess) return cbail(ce, "cudaMalloc tetLam");
  if ((ce = cudaMemset(d.edgeLam, 0, sizeof(float) * ((size_t)eDev + 4))) != cudaSuccess) return cbail(ce, "memset edgeLam");
  if ((ce = cudaMemset(d.tetLam, 0, sizeof(float) * ((size_t)tDev + 4))) != cudaSuccess) return cbail(ce, "memset tetLam");
  if ((ce = cudaMalloc((void**)&d.packed, sizeof(float) * (3 * Vtot + 1))) != cudaSuccess) return cbail(ce, "cudaMalloc packed");
  if ((ce = cudaMalloc((void**)&d.consts, sizeof(StepConsts))) != cudaSuccess) return cbail(ce, "cudaMalloc consts");
  B->bytes += sizeof(float4) * Vtot + sizeof(float) * ((size_t)eDev + tDev + 3 * Vtot);
  if ((ce = upload_vec(&B->blob, blob, B->bytes)) != cudaSuccess) return cbail(ce, "upload blob");
  if ((ce = upload_vec(&B->bodies, descs, B->bytes)) != cudaSuccess) return cbail(ce, "upload bodies");
  if ((ce = cudaFuncSetAttribute(B->kernel(), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B->smemBytes)) != cudaSuccess) return cbail(ce, "cudaFuncSetAttribute");
  int perSM = 0;
  if ((ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, B->kernel(), 512, B->smemBytes)) != cudaSuccess) return cbail(ce, "occupancy");
  if (perSM < 1) { bfail(PBD_ERR_UNSUPPORTED, "batch kernel does not fit on an SM", status); return nullptr; }
  B->grid = std::max(1u, std::min(nBodies, (uint32_t)(perSM * prop.multiProcessorCount)));
  {
    // work lists: the body-major sequence of (body, substep) units cut into one equal share per CTA (see BatchPiece)
    const uint32_t S = params->substeps > 1u ? params->substeps : 1u, G = B->grid;
    const uint64_t U = (uint64_t)nBodies * S;
    std::vector<BatchPiece> pieces;
    std::vector<uint32_t> begin(G + 1, 0);
    const bool split = getenv("PBD_BATCH_NOSPLIT") == nullptr;   // debug switch: whole bodies, dealt round-robin
    for (uint32_t c = 0; c < G; ++c) {
      begin[c] = (uint32_t)pieces.size();
      if (!split) {
        for (uint32_t b = c; b < nBodies; b += G) pieces.push_back({b, 0u, S, 0u});
        continue;
      }
      const uint64_t u0 = U * c / G, u1 = U * (c + 1) / G;
      const uint32_t bFirst = (uint32_t)(u0 / S), sFirst = (uint32_t)(u0 % S);     // body the share starts in, at substep sFirst
      const uint32_t bLast = (uint32_t)(u1 / S), sLast = (uint32_t)(u1 % S);       // body it ends in, before substep sLast
      if (sLast != 0 && bLast < nBodies) { pieces.push_back({bLast, 0u, sLast, 2u}); ++B->nSplit; }   // head of a split body: first
      for (uint32_t b = bFirst + (sFirst ? 1u : 0u); b < bLast; ++b) pieces.push_back({b, 0u, S, 0u});
      if (sFirst != 0) pieces.push_back({bFirst, sFirst, S, 1u});                                      // tail of a split body: last
    }
    begin[G] = (uint32_t)pieces.size();
    if ((ce = upload_vec(&B->pieces, pieces, B->bytes)) != cudaSuccess) return cbail(ce, "upload pieces");
    if ((ce = upload_vec(&B->pieceBegin, begin, B->bytes)) != cudaSuccess) return cbail(ce, "upload pieceBegin");
    if ((ce = cudaMalloc((void**)&B->pieceDone, sizeof(unsigned) * ((size_t)nBodies + 1))) != cudaSuccess) return cbail(ce, "cudaMalloc pieceDone");
    if ((ce = cudaMemset(B->pieceDone, 0, sizeof(unsigned) * ((size_t)nBodies + 1))) != cudaSuccess) return cbail(ce, "memset pieceDone");
    if ((ce = cudaHostAlloc((void**)&B->abortHost, 64, cudaHostAllocMapped)) != cudaSuccess) return cbail(ce, "cudaHostAlloc");
    *B->abortHost = 0u;
    if ((ce = cudaHostGetDevicePointer((void**)&B->abortHostDev, B->abortHost, 0)) != cudaSuccess) return cbail(ce, "cudaHostGetDevicePointer");
    if ((ce = cudaHostAlloc((void**)&B->stageHost, sizeof(uint32_t) * 4 * (size_t)std::max(1u, B->grid), cudaHostAllocMapped)) != cudaSuccess) return cbail(ce, "cudaHostAlloc");
    memset(B->stageHost, 0, sizeof(uint32_t) * 4 * (size_t)std::max(1u, B->grid));
    if ((ce = cudaHostGetDevicePointer((void**)&B->stageHostDev, B->stageHost, 0)) != cudaSuccess) return cbail(ce, "cudaHostGetDevicePointer");
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
    const char* e = getenv("PBD_SPIN_LIMIT_MS");
    B->spinLimit = (long long)((e ? atof(e) : 10000.0) * (double)std::max(khz, 1000000));
  }
  if ((ce = cudaDeviceSynchronize()) != cudaSuccess) return cbail(ce, "cudaDeviceSynchronize");
  B->uploadMs = wall_ms() - tUp;
  return B.release();
}

int pbd_batch_step_async(pbd_batch* b, float dt, uint32_t frames) {
  if (!b) return bfail(PBD_ERR_INVALID, "batch is null");
  DeviceScope onDevice(b->device);
  BCU(onDevice.err);
  const StepConsts k = make_consts(b->params, dt);
  BCU(cudaMemcpyAsync(b->d.consts, &k, sizeof(k), cudaMemcpyHostToDevice, b->stream));
  BatchParams P{};
  P.pos = b->d.pos; P.prev = b->d.prev; P.vel = b->d.vel; P.blob = b->blob; P.bodies = b->bodies;
  P.edgeLam = b->d.edgeLam; P.tetLam = b->d.tetLam; P.consts = b->d.consts;
  P.nBodies = b->nBodies; P.substeps = b->params.substeps > 1u ? b->params.substeps : 1u; P.iterations = b->params.iterations;
  P.recStride = b->recStride;
  P.pieces = b->pieces; P.pieceBegin = b->pieceBegin; P.pieceDone = b->pieceDone;
  P.spinLimit = b->spinLimit; P.abortHost = b->abortHostDev; P.stageHost = b->stageHostDev;
  BCU(cudaEventRecord(b->ev0, b->stream));
  for (uint32_t f = 0; f < frames; ++f) {
    P.frameStamp = ++b->frameStamp;
    if (b->nBodies) {
      void* args[] = {&P};
      BCU(cudaLaunchKernel(b->kernel(), dim3(b->grid), dim3(512), args, b->smemBytes, b->stream));
    }
  }
  BCU(cudaEventRecord(b->ev1, b->stream));
  b->pending = true;
  return PBD_OK;
}

int pbd_batch_sync(pbd_batch* b, double* device_ms) {
  if (!b) return bfail(PBD_ERR_INVALID, "batch is null");
  DeviceScope onDevice(b->device);
  BCU(onDevice.err);
  BCU(cudaStreamSynchronize(b->stream));
  if (device_ms) {
    float ms = 0.f;
    if (b->pending) BCU(cudaEventElapsedTime(&ms, b->ev0, b->ev1));
    *device_ms = ms;
  }
  b->pending = false;
  if (b->abortHost && *reinterpret_cast<volatile unsigned*>(b->abortHost) != 0u) {
    *b->abortHost = 0u;
    return bfail(PBD_ERR_CUDA, "a tail piece waited longer than PBD_SPIN_LIMIT_MS for the head piece of its body; the state is invalid");
  }
  return PBD_OK;
}

int pbd_batch_step(pbd_batch* b, float dt, pbd_step_stats* stats) {
  const double t0 = wall_ms();
  int rc = pbd_batch_step_async(b, dt, 1);
  if (rc != PBD_OK) return rc;
  double ms = 0.0;
  rc = pbd_batch_sync(b, &ms);
  if (rc != PBD_OK) return rc;
  if (stats) {
    // the frame is one kernel: predict / commit are shares of its device time (pbd_b200.h pbd_step_stats)
    double a0 = 0, a1 = 0, a2 = 0, tot = 0;
    for (uint32_t c = 0; b->stageHost && c < b->grid; ++c) {
      const volatile uint32_t* o = b->stageHost + 4u * c;
      a0 += o[0]; a1 += o[1]; a2 += o[2]; tot += o[3];
    }
    const double p = tot > 0 ? (a0 + 0.5 * a1) / tot : 0.0, cm = tot > 0 ? (0.5 * a1 + a2) / tot : 0.0;
    stats->predictMs += ms * p; stats->commitMs += ms * cm; stats->solveMs += ms * (1.0 - p - cm);
    stats->totalMs += wall_ms() - t0;
  }
  return PBD_OK;
}

int pbd_batch_read_positions(pbd_batch* b, float* out, double* packMs) {
  if (!b || !out) return bfail(PBD_ERR_INVALID, "null argument");
  const double t0 = wall_ms();
  DeviceScope onDevice(b->device);
  BCU(onDevice.err);
  BCU(launch_pack(b->d, b->stream));
  if (b->Vtot) BCU(cudaMemcpyAsync(out, b->d.packed, sizeof(float) * 3 * b->Vtot, cudaMemcpyDeviceToHost, b->stream));
  BCU(cudaStreamSynchronize(b->stream));
  if (packMs) *packMs += wall_ms() - t0;
  return PBD_OK;
}

int pbd_batch_get_schedule_order(const pbd_batch* b, uint32_t* edgeOrder, uint32_t* tetOrder) {
  if (!b) return bfail(PBD_ERR_INVALID, "batch is null");
  if (edgeOrder && b->Etot) memcpy(edgeOrder, b->edgeOrder.data(), sizeof(uint32_t) * b->Etot);
  if (tetOrder && b->Ttot) memcpy(tetOrder, b->tetOrder.data(), sizeof(uint32_t) * b->Ttot);
  return PBD_OK;
}

int pbd_batch_get_info(const pbd_batch* b, pbd_info* out) {
  if (!b || !out) return bfail(PBD_ERR_INVALID, "null argument");
  memset(out, 0, sizeof(*out));
  out->V = (uint32_t)b->Vtot; out->E = (uint32_t)b->Etot; out->T = (uint32_t)b->Ttot;
  out->backend = PBD_BACKEND_TILE;
  out->edge_colors = b->edgeColorsMax; out->tet_colors = b->tetColorsMax;
  out->edge_phases = b->Etot ? 1 : 0; out->tet_phases = b->Ttot ? 1 : 0;
  out->tiles = b->nBodies;
  out->launches_per_frame = 1;
  out->grid_blocks = b->grid; out->block_threads = 512;
  out->partitions = 1; out->lanes_per_tet = b->lanes;
  out->device_bytes = b->bytes;
  out->algorithmic_bytes_per_substep = algorithmic_bytes_per_substep((uint32_t)b->Vtot, (uint32_t)b->Etot, (uint32_t)b->Ttot, b->params.iterations);
  out->plan_ms = b->planMs;
  out->upload_ms = b->uploadMs;
  for (int ty = 0; ty < 2; ++ty)
    out->gather_wavefronts_permille[ty] = b->gatherIdeal[ty] ? (uint32_t)((1000ull * b->gatherWavefronts[ty] + b->gatherIdeal[ty] / 2) / b->gatherIdeal[ty]) : 0u;
  return PBD_OK;
}

void pbd_batch_destroy(pbd_batch* b) {
  if (!b) return;
  DeviceScope onDevice(b->device);
  if (b->stream) cudaStreamSynchronize(b->stream);
  delete b;
}

}  // extern "C"

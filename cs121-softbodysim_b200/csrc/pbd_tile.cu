// pbd_tile.cu -- "tile" backend: ONE persistent cooperative kernel per frame.
//
// Each CTA owns one shared-memory vertex tile per phase (schedule: pbd_tileplan.cpp):
//
//   for substep, iteration, phase:                        (reference loop nest, Sim.cpp:288-301)
//     for each tile of the phase assigned to this CTA:
//        load the tile's float4 (xStar, invMass) vertices HBM/L2 -> shared memory
//             phase 0 of iteration 0 fuses   [ground + commit of the previous substep] + predict
//             phase 0 of iteration > 0 fuses the ground clamp of the previous iteration
//        for each local colour group:  one thread per constraint, gather 2/4 vertices from shared
//             memory, project (pbd_math.cuh), scatter back;  __syncthreads()
//        store the tile back
//     grid barrier (release/acquire on one L2 counter)
//   final pass: ground + commit of the last substep.
//
// Vertex traffic is coalesced float4 for phase-0 tiles (slots are tile-major) and 16-byte gathers
// for the re-partitioned phases; constraint records stream as 8-byte (2x/4x u16 tile-local index
// [+ rest]) coalesced loads, prefetched one colour group ahead so the only latency on the
// dependent chain is shared memory + arithmetic + the block barrier.  Mutable arrays are accessed
// with .cg (L2) loads/stores so no stale L1 line can be observed after a grid barrier.
//
// Replaces (CProgram/src/Sim.cpp): predict_serial :178-185, solve_edges_xpbd_gs :100-130,
// solve_tets_xpbd_gs :132-173, project_ground_serial :187-195, commit_serial :197-222 and the
// loop nest of SerialStepper::step :280-305.
#include <algorithm>
#include <cstring>
#include <type_traits>
#include <vector>

#include "pbd_body.h"

namespace pbd {

namespace {

struct TileDesc {
  uint32_t vertBegin, vertCount, groupBegin, groupCount;
  uint32_t contiguous, isTet, pad0, pad1;
};
struct GroupDesc {
  uint32_t begin, count;
};
struct PhaseDesc {
  uint32_t tileBegin, tileCount;
};

struct TileParams {
  float4* pos;
  float4* prev;
  float4* vel;
  const uint2* edgeRec;   // {a | b << 16, float bits of rest}
  float* edgeLam;
  const uint2* tetIdx;    // {a | b << 16, c | d << 16}
  const float* tetRest;
  float* tetLam;
  const TileDesc* tiles;
  const GroupDesc* groups;
  const PhaseDesc* phases;
  const uint32_t* tileVerts;
  const uint32_t* tile0Begin;   // nTile0 + 1
  const StepConsts* consts;
  unsigned* barrier;
  uint32_t nTile0, nPhases, substeps, iterations;
};

enum LoadMode { LOAD_PLAIN = 0, LOAD_GROUND = 1, LOAD_PREDICT = 2, LOAD_COMMIT_PREDICT = 3 };

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// All CTAs are co-resident (cooperative launch).  The counter is zeroed before the launch.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += gridDim.x;
    red_release(counter, 1u);
    while (ld_acquire(counter) < epoch) {}
  }
  __syncthreads();
}

// Vertex stage applied while a phase-0 tile is loaded.  p arrives as (xStar, w) from HBM.
__device__ __forceinline__ float4 load_transform(const TileParams& P, const StepConsts& k, uint32_t s, int mode,
                                                 bool clampFirst) {
  float4 p = __ldcg(P.pos + s);
  if (mode == LOAD_GROUND) {
    ground_vertex(p, k);
  } else if (mode == LOAD_PREDICT) {
    const float4 x = __ldcg(P.prev + s);
    float4 v = __ldcg(P.vel + s);
    p = predict_vertex(x, v, p.w, k);
    __stcg(P.vel + s, v);
  } else if (mode == LOAD_COMMIT_PREDICT) {
    float4 x = __ldcg(P.prev + s), v;
    if (clampFirst) ground_vertex(p, k);
    commit_vertex(p, x, v, k);
    v.w = 0.0f;
    p = predict_vertex(x, v, p.w, k);
    __stcg(P.prev + s, x);
    __stcg(P.vel + s, v);
  }
  return p;
}

__device__ __forceinline__ void sweep_edges(const TileParams& P, const TileDesc& td, float4* sv, float alpha) {
  const uint32_t tid = threadIdx.x;
  uint32_t gi = td.groupBegin;
  const uint32_t gend = gi + td.groupCount;
  uint2 rec = make_uint2(0, 0), nrec = rec;
  float lam = 0.f, nlam = 0.f;
  uint32_t e = 0xffffffffu, ne = 0xffffffffu;
  if (gi < gend) {
    const uint2 gq = __ldg(reinterpret_cast<const uint2*>(P.groups + gi)); const GroupDesc g{gq.x, gq.y};
    if (tid < g.count) { e = g.begin + tid; rec = __ldg(P.edgeRec + e); lam = __ldcg(P.edgeLam + e); }
  }
  for (; gi < gend; ++gi) {
    ne = 0xffffffffu;
    if (gi + 1 < gend) {   // prefetch the next colour group's records (independent of the vertex data)
      const uint2 gq = __ldg(reinterpret_cast<const uint2*>(P.groups + gi + 1)); const GroupDesc g{gq.x, gq.y};
      if (tid < g.count) { ne = g.begin + tid; nrec = __ldg(P.edgeRec + ne); nlam = __ldcg(P.edgeLam + ne); }
    }
    if (e != 0xffffffffu) {
      const uint32_t a = rec.x & 0xffffu, b = rec.x >> 16;
      float4 p0 = sv[a], p1 = sv[b];
      if (project_edge(p0, p1, __uint_as_float(rec.y), lam, alpha)) {
        sv[a] = p0;
        sv[b] = p1;
        __stcg(P.edgeLam + e, lam);
      }
    }
    __syncthreads();
    e = ne; rec = nrec; lam = nlam;
  }
}

__device__ __forceinline__ void sweep_tets(const TileParams& P, const TileDesc& td, float4* sv, float alpha) {
  const uint32_t tid = threadIdx.x;
  uint32_t gi = td.groupBegin;
  const uint32_t gend = gi + td.groupCount;
  uint2 rec = make_uint2(0, 0), nrec = rec;
  float lam = 0.f, nlam = 0.f, rest = 0.f, nrest = 0.f;
  uint32_t t = 0xffffffffu, nt = 0xffffffffu;
  if (gi < gend) {
    const uint2 gq = __ldg(reinterpret_cast<const uint2*>(P.groups + gi)); const GroupDesc g{gq.x, gq.y};
    if (tid < g.count) { t = g.begin + tid; rec = __ldg(P.tetIdx + t); rest = __ldg(P.tetRest + t); lam = __ldcg(P.tetLam + t); }
  }
  for (; gi < gend; ++gi) {
    nt = 0xffffffffu;
    if (gi + 1 < gend) {
      const uint2 gq = __ldg(reinterpret_cast<const uint2*>(P.groups + gi + 1)); const GroupDesc g{gq.x, gq.y};
      if (tid < g.count) { nt = g.begin + tid; nrec = __ldg(P.tetIdx + nt); nrest = __ldg(P.tetRest + nt); nlam = __ldcg(P.tetLam + nt); }
    }
    if (t != 0xffffffffu) {
      const uint32_t a = rec.x & 0xffffu, b = rec.x >> 16, c = rec.y & 0xffffu, d = rec.y >> 16;
      float4 pa = sv[a], pb = sv[b], pc = sv[c], pd = sv[d];
      if (project_tet(pa, pb, pc, pd, rest, lam, alpha)) {
        sv[a] = pa;
        sv[b] = pb;
        sv[c] = pc;
        sv[d] = pd;
        __stcg(P.tetLam + t, lam);
      }
    }
    __syncthreads();
    t = nt; rec = nrec; rest = nrest; lam = nlam;
  }
}

__device__ __forceinline__ void run_tile(const TileParams& P, const StepConsts& k, uint32_t tile, int mode,
                                         float4* sv) {
  const TileDesc td = P.tiles[tile];
  const uint32_t tid = threadIdx.x, nth = blockDim.x;
  if (td.contiguous) {
    for (uint32_t i = tid; i < td.vertCount; i += nth) sv[i] = load_transform(P, k, td.vertBegin + i, mode, true);
  } else {
    for (uint32_t i = tid; i < td.vertCount; i += nth) sv[i] = __ldcg(P.pos + __ldg(P.tileVerts + td.vertBegin + i));
  }
  __syncthreads();
  if (td.isTet) sweep_tets(P, td, sv, k.alphaTet); else sweep_edges(P, td, sv, k.alphaEdge);
  if (td.contiguous) {
    for (uint32_t i = tid; i < td.vertCount; i += nth) __stcg(P.pos + td.vertBegin + i, sv[i]);
  } else {
    for (uint32_t i = tid; i < td.vertCount; i += nth) __stcg(P.pos + __ldg(P.tileVerts + td.vertBegin + i), sv[i]);
  }
  __syncthreads();   // sv is reused by the next tile of this CTA
}

// vertex-only pass over the phase-0 partition (no constraints): used when there is nothing to
// sweep and for the final commit.  finalCommit: ground (if clamp) + commit, no predict.
__device__ __forceinline__ void vertex_pass(const TileParams& P, const StepConsts& k, int mode, bool clamp,
                                            bool finalCommit) {
  for (uint32_t t = blockIdx.x; t < P.nTile0; t += gridDim.x) {
    const uint32_t b = P.tile0Begin[t], e = P.tile0Begin[t + 1];
    for (uint32_t s = b + threadIdx.x; s < e; s += blockDim.x) {
      if (finalCommit) {
        float4 p = __ldcg(P.pos + s), x = __ldcg(P.prev + s), v;
        if (clamp) ground_vertex(p, k);
        commit_vertex(p, x, v, k);
        v.w = 0.0f;
        __stcg(P.prev + s, x);
        __stcg(P.vel + s, v);
        __stcg(P.pos + s, p);
      } else {
        const float4 p = load_transform(P, k, s, mode, clamp);
        __stcg(P.pos + s, p);
      }
    }
  }
}

__global__ void __launch_bounds__(1024, 1) tile_frame_kernel(const TileParams P) {
  extern __shared__ float4 sv[];
  const StepConsts k = *P.consts;
  unsigned epoch = 0;
  const bool sweeping = P.iterations > 0 && P.nPhases > 0;
  const bool clamp = P.iterations > 0;   // the reference clamps once per iteration (Sim.cpp:296)
  for (uint32_t sub = 0; sub < P.substeps; ++sub) {
    if (!sweeping) {
      // vertex-local work only: a vertex always belongs to the same CTA, no grid barrier needed
      vertex_pass(P, k, sub == 0 ? LOAD_PREDICT : LOAD_COMMIT_PREDICT, clamp, false);
      continue;
    }
    for (uint32_t it = 0; it < P.iterations; ++it) {
      for (uint32_t ph = 0; ph < P.nPhases; ++ph) {
        const PhaseDesc pd = P.phases[ph];
        const int mode = ph != 0 ? LOAD_PLAIN : it != 0 ? LOAD_GROUND : sub != 0 ? LOAD_COMMIT_PREDICT : LOAD_PREDICT;
        for (uint32_t t = blockIdx.x; t < pd.tileCount; t += gridDim.x) run_tile(P, k, pd.tileBegin + t, mode, sv);
        grid_barrier(P.barrier, epoch);
      }
    }
  }
  vertex_pass(P, k, LOAD_PLAIN, clamp, true);
}

class TileBackend final : public Backend {
 public:
  TileBackend(const pbd_options& o, int device) : opts_(o), device_(device) {}
  ~TileBackend() override {
    cudaFree(edgeRec_); cudaFree(tetIdx_); cudaFree(tiles_); cudaFree(groups_); cudaFree(phases_);
    cudaFree(tileVerts_); cudaFree(tile0Begin_); cudaFree(barrier_);
  }
  const char* name() const override { return "b200-tile"; }

  cudaError_t upload(const Plan& plan, const MeshView& m, DeviceArrays& d) override {
    (void)m;
    cudaError_t err;
    block_ = plan.blockThreads ? plan.blockThreads : 512;
    nPhases_ = (uint32_t)plan.phases.size();
    nTile0_ = (uint32_t)plan.tile0Begin.size() - 1;
    nTiles_ = (uint32_t)plan.tiles.size();
    smemBytes_ = sizeof(float4) * (size_t)std::max(plan.tileVertexCapacity, 1u);

    // device copies of the rest values are already in schedule order (pbd_capi.cu); pack the
    // edge rest next to the indices so one 8-byte load fetches the whole edge record
    std::vector<float> eRest(plan.E);
    if (plan.E && (err = cudaMemcpy(eRest.data(), d.edgeRest, sizeof(float) * plan.E, cudaMemcpyDeviceToHost)) != cudaSuccess) return err;
    std::vector<uint2> er(plan.E), ti(plan.T);
    for (uint32_t k = 0; k < plan.E; ++k) {
      uint32_t bits;
      memcpy(&bits, &eRest[k], 4);
      er[k] = make_uint2((uint32_t)plan.edgeLocal[2 * (size_t)k] | ((uint32_t)plan.edgeLocal[2 * (size_t)k + 1] << 16), bits);
    }
    for (uint32_t k = 0; k < plan.T; ++k) {
      const uint16_t* l = &plan.tetLocal[4 * (size_t)k];
      ti[k] = make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
    }
    std::vector<TileDesc> td(plan.tiles.size());
    for (size_t i = 0; i < td.size(); ++i) {
      const Tile& t = plan.tiles[i];
      td[i] = TileDesc{t.vertBegin, t.vertCount, t.groupBegin, t.groupCount, t.contiguous, t.isTet, 0, 0};
    }
    std::vector<GroupDesc> gd(plan.groups.size());
    for (size_t i = 0; i < gd.size(); ++i) gd[i] = GroupDesc{plan.groups[i].begin, plan.groups[i].count};
    std::vector<PhaseDesc> pd(plan.phases.size());
    maxTilesPerPhase_ = nTile0_;
    for (size_t i = 0; i < pd.size(); ++i) {
      pd[i] = PhaseDesc{plan.phases[i].tileBegin, plan.phases[i].tileCount};
      maxTilesPerPhase_ = std::max(maxTilesPerPhase_, plan.phases[i].tileCount);
    }
    auto up = [&](auto** dst, const auto& src) -> cudaError_t {
      using T = typename std::remove_reference<decltype(src)>::type::value_type;
      cudaError_t e = cudaMalloc((void**)dst, sizeof(T) * (src.size() + 1));
      if (e != cudaSuccess) return e;
      bytes_ += sizeof(T) * src.size();
      if (!src.empty()) e = cudaMemcpy(*dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice);
      return e;
    };
    if ((err = up(&edgeRec_, er)) != cudaSuccess) return err;
    if ((err = up(&tetIdx_, ti)) != cudaSuccess) return err;
    if ((err = up(&tiles_, td)) != cudaSuccess) return err;
    if ((err = up(&groups_, gd)) != cudaSuccess) return err;
    if ((err = up(&phases_, pd)) != cudaSuccess) return err;
    if ((err = up(&tileVerts_, plan.tileVerts)) != cudaSuccess) return err;
    if ((err = up(&tile0Begin_, plan.tile0Begin)) != cudaSuccess) return err;
    if ((err = cudaMalloc((void**)&barrier_, 256)) != cudaSuccess) return err;

    if ((err = cudaFuncSetAttribute(tile_frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes_)) != cudaSuccess) return err;
    int perSM = 0, nSM = 0, coop = 0;
    if ((err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, tile_frame_kernel, (int)block_, smemBytes_)) != cudaSuccess) return err;
    cudaDeviceGetAttribute(&nSM, cudaDevAttrMultiProcessorCount, device_);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device_);
    if (!coop || perSM < 1) return cudaErrorCooperativeLaunchTooLarge;
    grid_ = std::max(1u, std::min(maxTilesPerPhase_, (uint32_t)(perSM * nSM)));
    return cudaSuccess;
  }

  cudaError_t enqueue_frame(const DeviceArrays& d, const FrameShape& f, cudaStream_t s) override {
    TileParams P{};
    P.pos = d.pos; P.prev = d.prev; P.vel = d.vel;
    P.edgeRec = edgeRec_; P.edgeLam = d.edgeLam;
    P.tetIdx = tetIdx_; P.tetRest = d.tetRest; P.tetLam = d.tetLam;
    P.tiles = tiles_; P.groups = groups_; P.phases = phases_; P.tileVerts = tileVerts_;
    P.tile0Begin = tile0Begin_; P.consts = d.consts; P.barrier = barrier_;
    P.nTile0 = nTile0_; P.nPhases = nPhases_; P.substeps = f.substeps; P.iterations = f.iterations;
    cudaError_t err = cudaMemsetAsync(barrier_, 0, sizeof(unsigned), s);
    if (err != cudaSuccess) return err;
    void* args[] = {&P};
    return cudaLaunchCooperativeKernel((const void*)tile_frame_kernel, dim3(grid_), dim3(block_), args, smemBytes_, s);
  }

  uint32_t launches_per_frame(const FrameShape&) const override { return 1; }
  uint64_t device_bytes() const override { return bytes_; }
  void fill_info(pbd_info& info) const override {
    info.grid_blocks = grid_;
    info.block_threads = block_;
  }

 private:
  pbd_options opts_;
  int device_;
  uint2* edgeRec_ = nullptr;
  uint2* tetIdx_ = nullptr;
  TileDesc* tiles_ = nullptr;
  GroupDesc* groups_ = nullptr;
  PhaseDesc* phases_ = nullptr;
  uint32_t* tileVerts_ = nullptr;
  uint32_t* tile0Begin_ = nullptr;
  unsigned* barrier_ = nullptr;
  uint32_t block_ = 512, grid_ = 1, nPhases_ = 0, nTile0_ = 0, nTiles_ = 0, maxTilesPerPhase_ = 0;
  size_t smemBytes_ = 0;
  uint64_t bytes_ = 0;
};

}  // namespace

Backend* make_tile_backend(const pbd_options& opts, int device) { return new TileBackend(opts, device); }

}  // namespace pbd

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
#include <cstdio>
#include <cstdlib>
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
  long long* ftrace;           // debug: fine-grained clock64 stamps of CTA 0 (phase-major, 128 per phase)
  unsigned long long* trace;   // debug: [phase][cta][2] globaltimer ns of (start, arrive) in substep 0, last iteration
  uint32_t nTile0, nPhases, substeps, iterations;
  uint32_t dbgFlags;   // experiments only (PBD_TILE_DBG): 1 = skip lambda stores, 2 = default-policy stores
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

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
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

// One prefetched constraint record (registers).  id == NONE32: this thread idles in that group.
constexpr uint32_t NONE32 = 0xffffffffu;
constexpr int kPrefetch = 3;        // colour groups in flight per thread (hides ~L2 latency)
constexpr uint32_t kMaxGroupsSmem = 1024;  // group descriptors staged in shared memory per tile

struct Rec {
  uint32_t id;
  uint2 idx;
  float rest, lam;
};

template <bool TET>
__device__ __forceinline__ void fetch_rec(const TileParams& P, const uint2* sg, uint32_t g, uint32_t gcount, Rec& r) {
  r.id = NONE32;
  if (g < gcount) {
    const uint2 gd = sg[g];
    const uint32_t item = TET ? (threadIdx.x >> 2) : threadIdx.x;   // a tet is shared by 4 lanes
    if (item < gd.y) {
      r.id = gd.x + item;
      if (TET) {
        r.idx = __ldg(P.tetIdx + r.id);
        r.rest = __ldg(P.tetRest + r.id);
        r.lam = __ldcg(P.tetLam + r.id);
      } else {
        r.idx = __ldg(P.edgeRec + r.id);
        r.rest = __uint_as_float(r.idx.y);
        r.lam = __ldcg(P.edgeLam + r.id);
      }
    }
  }
}

// `z` is a zero read from shared memory AFTER the preceding block barrier.  Mixing it into the
// prefetched record makes every use of the record data-dependent on a post-barrier load, so
// neither nvcc nor ptxas can hoist the consumers (index unpacking, alpha*lambda) up to the point
// right after the global load -- where an in-order warp would stall for the whole L2 latency
// and defeat the prefetch (measured: ~1 us per colour step before this, see profiles/).
template <bool TET>
__device__ __forceinline__ void apply_rec(const TileParams& P, float4* sv, Rec& r, float alpha, uint32_t z) {
  if (r.id == NONE32) return;
  r.idx.x ^= z; r.idx.y ^= z;
  r.rest = __uint_as_float(__float_as_uint(r.rest) ^ z);
  r.lam = __uint_as_float(__float_as_uint(r.lam) ^ z);
  if (TET) {
    // One tet per 4 adjacent lanes, lane `role` owns vertex `role` (a,b,c,d).  All four gradients
    // have the form cross(x - o, y - o)/6 (Sim.cpp:146-149):
    //   ga: o=b x=d y=c | gb: o=a x=c y=d | gc: o=a x=d y=b | gd: o=a x=b y=c
    // so every lane runs the same instructions on role-selected operands; the reduction terms are
    // exchanged with quad shuffles and summed in the reference's order, which keeps the result
    // bit-identical while cutting the per-step dependent instruction stream ~3x.
    const uint32_t role = threadIdx.x & 3u;
    const uint32_t lane = threadIdx.x & 31u, base = lane & ~3u;
    const unsigned m = __activemask();
    auto pick = [&](uint32_t f) -> uint32_t { return (((f & 2u) ? r.idx.y : r.idx.x) >> ((f & 1u) * 16u)) & 0xffffu; };
    const uint32_t fo = (0x00000001u >> (role * 8u)) & 3u;        // {1,0,0,0}
    const uint32_t fx = (0x01030203u >> (role * 8u)) & 3u;        // {3,2,3,1}
    const uint32_t fy = (0x02010302u >> (role * 8u)) & 3u;        // {2,3,1,2}
    const uint32_t iown = pick(role);
    float4 own = sv[iown];
    const float4 o = sv[pick(fo)], x = sv[pick(fx)], y = sv[pick(fy)];
    const float wa = __shfl_sync(m, own.w, base), wb = __shfl_sync(m, own.w, base + 1),
                wc = __shfl_sync(m, own.w, base + 2), wd = __shfl_sync(m, own.w, base + 3);
    if (fadd(fadd(fadd(wa, wb), wc), wd) != 0.0f) {               // quad-uniform
      const float k6 = 1.0f / 6.0f;
      const float ux = fsub(x.x, o.x), uy = fsub(x.y, o.y), uz = fsub(x.z, o.z);
      const float vx = fsub(y.x, o.x), vy = fsub(y.y, o.y), vz = fsub(y.z, o.z);
      const float nx = cross_c(uy, vz, uz, vy), ny = cross_c(uz, vx, ux, vz), nz = cross_c(ux, vy, uy, vx);
      const float gx = fmul(nx, k6), gy = fmul(ny, k6), gz = fmul(nz, k6);
      const float t = fmul(own.w, dot3(gx, gy, gz, gx, gy, gz));
      const float ta = __shfl_sync(m, t, base), tb = __shfl_sync(m, t, base + 1), tc = __shfl_sync(m, t, base + 2),
                  td = __shfl_sync(m, t, base + 3);
      const float wSum = fadd(fadd(fadd(ta, tb), tc), td);
      // role 3 holds n = cross(pb-pa, pc-pa) and own - o = pd - pa: the volume numerator
      const float vn = dot3(nx, ny, nz, fsub(own.x, o.x), fsub(own.y, o.y), fsub(own.z, o.z));
      const float vol = fdiv(__shfl_sync(m, vn, base + 3), 6.0f);
      if (!(wSum < 1e-20f)) {                                      // quad-uniform
        const float C = fsub(vol, r.rest);
        const float dl = fdiv(fsub(-C, fmul(alpha, r.lam)), fadd(wSum, alpha));
        const float sc = fmul(own.w, dl);
        own.x = fadd(own.x, fmul(gx, sc)); own.y = fadd(own.y, fmul(gy, sc)); own.z = fadd(own.z, fmul(gz, sc));
        sv[iown] = own;
        if (role == 0) {
          const float l = fadd(r.lam, dl);
          if (P.dbgFlags & 2) P.tetLam[r.id] = l; else if (!(P.dbgFlags & 1)) __stcg(P.tetLam + r.id, l);
        }
      }
    }
  } else {
    const uint32_t a = r.idx.x & 0xffffu, b = r.idx.x >> 16;
    float4 p0 = sv[a], p1 = sv[b];
    if (project_edge(p0, p1, r.rest, r.lam, alpha)) {
      sv[a] = p0; sv[b] = p1;
      if (P.dbgFlags & 2) P.edgeLam[r.id] = r.lam; else if (!(P.dbgFlags & 1)) __stcg(P.edgeLam + r.id, r.lam);
    }
  }
}

// Sweep the colour groups of one tile.  sg: the tile's group descriptors in shared memory.
template <bool TET>
__device__ __forceinline__ void sweep(const TileParams& P, const uint2* sg, uint32_t gcount, float4* sv, float alpha,
                                      const volatile uint32_t* zero, long long* ft) {
  Rec r[kPrefetch];
  int fi = 4;
#pragma unroll
  for (int d = 0; d < kPrefetch; ++d) fetch_rec<TET>(P, sg, d, gcount, r[d]);
  for (uint32_t g = 0; g < gcount; g += kPrefetch) {
#pragma unroll
    for (int d = 0; d < kPrefetch; ++d) {
      if (g + d < gcount) {   // block-uniform
        apply_rec<TET>(P, sv, r[d], alpha, *zero);
        __syncthreads();
        if (ft && threadIdx.x == 0 && fi < 120) ft[fi++] = clock64();
        fetch_rec<TET>(P, sg, g + d + kPrefetch, gcount, r[d]);
      }
    }
  }
}

__device__ __forceinline__ void run_tile(const TileParams& P, const StepConsts& k, uint32_t tile, int mode,
                                         float4* sv, uint2* sg, const volatile uint32_t* zero, long long* ft) {
  if (ft && threadIdx.x == 0) ft[0] = clock64();
  const TileDesc td = P.tiles[tile];
  const uint32_t tid = threadIdx.x, nth = blockDim.x;
  if (td.contiguous) {
    for (uint32_t i = tid; i < td.vertCount; i += nth) sv[i] = load_transform(P, k, td.vertBegin + i, mode, true);
  } else {
    for (uint32_t i = tid; i < td.vertCount; i += nth) sv[i] = __ldcg(P.pos + __ldg(P.tileVerts + td.vertBegin + i));
  }
  if (ft && threadIdx.x == 0) ft[1] = clock64();
  // group descriptors -> shared memory (the record prefetch must not wait on a dependent global load)
  for (uint32_t gb = 0; gb < td.groupCount; gb += kMaxGroupsSmem) {
    const uint32_t gn = min(kMaxGroupsSmem, td.groupCount - gb);
    for (uint32_t i = tid; i < gn; i += nth) sg[i] = __ldg(reinterpret_cast<const uint2*>(P.groups + td.groupBegin + gb + i));
    __syncthreads();
    if (ft && threadIdx.x == 0) ft[2] = clock64();
    if (td.isTet) sweep<true>(P, sg, gn, sv, k.alphaTet, zero, ft); else sweep<false>(P, sg, gn, sv, k.alphaEdge, zero, ft);
  }
  if (td.groupCount == 0) __syncthreads();
  if (td.contiguous) {
    for (uint32_t i = tid; i < td.vertCount; i += nth) __stcg(P.pos + td.vertBegin + i, sv[i]);
  } else {
    for (uint32_t i = tid; i < td.vertCount; i += nth) __stcg(P.pos + __ldg(P.tileVerts + td.vertBegin + i), sv[i]);
  }
  __syncthreads();   // sv is reused by the next tile of this CTA
  if (ft && threadIdx.x == 0) { ft[3] = clock64(); ft[127] = td.groupCount; }
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

__global__ void __launch_bounds__(512, 1) tile_frame_kernel(const TileParams P) {
  extern __shared__ float4 sv[];
  __shared__ uint2 sg[kMaxGroupsSmem];
  __shared__ uint32_t szero;
  if (threadIdx.x == 0) szero = 0u;
  __syncthreads();
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
        const bool tr = P.trace && sub == 0 && it + 1 == P.iterations && threadIdx.x == 0;
        if (tr) P.trace[2 * ((size_t)ph * gridDim.x + blockIdx.x)] = globaltimer_ns();
        for (uint32_t t = blockIdx.x; t < pd.tileCount; t += gridDim.x) run_tile(P, k, pd.tileBegin + t, mode, sv, sg, &szero, (P.ftrace && tr && blockIdx.x == 0) ? P.ftrace + 128 * ph : nullptr);
        if (tr) P.trace[2 * ((size_t)ph * gridDim.x + blockIdx.x) + 1] = globaltimer_ns();
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
    if (getenv("PBD_TILE_TRACE")) {
      traceN_ = 2 * (size_t)(nPhases_ + 1) * 4096;
      if ((err = cudaMalloc((void**)&trace_, sizeof(unsigned long long) * traceN_)) != cudaSuccess) return err;
      cudaMemset(trace_, 0, sizeof(unsigned long long) * traceN_);
      cudaMalloc((void**)&ftrace_, sizeof(long long) * 128 * (nPhases_ + 1));
      cudaMemset(ftrace_, 0, sizeof(long long) * 128 * (nPhases_ + 1));
    }

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
    P.tile0Begin = tile0Begin_; P.consts = d.consts; P.barrier = barrier_; P.trace = trace_; P.ftrace = ftrace_;
    P.dbgFlags = getenv("PBD_TILE_DBG") ? (uint32_t)atoi(getenv("PBD_TILE_DBG")) : 0u;
    P.nTile0 = nTile0_; P.nPhases = nPhases_; P.substeps = f.substeps; P.iterations = f.iterations;
    cudaError_t err = cudaMemsetAsync(barrier_, 0, sizeof(unsigned), s);
    if (err != cudaSuccess) return err;
    void* args[] = {&P};
    return cudaLaunchCooperativeKernel((const void*)tile_frame_kernel, dim3(grid_), dim3(block_), args, smemBytes_, s);
  }

  void debug_dump() override {
    if (!trace_) return;
    std::vector<unsigned long long> t(traceN_);
    cudaMemcpy(t.data(), trace_, sizeof(unsigned long long) * traceN_, cudaMemcpyDeviceToHost);
    unsigned long long prevEnd = 0;
    for (uint32_t ph = 0; ph < nPhases_; ++ph) {
      unsigned long long s0 = ~0ull, sMax = 0, aMax = 0, busySum = 0, busyMax = 0;
      uint32_t busyN = 0;
      for (uint32_t c = 0; c < grid_; ++c) {
        const unsigned long long s = t[2 * ((size_t)ph * grid_ + c)], a = t[2 * ((size_t)ph * grid_ + c) + 1];
        s0 = std::min(s0, s); sMax = std::max(sMax, s); aMax = std::max(aMax, a);
        busySum += a - s; busyMax = std::max(busyMax, a - s);
        busyN += (a - s) > 300;
      }
      fprintf(stderr, "[pbd-trace] phase %u: start skew %.2f us, phase span %.2f us (busiest CTA %.2f us, mean busy %.2f us, %u CTAs busy), gap since prev %.2f us\n",
              ph, (sMax - s0) * 1e-3, (aMax - s0) * 1e-3, busyMax * 1e-3, busySum * 1e-3 / grid_, busyN,
              prevEnd ? (double)(s0 - prevEnd) * 1e-3 : 0.0);
      prevEnd = aMax;
    }
    std::vector<long long> f(128 * (size_t)nPhases_);
    cudaMemcpy(f.data(), ftrace_, sizeof(long long) * f.size(), cudaMemcpyDeviceToHost);
    for (uint32_t ph = 0; ph < nPhases_; ++ph) {
      const long long* q = &f[128 * (size_t)ph];
      fprintf(stderr, "[pbd-ftrace] phase %u CTA0: groups %lld | vertex load %lld cyc | desc staging %lld | sweep+store %lld | steps:", ph, q[127], q[1] - q[0], q[2] - q[1], q[3] - q[2]);
      long long prev = q[2];
      for (int i = 4; i < 120 && q[i]; ++i) { fprintf(stderr, " %lld", q[i] - prev); prev = q[i]; }
      fprintf(stderr, "\n");
    }
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
  unsigned long long* trace_ = nullptr;
  long long* ftrace_ = nullptr;
  size_t traceN_ = 0;
  uint32_t block_ = 512, grid_ = 1, nPhases_ = 0, nTile0_ = 0, nTiles_ = 0, maxTilesPerPhase_ = 0;
  size_t smemBytes_ = 0;
  uint64_t bytes_ = 0;
};

}  // namespace

Backend* make_tile_backend(const pbd_options& opts, int device) { return new TileBackend(opts, device); }

}  // namespace pbd

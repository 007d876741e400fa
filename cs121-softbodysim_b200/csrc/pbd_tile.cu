// pbd_tile.cu -- "tile" backend: ONE persistent cooperative kernel per frame.
//
// Every CTA walks the schedule of pbd_tileplan.cpp:
//
//   for substep, iteration, phase:                        (reference loop nest, Sim.cpp:288-301)
//     for each tile of the phase assigned to this CTA (normally exactly one):
//        wait for the tile's RECORD BLOCK in shared memory.  It was fetched one tile ahead by
//             cp.async.bulk (TMA bulk copies completing on an mbarrier): tile-local u16 vertex
//             indices, rest values, colour-group table, gathered vertex slots, and the tile's
//             lambdas -- everything a sweep needs except the vertex positions
//        load the tile's float4 (xStar, invMass) vertices L2 -> shared memory
//             phase 0 of iteration 0 fuses   [ground + commit of the previous substep] + predict
//             phase 0 of iteration > 0 fuses the ground clamp of the previous iteration
//        for each local colour group of edges, then of tets: one thread (or four lanes) per
//             constraint, gather 2/4 vertices from shared memory, project (pbd_math.cuh),
//             scatter back;  __syncthreads()
//        store the vertices back; bulk-store the tile's lambdas
//     grid barrier (release/acquire on one L2 counter)
//   final pass: ground + commit of the last substep.
//
// The dependent chain of a sweep is therefore shared memory + arithmetic + block barrier per
// colour, and L2 only once per phase.  Mutable vertex arrays are accessed with .cg (L2)
// loads/stores so no stale L1 line can be observed after a grid barrier.
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
#include "pbd_device.cuh"
#include "pbd_sweep.cuh"

namespace pbd {

namespace {

// what the copy-issuing thread needs per tile (global memory)
struct TileCopy {
  unsigned long long blobOff;   // byte offset of the static part in the blob
  uint32_t staticBytes;
  uint32_t edgeDevBegin, edgeLamBytes;   // lambda range: first device index, bytes (multiple of 16)
  uint32_t tetDevBegin, tetLamBytes;
  uint32_t pad;
};

struct PhaseDesc {
  uint32_t tileBegin, tileCount;
};

constexpr uint32_t kMaxItems = 512;      // tiles one CTA visits per iteration
constexpr uint32_t kItemCopySmem = 128;  // ... whose copy descriptors are cached in shared memory
constexpr uint32_t kMaxRanks = 8;     // GPUs one body can be spread over (one node)

struct TileParams {
  float4* pos;
  uint4* posT;                  // tagged hand-over: {x, y, z, tag} per vertex (one 128-bit word), else null
  const float* invMass;         // ... and the inverse masses (static), slot order
  float4* prev;
  float4* vel;
  const unsigned char* blob;
  const TileCopy* copies;
  float* edgeLam;
  float* tetLam;
  const PhaseDesc* phases;
  const uint32_t* tile0Begin;   // nTile0 + 1
  const StepConsts* consts;
  const ColliderSet* colliders;  // primitive colliders of the clamp stage (nColliders == 0: none)
  uint32_t nColliders;
  unsigned* barrier;
  unsigned* done;               // per tile: completed visits (point-to-point sync), or null: grid barrier per phase
  // one body across several GPUs: every rank steps the tiles it owns; a vertex / a done counter
  // lives on the rank that owns its home tile / its tile and is read and written in place over
  // NVLink (peer pointers, opened with CUDA IPC or peer access).  world == 1: entry 0 = this GPU.
  float4* posPeers[kMaxRanks];
  unsigned* donePeers[kMaxRanks];
  const uint32_t* tileList;     // this rank's tiles, phase by phase (PhaseDesc indexes it)
  const uint32_t* homeList;     // this rank's home tiles (indices into tile0Begin)
  uint32_t nHome, world, rank;
  uint32_t iterBase;            // iterations completed by earlier frames: the done counters never reset
  unsigned long long* trace;    // debug: [phase][cta][2] globaltimer ns of (start, arrive), substep 0, last iteration
  long long* ftrace;            // debug: clock64 stamps of CTA 0, 128 per phase
  uint32_t nTile0, nPhases, substeps, iterations;
  uint32_t recStride;           // bytes of one record buffer
  long long spinLimit;          // cycles a CTA may wait for another CTA before it gives up (bounded spins)
  unsigned* abortWord;          // device: set by the first wait that gives up; every later wait bails out at once
  unsigned* abortHost;          // mapped host word: what pbd_sync checks after the frame
  uint32_t* stageHost;          // mapped host memory, 4 words per CTA: cycles its thread 0 spent in (predict loads,
                                // fused commit+predict loads, final commit pass, the whole frame) -- pbd_step_stats
  uint32_t stagger;             // cycles by which every second CTA of an SM delays its sweeps (tiles_per_sm >= 2)
};


// Every wait on another CTA (or another GPU) is BOUNDED: a peer rank that was never launched, or a
// lost update, must not hang the GPU until reset.  Every 1024 polls the waiter looks at the clock and
// at the abort word; the first waiter that exceeds P.spinLimit sets the word (and its mapped host
// copy), every other wait then bails out at its next check, the kernel runs to its end with
// meaningless results and pbd_sync reports PBD_ERR_CUDA instead of hanging.
struct SpinGuard {
  uint32_t polls = 0;
  long long t0 = 0;
};
__device__ __forceinline__ bool spin_expired(SpinGuard& g, unsigned* abortWord, unsigned* abortHost, long long spinLimit) {
  if ((++g.polls & 1023u) != 0u) return false;
  if (g.polls == 1024u) g.t0 = clock64();
  if (ld_acquire(abortWord) != 0u) return true;
  if (clock64() - g.t0 > spinLimit) {
    atomicExch(abortWord, 1u);
    *reinterpret_cast<volatile unsigned*>(abortHost) = 1u;
    __threadfence_system();
    return true;
  }
  return false;
}
__device__ __forceinline__ bool spin_expired(SpinGuard& g, const TileParams& P) { return spin_expired(g, P.abortWord, P.abortHost, P.spinLimit); }
// done counters and barrier epochs only ever grow and may wrap: compare by signed distance
__device__ __forceinline__ bool reached(unsigned have, unsigned need) { return (int)(have - need) >= 0; }

// All CTAs are co-resident (cooperative launch).  The counter is zeroed before the launch.
__device__ __forceinline__ void grid_barrier(const TileParams& P, unsigned* counter, unsigned& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += gridDim.x;
    red_release(counter, 1u);
    SpinGuard g;
    while (!reached(ld_acquire(counter), epoch) && !spin_expired(g, P)) {}
  }
  __syncthreads();
}

// ---- tagged hand-over (PBD_FLAG_TAGGED_HANDOVER) ------------------------------------------------
// A position travels between tiles as ONE 128-bit word {x, y, z, tag}, written and read with scalar
// 128-bit accesses (st/ld.relaxed.gpu.b128 -> STG/LDG.E.128.STRONG.GPU): a naturally aligned scalar
// access is a single memory operation, so a reader that finds the expected tag holds the x, y, z
// written WITH that tag -- no release fence on the writer (MEMBAR.GPU + L1 invalidate, ~1.3 k cycles
// per tile visit), no done flag, no separate poll round trip.  The inverse mass never changes and
// lives in its own read-only array (P.invMass, slot order).  Tags: a tile visit with sequence number
// seq = iteration * nPhases + phase (iterations counted across frames) writes 2*seq + 2 and, every
// phase covering every vertex exactly once, expects 2*seq (the visit before it); the commit pass at
// the end of a frame writes 2*seqEnd + 1, which is what the first visit of the next frame expects.
// Tags are compared for equality only, so the 32-bit wrap is harmless.
// `sys`: the word may live on / be written from another GPU of the node (one body across several GPUs):
// system scope, STG/LDG.E.128.STRONG.SYS over NVLink; one GPU: gpu scope.
__device__ __forceinline__ uint4 ld_tagged(const uint4* p, bool sys) {
  unsigned __int128 v;
  if (sys) asm volatile("ld.relaxed.sys.global.b128 %0, [%1];" : "=q"(v) : "l"(p) : "memory");
  else asm volatile("ld.relaxed.gpu.global.b128 %0, [%1];" : "=q"(v) : "l"(p) : "memory");
  return make_uint4((uint32_t)v, (uint32_t)(v >> 32), (uint32_t)(v >> 64), (uint32_t)(v >> 96));
}
__device__ __forceinline__ void st_tagged(uint4* p, float4 q, uint32_t tag, bool sys) {
  const unsigned __int128 v = (unsigned __int128)__float_as_uint(q.x) | ((unsigned __int128)__float_as_uint(q.y) << 32) |
                              ((unsigned __int128)__float_as_uint(q.z) << 64) | ((unsigned __int128)tag << 96);
  if (sys) asm volatile("st.relaxed.sys.global.b128 [%0], %1;" ::"l"(p), "q"(v) : "memory");
  else asm volatile("st.relaxed.gpu.global.b128 [%0], %1;" ::"l"(p), "q"(v) : "memory");
}
__device__ __forceinline__ float4 tagged_value(uint4 r, float w) {
  return make_float4(__uint_as_float(r.x), __uint_as_float(r.y), __uint_as_float(r.z), w);
}
// (out of line, and with scalar arguments only -- a reference to the launch parameters would force a copy of them into local
// memory: the slow path of a vertex load, taken when the word's tag was not there yet.  Keeping its spin and the
// bounded-wait bookkeeping out of the unrolled load loops shortens the kernel and frees registers around them.)
__device__ __noinline__ float4 tagged_wait_slow(const uint4* word, uint32_t expect, float w, bool sys, unsigned* abortWord,
                                                unsigned* abortHost, long long spinLimit) {
  SpinGuard g;
  uint4 r;
  do { r = ld_tagged(word, sys); } while (r.w != expect && !spin_expired(g, abortWord, abortHost, spinLimit));
  return tagged_value(r, w);
}
__device__ __forceinline__ float4 tagged_wait_load(const TileParams& P, const uint4* word, uint32_t expect, float w, bool sys) {
#ifdef PBD_X_INLINE_WAIT
  SpinGuard g;
  uint4 r;
  do { r = ld_tagged(word, sys); } while (r.w != expect && !spin_expired(g, P));
  return tagged_value(r, w);
#else
  return tagged_wait_slow(word, expect, w, sys, P.abortWord, P.abortHost, P.spinLimit);
#endif
}
__device__ __forceinline__ float4 tagged_load_any(const TileParams& P, uint32_t s) {   // own earlier write: no wait
  return tagged_value(ld_tagged(P.posT + s, P.world > 1), __ldg(P.invMass + s));
}
__device__ __forceinline__ void tagged_store(uint4* word, float4 p, uint32_t tag, bool sys) { st_tagged(word, p, tag, sys); }

// vertex-only pass over the phase-0 partition (no constraints): used when there is nothing to
// sweep and for the final commit.  finalCommit: ground (if clamp) + commit, no predict.
// TAGGED: `waitTag` != 0: wait for that tag (values written by other CTAs); 0: the vertex was last
// written by this very thread.  Stores carry `writeTag`.
template <bool TAGGED>
__device__ __forceinline__ void vertex_pass(const TileParams& P, const StepConsts& k, int mode, bool clamp,
                                            bool finalCommit, uint32_t waitTag = 0, uint32_t writeTag = 0) {
  for (uint32_t q = blockIdx.x; q < P.nHome; q += gridDim.x) {
    const uint32_t t = P.homeList[q];
    const uint32_t b = P.tile0Begin[t], e = P.tile0Begin[t + 1];
    for (uint32_t s = b + threadIdx.x; s < e; s += blockDim.x) {
      if (finalCommit) {
        float4 p, x = __ldcg(P.prev + s), v;
        if (TAGGED) p = waitTag ? tagged_wait_load(P, P.posT + s, waitTag, __ldg(P.invMass + s), P.world > 1) : tagged_load_any(P, s);
        else p = __ldcg(P.pos + s);
        if (clamp) { ground_vertex(p, k); if (P.nColliders) collide_vertex(p, P.colliders, P.nColliders); }
        commit_vertex(p, x, v, k);
        v.w = 0.0f;
        __stcg(P.prev + s, x);
        __stcg(P.vel + s, v);
        if (TAGGED) tagged_store(P.posT + s, p, writeTag, P.world > 1); else __stcg(P.pos + s, p);
      } else if (TAGGED) {
        VertexIn in;
        in.p = tagged_load_any(P, s);
        in.x = in.v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (mode == LOAD_PREDICT || mode == LOAD_COMMIT_PREDICT) in.x = __ldcg(P.prev + s);
        if (mode == LOAD_PREDICT) in.v = __ldcg(P.vel + s);
        tagged_store(P.posT + s, finish_vertex(P, k, s, mode, clamp, in), writeTag, P.world > 1);
      } else {
        const float4 p = load_transform(P, k, s, mode, clamp);
        __stcg(P.pos + s, p);
      }
    }
  }
}

// ---------------------------------------------------------------- the frame kernel

// TETLAM = false (only with FAST): zero volume compliance (the reference's default, PBDServer.h:154) makes alpha == 0, so
// a tet's multiplier never enters a correction.  The fast mode then does not carry it at all -- 8 bytes per tet and visit of
// L2 traffic less, one LDS and one STS per projection less (the lambda write-back of a visit alone measured 3 % of the
// frame).  PBD_ARRAY_TET_LAMBDA keeps whatever it held; the exact mode always accumulates it like the reference.  A
// compile-time switch: the kernel sits at its register limit and a run-time flag spilled (measured -4 %).
// RES = true (only with TAGGED; three record buffers): a CTA with exactly four tile visits per iteration -- the one-wave
// case, four shifted partitions -- keeps the record blocks of visits 0 and 2 RESIDENT in shared memory for the whole frame
// (fetched once, their lambdas written back once, after the tile's last visit) and streams the blocks of visits 1 and 3
// through the third buffer, which is free while a resident visit runs.  Half the record traffic (static block + lambdas,
// ~26 of the ~56 KB a visit moves through L2) disappears; what a visit costs follows those bytes (the lambda write-back
// alone measured 3 % of the frame).  CTAs with another visit count fall back to the two streaming buffers.
template <int LANES, bool TAGGED, bool FAST, bool TETLAM = true, bool RES = false>
__global__ void __launch_bounds__(512, 1) tile_frame_kernel(const TileParams P) {
  static_assert(!RES || TAGGED, "resident record blocks ride on the tagged write-back path");
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long mbar[3];
  __shared__ TileCopy itemCopy[kItemCopySmem];   // copy descriptors of this CTA's tiles: no global latency when prefetching
  __shared__ uint32_t itemTile[kMaxItems];
  __shared__ uint32_t nItemsS;
  __shared__ float4* posPeerS[kMaxRanks];
  __shared__ unsigned* donePeerS[kMaxRanks];
  if (threadIdx.x < kMaxRanks) { posPeerS[threadIdx.x] = P.posPeers[threadIdx.x]; donePeerS[threadIdx.x] = P.donePeers[threadIdx.x]; }
  const bool multi = P.world > 1;
  const uint32_t svOff = (RES ? 3u : 2u) * P.recStride;
  float4* const sv = reinterpret_cast<float4*>(smem + svOff);

  const StepConsts k = *P.consts;
  const uint32_t tid = threadIdx.x, nth = blockDim.x;
  const bool sweeping = P.iterations > 0 && P.nPhases > 0;
  constexpr bool tetLam = TETLAM;
  const bool clamp = P.iterations > 0;   // the reference clamps once per iteration (Sim.cpp:296)

  if (!sweeping) {
    // vertex-local work only: a vertex always belongs to the same CTA, no grid barrier needed
    const uint32_t frameTag = 2u * (P.iterBase * P.nPhases) + 1u;   // nothing advances: keep the frame-start tag
    for (uint32_t sub = 0; sub < P.substeps; ++sub)
      vertex_pass<TAGGED>(P, k, sub == 0 ? LOAD_PREDICT : LOAD_COMMIT_PREDICT, clamp, false, 0u, frameTag);
    vertex_pass<TAGGED>(P, k, LOAD_PLAIN, clamp, true, 0u, frameTag);
    return;
  }

  // the tiles this CTA visits in one iteration, in order
  if (tid == 0) {
    uint32_t n = 0;
    for (uint32_t ph = 0; ph < P.nPhases; ++ph) {
      const PhaseDesc pd = P.phases[ph];
      for (uint32_t t = blockIdx.x; t < pd.tileCount; t += gridDim.x)
        if (n < kMaxItems) {
          const uint32_t tl = P.tileList[pd.tileBegin + t];
          itemTile[n] = tl;
          if (n < kItemCopySmem) itemCopy[n] = P.copies[tl];
          ++n;
        }
    }
    nItemsS = n;
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    mbar_init(&mbar[2], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
  }
  __syncthreads();
  const uint32_t nItems = nItemsS;
  const uint32_t totalItems = nItems * P.iterations * P.substeps;
  // CTAs that share an SM start every phase in lockstep (same barrier, similar tiles) and would
  // hit the shared-memory pipe and the FP32 pipe at the same moments; delaying every second one by
  // about half a colour step lets one CTA's gathers overlap the other's arithmetic.
  __shared__ uint32_t staggerS;
  if (tid == 0) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    staggerS = (P.stagger && (atomicAdd(P.barrier + 16 + (smid & 255u), 1u) & 1u)) ? P.stagger : 0u;
  }
  __syncthreads();
  const uint32_t stagger = staggerS;
  // per-stage accounting for pbd_step_stats (a16): thread 0 brackets the vertex stages that carry the reference's
  // predict / commit work with the cycle counter, in shared memory (no registers held across the frame)
  __shared__ uint32_t stageS[4];
  if (tid == 0) { stageS[0] = stageS[1] = stageS[2] = 0u; stageS[3] = (uint32_t)clock64(); }

  // fetch the record block of this CTA's item `ji` into buffer `b` (thread 0 only)
  auto fetch = [&](uint32_t ji, uint32_t b) {
    const TileCopy c = ji < kItemCopySmem ? itemCopy[ji] : P.copies[itemTile[ji]];
    unsigned char* dst = smem + b * P.recStride;
    const uint32_t tetBytes = tetLam ? c.tetLamBytes : 0u;
    mbar_expect_tx(&mbar[b], c.staticBytes + c.edgeLamBytes + tetBytes);
    bulk_load(dst, P.blob + c.blobOff, c.staticBytes, &mbar[b]);
    if (c.edgeLamBytes) bulk_load(dst + c.staticBytes, P.edgeLam + c.edgeDevBegin, c.edgeLamBytes, &mbar[b]);
    if (tetBytes) bulk_load(dst + c.staticBytes + c.edgeLamBytes, P.tetLam + c.tetDevBegin, tetBytes, &mbar[b]);
  };

  const bool res = RES && nItems == 4;   // visits 0, 2: buffers 0, 1 (resident); visits 1, 3: buffer 2
  if (tid == 0 && nItems) { fetch(0, 0); if (res) fetch(2, 1); }

  unsigned epoch = 0;
  uint32_t item = 0;      // items processed so far by this CTA
  uint32_t buf = 0;
  uint32_t parityBits = 0;   // bit b: phase parity of mbar[b]
  bool needWait = true;
  bool recReady = false;  // tagged path: the previous visit already waited for this visit's record block
  uint32_t j = 0;         // position in itemTile

  for (uint32_t sub = 0; sub < P.substeps; ++sub) {
    for (uint32_t it = 0; it < P.iterations; ++it) {
      for (uint32_t ph = 0; ph < P.nPhases; ++ph) {
        const PhaseDesc pd = P.phases[ph];
        const int mode = ph != 0 ? LOAD_PLAIN : it != 0 ? LOAD_GROUND : sub != 0 ? LOAD_COMMIT_PREDICT : LOAD_PREDICT;
        const bool tr = P.trace && sub == 0 && it + 1 == P.iterations && tid == 0;
        long long* ft = (P.ftrace && tr && blockIdx.x == 0) ? P.ftrace + 256 * ph : nullptr;
        if (tr) P.trace[2 * ((size_t)ph * gridDim.x + blockIdx.x)] = globaltimer_ns();
        for (uint32_t t = blockIdx.x; t < pd.tileCount; t += gridDim.x) {
          if (ft) ft[0] = clock64();
          // ---- record block
          if (needWait && !recReady) {
            // one thread waits on the mbarrier, the block barrier releases the rest (16 warps
            // polling the same mbarrier serialise: ~45 cycles each)
            if (tid == 0) while (!mbar_try_wait(&mbar[buf], (parityBits >> buf) & 1u)) {}
            __syncthreads();
            parityBits ^= 1u << buf;
          }
          recReady = false;
          const uint32_t recOff = buf * P.recStride;
          unsigned char* rec = smem + recOff;
          const TileHdr h = *reinterpret_cast<const TileHdr*>(rec);
          const bool contiguous = (h.flags & 1u) != 0u;
          const bool sysScope = multi && (h.flags & 2u) != 0u;   // this tile exchanges data with another GPU
          // tagged hand-over: what this visit expects in its vertices and what it leaves in them
          const uint32_t seq = (P.iterBase + sub * P.iterations + it) * P.nPhases + ph;
          const uint32_t expectTag = (sub == 0 && it == 0 && ph == 0) ? 2u * seq + 1u : 2u * seq;
          const uint32_t writeTag = 2u * seq + 2u;
          if (ft) ft[1] = clock64();
          // ---- prefetch the next tile's block into the other buffer
          const uint32_t jn = (j + 1 == nItems) ? 0u : j + 1;
          const bool hasNext = item + 1 < totalItems;
          const uint32_t nb = res ? ((jn & 1u) ? 2u : (jn >> 1)) : (buf ^ 1u);   // the next visit's buffer
          // ---- vertices L2 -> shared memory (all of a thread's loads are issued before the first use)
          // ---- point-to-point sync: wait until the tiles that last wrote my vertices have stored them
          if (!TAGGED && P.done) {
            const uint32_t np = (h.flags >> 8) & 0xffu;
            if (tid < np) {
              // entry: tile (bits 0..23) | owner rank (24..27) | bit 31 = written earlier in THIS iteration
              const uint32_t e = reinterpret_cast<const uint32_t*>(rec + 64)[tid];
              const uint32_t need = P.iterBase + sub * P.iterations + it + (e >> 31);
              const uint32_t own = (e >> 24) & 0xfu;
              const unsigned* f = donePeerS[own] + (e & 0xffffffu);
              SpinGuard g;
              if (multi && own != P.rank) { while (!reached(ld_acquire_sys(f), need) && !spin_expired(g, P)) {} }
              else { while (!reached(ld_acquire(f), need) && !spin_expired(g, P)) {} }
            }
            __syncthreads();
          }
          if (ft) ft[14] = clock64();
          if (tid == 0 && mode != LOAD_PLAIN && mode != LOAD_GROUND) stageS[mode == LOAD_PREDICT ? 0 : 1] -= (uint32_t)clock64();
          if (TAGGED) {
            // the values ARE the synchronisation: issue every load of the batch, then wait only for
            // the vertices whose tag is not there yet
            const uint32_t* vidx = reinterpret_cast<const uint32_t*>(smem + recOff + h.offVertIdx);
            for (uint32_t base = 0; base < h.vertCount; base += 3u * nth) {
              uint4 raw[3];
              float w[3];
              VertexIn in[3];
              uint32_t slot[3];
              const uint4* word[3];
#pragma unroll
              for (int u = 0; u < 3; ++u) {
                const uint32_t i = base + u * nth + tid;
                if (i < h.vertCount) {
                  // every load is LOCAL: whoever wrote the vertex last stored it into this rank's array (push model;
                  // the rank bits of a gathered-slot entry name the destination of this tile's own store)
                  slot[u] = contiguous ? h.vertBegin + i : (vidx[i] & 0x0fffffffu);
                  word[u] = P.posT + slot[u];
                  raw[u] = ld_tagged(word[u], sysScope);
                  w[u] = __ldg(P.invMass + slot[u]);
                  in[u].x = in[u].v = make_float4(0.f, 0.f, 0.f, 0.f);
                  if (mode == LOAD_PREDICT || mode == LOAD_COMMIT_PREDICT) in[u].x = __ldcg(P.prev + slot[u]);
                  if (mode == LOAD_PREDICT) in[u].v = __ldcg(P.vel + slot[u]);
                }
              }
#pragma unroll
              for (int u = 0; u < 3; ++u) {
                const uint32_t i = base + u * nth + tid;
                if (i < h.vertCount) {
                  in[u].p = raw[u].w == expectTag ? tagged_value(raw[u], w[u]) : tagged_wait_load(P, word[u], expectTag, w[u], sysScope);
                  sv[i] = finish_vertex(P, k, slot[u], mode, true, in[u]);   // (mode != LOAD_PLAIN only on contiguous home tiles)
                }
              }
            }
          } else if (contiguous) {
            for (uint32_t base = 0; base < h.vertCount; base += 3u * nth) {
              VertexIn in[3];
#pragma unroll
              for (int u = 0; u < 3; ++u) {
                const uint32_t i = base + u * nth + tid;
                if (i < h.vertCount) in[u] = load_vertex(P, h.vertBegin + i, mode);
              }
#pragma unroll
              for (int u = 0; u < 3; ++u) {
                const uint32_t i = base + u * nth + tid;
                if (i < h.vertCount) sv[i] = finish_vertex(P, k, h.vertBegin + i, mode, true, in[u]);
              }
            }
          } else {
            const uint32_t* vidx = reinterpret_cast<const uint32_t*>(smem + recOff + h.offVertIdx);
            for (uint32_t base = 0; base < h.vertCount; base += 4u * nth) {
              float4 v[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const uint32_t i = base + u * nth + tid;
                if (i < h.vertCount) { const uint32_t e = vidx[i]; v[u] = __ldcg(posPeerS[e >> 28] + (e & 0x0fffffffu)); }
              }
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const uint32_t i = base + u * nth + tid;
                if (i < h.vertCount) sv[i] = v[u];
              }
            }
          }
          __syncthreads();
          if (tid == 0 && mode != LOAD_PLAIN && mode != LOAD_GROUND) stageS[mode == LOAD_PREDICT ? 0 : 1] += (uint32_t)clock64();
          if (ft) ft[2] = clock64();
          // ---- prefetch the next tile's record block into the other buffer.  Issued here, behind the
          // barrier that ends the vertex load: every thread has finished the PREVIOUS visit's write-back
          // (which reads that buffer's gathered-slot list) by then, and the lambda write-back that read
          // it was issued a whole vertex load ago.  Its global writes need to be complete only before the
          // same tile's lambdas are fetched again, nItems - 1 visits later.
          // (resident mode: only a resident visit prefetches -- the streamed block of the next visit; the block after a
          // streamed visit is resident)
          if (tid == 0 && hasNext && nItems > 1 && !(res && (j & 1u))) {
            bulk_wait_read();
            if (nItems >= 4) bulk_wait_pending<2>(); else if (nItems == 3) bulk_wait_pending<1>(); else bulk_wait_all();
            fetch(jn, nb);
          }
          // ---- sweeps
          if (stagger) {
            const long long t0 = clock64();
            while (clock64() - t0 < (long long)stagger) {}
          }
          if (LANES == 1 && (h.flags & 4u)) {
            sweep_mixed<FAST>(h, recOff, svOff, k.alphaEdge, k.alphaTet, ft, tetLam);   // edges and tets share the colour steps
            if (ft) ft[3] = clock64();
          } else {
            sweep_edges<FAST>(h, recOff, svOff, k.alphaEdge, ft);
            if (ft) ft[3] = clock64();
            sweep_tets<LANES, FAST>(h, recOff, svOff, k.alphaTet, ft ? ft + 20 : nullptr, k.alphaEdge, tetLam);
          }
          if (ft) ft[4] = clock64();
          // ---- write back
          if (TAGGED && nItems > 1) {
            // Tagged hand-over: nothing to publish after the stores, so the visit ends WITHOUT a block
            // barrier.  One barrier right after the sweeps (a) makes the lambdas visible to the bulk
            // store and (b) hands every thread the NEXT record block, which thread 0 waits for here --
            // it was requested a whole sweep ago.  A thread writes back exactly the shared-memory
            // entries (i = tid mod block size) that it overwrites itself when it loads the next tile.
            fence_async_smem();
            // a resident block is waited for once, before its first visit
            const bool waitNext = hasNext && (!res || (jn & 1u) || item + 1 < nItems);
            if (tid == 0 && waitNext) while (!mbar_try_wait(&mbar[nb], (parityBits >> nb) & 1u)) {}
            __syncthreads();
            if (waitNext) parityBits ^= 1u << nb;
            if (hasNext) recReady = true;
            if (ft) ft[11] = ft[12] = ft[13] = clock64();
            // (Writing the lambdas back with plain 16-byte stores by every thread instead was measured: no gain --
            // what the write-back costs is its traffic, not thread 0's issue time.)
            if (tid == 0) {
              // resident block: its lambdas leave shared memory once, after the tile's last visit of the frame; every
              // visit commits a (possibly empty) group, so the prefetch's wait_group counts stay what they were
              if (!res || (j & 1u) || item + nItems >= totalItems) {
                const TileCopy c = j < kItemCopySmem ? itemCopy[j] : P.copies[itemTile[j]];
                if (c.edgeLamBytes) bulk_store(P.edgeLam + c.edgeDevBegin, rec + h.offEdgeLam, c.edgeLamBytes);
                if (c.tetLamBytes && tetLam) bulk_store(P.tetLam + c.tetDevBegin, rec + h.offTetLam, c.tetLamBytes);
              }
              bulk_commit();
            }
            const uint32_t* vidx = reinterpret_cast<const uint32_t*>(smem + recOff + h.offVertIdx);
            for (uint32_t i = tid; i < h.vertCount; i += nth) {
              const uint32_t e = contiguous ? 0u : vidx[i];
              tagged_store((contiguous ? P.posT + h.vertBegin + i : reinterpret_cast<uint4*>(posPeerS[e >> 28]) + (e & 0x0fffffffu)), sv[i], writeTag, sysScope);
            }
            buf = nb;
            j = jn;
            ++item;
            if (ft) { ft[5] = clock64(); ft[6] = h.nEdgeGroups; ft[7] = h.nTetGroups; ft[8] = h.vertCount; ft[9] = h.nEdges; ft[10] = h.nTets; }
            continue;
          }
          if (TAGGED) {
            const uint32_t* vidx = reinterpret_cast<const uint32_t*>(smem + recOff + h.offVertIdx);
            for (uint32_t i = tid; i < h.vertCount; i += nth) {
              const uint32_t e = contiguous ? 0u : vidx[i];
              tagged_store((contiguous ? P.posT + h.vertBegin + i : reinterpret_cast<uint4*>(posPeerS[e >> 28]) + (e & 0x0fffffffu)), sv[i], writeTag, sysScope);
            }
          } else if (contiguous) {
            for (uint32_t i = tid; i < h.vertCount; i += nth) __stcg(P.pos + h.vertBegin + i, sv[i]);
          } else {
            const uint32_t* vidx = reinterpret_cast<const uint32_t*>(smem + recOff + h.offVertIdx);
            for (uint32_t i = tid; i < h.vertCount; i += nth) { const uint32_t e = vidx[i]; __stcg(posPeerS[e >> 28] + (e & 0x0fffffffu), sv[i]); }
          }
          if (ft) ft[11] = clock64();
          fence_async_smem();   // lambdas written by the sweeps -> visible to the bulk store
          __syncthreads();      // also: sv and rec are free for the next tile
          if (ft) ft[12] = clock64();
          if (tid == 0) {
            // publish first: the release fence waits for the writes issued before it, and nobody but
            // this tile reads its lambdas -- their write-back need not be part of that wait
            if (!TAGGED && P.done) {   // my vertices are in L2 (of their owners)
              const uint32_t v = P.iterBase + sub * P.iterations + it + 1u;
              // system scope only when another GPU reads this counter or holds vertices this tile
              // wrote (flags bit 1): a system-scope release costs several microseconds
              if (multi && (h.flags & 2u)) st_release_sys(P.done + itemTile[j], v); else st_release(P.done + itemTile[j], v);
            }
            if (ft) ft[13] = clock64();
            const TileCopy c = j < kItemCopySmem ? itemCopy[j] : P.copies[itemTile[j]];
            if (c.edgeLamBytes) bulk_store(P.edgeLam + c.edgeDevBegin, rec + h.offEdgeLam, c.edgeLamBytes);
            if (c.tetLamBytes && tetLam) bulk_store(P.tetLam + c.tetDevBegin, rec + h.offTetLam, c.tetLamBytes);
            bulk_commit();
            if (nItems == 1) bulk_wait_read();   // the block stays resident and is swept again next iteration
          }
          if (nItems == 1) {
            __syncthreads();
            needWait = false;
          } else {
            buf ^= 1u;
          }
          j = jn;
          ++item;
          if (ft) { ft[5] = clock64(); ft[6] = h.nEdgeGroups; ft[7] = h.nTetGroups; ft[8] = h.vertCount; ft[9] = h.nEdges; ft[10] = h.nTets; }
        }
        if (tr) P.trace[2 * ((size_t)ph * gridDim.x + blockIdx.x) + 1] = globaltimer_ns();
        if (!TAGGED && !P.done) grid_barrier(P, P.barrier, epoch);
      }
    }
  }
  if (tid == 0) bulk_wait_all();
  auto stage_report = [&](uint32_t tCommit) {   // thread 0, after the final commit pass
    const uint32_t now = (uint32_t)clock64();
    uint32_t* o = P.stageHost + 4u * blockIdx.x;
    o[0] = stageS[0]; o[1] = stageS[1]; o[2] = now - tCommit; o[3] = now - stageS[3];
  };
  uint32_t tCommit = 0;
  if (tid == 0) tCommit = (uint32_t)clock64();
  if (TAGGED) {
    // the final commit waits for the tag of the frame's last visit in every vertex it commits
    const uint32_t seqEnd = (P.iterBase + P.substeps * P.iterations) * P.nPhases;
    vertex_pass<true>(P, k, LOAD_PLAIN, clamp, true, 2u * seqEnd, 2u * seqEnd + 1u);
    if (tid == 0 && P.stageHost) stage_report(tCommit);
    return;
  }
  if (P.done) {
    // the final commit of a home tile reads its vertices' last values: wait for the tiles that wrote them
    const uint32_t need = P.iterBase + P.substeps * P.iterations;
    for (uint32_t q = blockIdx.x; q < P.nHome; q += gridDim.x) {
      const TileCopy c = P.copies[P.tileList[P.phases[0].tileBegin + q]];   // the q-th home tile I own = my q-th phase-0 tile
      const uint32_t* hdr = reinterpret_cast<const uint32_t*>(P.blob + c.blobOff);
      const uint32_t np = (__ldg(hdr + 1) >> 8) & 0xffu;
      if (tid < np) {
        const uint32_t e = __ldg(hdr + 16 + tid);
        const uint32_t own = (e >> 24) & 0xfu;
        const unsigned* f = donePeerS[own] + (e & 0xffffffu);
        SpinGuard g;
        if (multi && own != P.rank) { while (!reached(ld_acquire_sys(f), need) && !spin_expired(g, P)) {} }
        else { while (!reached(ld_acquire(f), need) && !spin_expired(g, P)) {} }
      }
    }
    __syncthreads();
  }
  if (tid == 0) tCommit = (uint32_t)clock64();
  vertex_pass<false>(P, k, LOAD_PLAIN, clamp, true);
  if (tid == 0 && P.stageHost) stage_report(tCommit);
}

// float4 positions <-> tagged words (upload / host reads only)
__global__ void to_tagged_kernel(const float4* pos, uint4* posT, float* invMass, uint32_t n, uint32_t tag) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const float4 p = pos[i]; st_tagged(posT + i, p, tag, false); invMass[i] = p.w; }
}
__global__ void from_tagged_kernel(const uint4* posT, float4* pos, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) pos[i] = tagged_value(ld_tagged(posT + i, true), pos[i].w);   // pos keeps the inverse mass it was uploaded with
}

class TileBackend final : public Backend {
 public:
  TileBackend(const pbd_options& o, int device) : opts_(o), device_(device) {}
  ~TileBackend() override {
    for (void* q : ipcOpened_) cudaIpcCloseMemHandle(q);
    cudaFree(done_); cudaFree(tileList_); cudaFree(homeList_);
    cudaFree(blob_); cudaFree(copies_); cudaFree(phases_); cudaFree(tile0Begin_); cudaFree(barrier_);
    cudaFree(trace_); cudaFree(ftrace_); cudaFree(posT_); cudaFree(invMass_);
    if (abortHost_) cudaFreeHost(abortHost_);
    if (stageHost_) cudaFreeHost(stageHost_);
  }
  const char* name() const override {
    return tagged_ ? (fast_ ? "b200-tile-tagged-fast" : "b200-tile-tagged") : (fast_ ? "b200-tile-fast" : "b200-tile");
  }

  cudaError_t upload(const Plan& plan, const MeshView& m, DeviceArrays& d) override {
    (void)m;
    cudaError_t err;
    block_ = plan.blockThreads ? std::min(plan.blockThreads, 512u) : 512u;   // colour groups were cut to fit one pass of this block
    nPhases_ = (uint32_t)plan.phases.size();
    nTile0_ = (uint32_t)plan.tile0Begin.size() - 1;
    tilesPerSm_ = std::max(1u, plan.tilesPerSm);
    lanes_ = opts_.lanes_per_tet == 2 ? 2u : opts_.lanes_per_tet == 4 ? 4u : 1u;   // auto: one thread per tet (fastest measured)
    fast_ = (opts_.flags & PBD_FLAG_FAST_ARITH) != 0u;
    if (fast_) lanes_ = 1;   // the fast forms exist for one thread per constraint

    // ---- tagged hand-over: one thread per tet, and every phase must cover every vertex exactly once
    // (then a vertex's previous writer is always the previous visit)
    tagged_ = false;
    if ((opts_.flags & PBD_FLAG_TAGGED_HANDOVER) && lanes_ == 1 && !plan.phases.empty()) {
      bool full = true;
      for (const Phase& ph : plan.phases) {
        uint64_t covered = 0;
        for (uint32_t ti = ph.tileBegin; ti < ph.tileBegin + ph.tileCount; ++ti) covered += plan.tiles[ti].vertCount;
        full &= covered == plan.V;   // tiles of one phase are vertex-disjoint
      }
      tagged_ = full;
    }

    // rest values are already on the device at the plan's device indices (pbd_capi.cu)
    std::vector<float> eRest(plan.edgeDevCount), tRest(plan.tetDevCount);
    if (plan.edgeDevCount && (err = cudaMemcpy(eRest.data(), d.edgeRest, sizeof(float) * plan.edgeDevCount, cudaMemcpyDeviceToHost)) != cudaSuccess) return err;
    if (plan.tetDevCount && (err = cudaMemcpy(tRest.data(), d.tetRest, sizeof(float) * plan.tetDevCount, cudaMemcpyDeviceToHost)) != cudaSuccess) return err;

    // ---- ownership (one body across `world` GPUs; world == 1: everything is mine).  Home tiles are
    // dealt out in contiguous runs (they are slot-contiguous and spatially ordered), a vertex belongs
    // to the rank of its home tile, any other tile to the rank that owns most of its vertices.
    world_ = std::max(1u, opts_.shard_world);
    rank_ = opts_.shard_rank;
    if (world_ > kMaxRanks || rank_ >= world_ || plan.V >= (1u << 28) || plan.tiles.size() >= (1u << 24)) return cudaErrorInvalidValue;
    std::vector<uint32_t> rankSlotBegin(world_ + 1, plan.V);
    rankSlotBegin[0] = 0;
    homeOwner_.assign(nTile0_, 0);
    for (uint32_t t = 0; t < nTile0_; ++t) {
      const uint32_t r = (uint32_t)(((uint64_t)t * world_) / std::max(1u, nTile0_));
      homeOwner_[t] = (uint8_t)r;
      rankSlotBegin[r] = std::min(rankSlotBegin[r], plan.tile0Begin[t]);
    }
    for (uint32_t r = world_; r-- > 0;) rankSlotBegin[r] = std::min(rankSlotBegin[r], rankSlotBegin[r + 1]);
    auto owner_of_slot = [&](uint32_t s) {
      uint32_t r = 0;
      while (r + 1 < world_ && s >= rankSlotBegin[r + 1]) ++r;
      return r;
    };
    slotOwnerBegin_ = rankSlotBegin;
    std::vector<uint8_t> tileOwner(plan.tiles.size(), 0);
    for (size_t ti = 0; ti < plan.tiles.size(); ++ti) {
      const Tile& t = plan.tiles[ti];
      if (world_ == 1 || t.vertCount == 0) continue;
      if (t.contiguous) { tileOwner[ti] = (uint8_t)owner_of_slot(t.vertBegin); continue; }
      uint32_t cnt[kMaxRanks] = {0};
      for (uint32_t i = 0; i < t.vertCount; ++i) cnt[owner_of_slot(plan.tileVerts[t.vertBegin + i])]++;
      uint32_t best = 0;
      for (uint32_t r = 1; r < world_; ++r)
        if (cnt[r] > cnt[best]) best = r;
      tileOwner[ti] = (uint8_t)best;
    }

    // ---- point-to-point dependencies: for every tile, the tiles that last wrote one of its vertices
    // (entry: tile | owner rank << 24 | bit 31 = earlier in the same iteration, else previous iteration)
    std::vector<std::vector<uint32_t>> preds(plan.tiles.size());
    bool flagsOk = !getenv("PBD_TILE_GRIDSYNC");
    {
      std::vector<uint32_t> lastTile(plan.V, 0xffffffffu), lastPhase(plan.V, 0);
      auto for_verts = [&](const Tile& t, auto&& fn) {
        for (uint32_t i = 0; i < t.vertCount; ++i) fn(t.contiguous ? t.vertBegin + i : plan.tileVerts[t.vertBegin + i]);
      };
      for (int pass = 0; pass < 2; ++pass)
        for (size_t ph = 0; ph < plan.phases.size(); ++ph)
          for (uint32_t ti = plan.phases[ph].tileBegin; ti < plan.phases[ph].tileBegin + plan.phases[ph].tileCount; ++ti) {
            const Tile& t = plan.tiles[ti];
            if (pass == 1) {
              std::vector<uint32_t>& pr = preds[ti];
              for_verts(t, [&](uint32_t s) {
                if (lastTile[s] != 0xffffffffu && lastTile[s] != ti)
                  pr.push_back(lastTile[s] | ((uint32_t)tileOwner[lastTile[s]] << 24) | (lastPhase[s] < ph ? 0x80000000u : 0u));
              });
              std::sort(pr.begin(), pr.end());
              pr.erase(std::unique(pr.begin(), pr.end()), pr.end());
              if (pr.size() > kMaxPreds) flagsOk = false;
            }
            for_verts(t, [&](uint32_t s) { lastTile[s] = ti; lastPhase[s] = (uint32_t)ph; });
          }
    }
    useFlags_ = flagsOk;
    if (world_ > 1 && !useFlags_) return cudaErrorNotSupported;   // several GPUs cannot share a grid barrier

    // A tile that talks to another GPU (a predecessor or a vertex there) goes to the front of its
    // phase's list, i.e. into the first wave of CTAs: its peers on the other rank are first in line
    // too, so the NVLink round trips overlap with the rank's interior tiles instead of ending the phase.
    std::vector<uint8_t> remote(plan.tiles.size(), 0);
    if (world_ > 1)
      for (size_t ti = 0; ti < plan.tiles.size(); ++ti) {
        const Tile& t = plan.tiles[ti];
        for (uint32_t e : preds[ti]) remote[ti] |= ((e >> 24) & 0xfu) != rank_;
        if (!t.contiguous)
          for (uint32_t i = 0; i < t.vertCount && !remote[ti]; ++i) remote[ti] |= owner_of_slot(plan.tileVerts[t.vertBegin + i]) != rank_;
      }
    for (uint32_t ti = 0; ti < plan.tiles.size(); ++ti)   // ... and so does a tile another rank waits for
      for (uint32_t e : preds[ti])
        if (tileOwner[ti] != tileOwner[e & 0xffffffu]) remote[e & 0xffffffu] = 1;

    // ---- record blocks
    std::vector<TileCopy> copies(plan.tiles.size());
    std::vector<unsigned char> blob;
    uint32_t recMax = 64, recMaxNoTetLam = 64;
    // Tagged hand-over across GPUs: PUSH model.  A tile stores each vertex into the memory of the rank
    // that reads it NEXT (the owner of the tile that holds the vertex in the following phase), so every
    // load -- and every retry while a tag is not there yet -- is local; only posted 16-byte stores cross
    // NVLink (a remote load is a ~2 us round trip on the consumer's critical path, a remote store is not
    // waited for by anyone).  The rank bits of a gathered-slot entry then name the DESTINATION rank of the
    // store, and home tiles carry a slot list too (their vertices scatter to several ranks).
    const bool pushModel = tagged_ && world_ > 1;
    std::vector<uint8_t> destRank;   // [phase * V + slot]
    if (pushModel) {
      const size_t nPh = plan.phases.size();
      std::vector<uint32_t> tileIn(nPh * (size_t)plan.V, 0);
      for (size_t ph = 0; ph < nPh; ++ph)
        for (uint32_t ti = plan.phases[ph].tileBegin; ti < plan.phases[ph].tileBegin + plan.phases[ph].tileCount; ++ti) {
          const Tile& t = plan.tiles[ti];
          for (uint32_t i = 0; i < t.vertCount; ++i)
            tileIn[ph * plan.V + (t.contiguous ? t.vertBegin + i : plan.tileVerts[t.vertBegin + i])] = ti;
        }
      destRank.resize(nPh * (size_t)plan.V);
      for (size_t ph = 0; ph < nPh; ++ph)
        for (uint32_t sl = 0; sl < plan.V; ++sl) destRank[ph * plan.V + sl] = tileOwner[tileIn[((ph + 1) % nPh) * plan.V + sl]];
      // system scope only for the tiles that exchange words with another GPU (flags bit 1): a tile whose
      // vertices were last written by a tile of another rank, or that stores into another rank's memory
      std::fill(remote.begin(), remote.end(), (uint8_t)0);
      for (size_t ph = 0; ph < nPh; ++ph)
        for (uint32_t ti = plan.phases[ph].tileBegin; ti < plan.phases[ph].tileBegin + plan.phases[ph].tileCount; ++ti) {
          const Tile& t = plan.tiles[ti];
          for (uint32_t i = 0; i < t.vertCount && !remote[ti]; ++i) {
            const uint32_t sl = t.contiguous ? t.vertBegin + i : plan.tileVerts[t.vertBegin + i];
            const uint32_t writer = tileIn[((ph + nPh - 1) % nPh) * plan.V + sl];
            remote[ti] = tileOwner[writer] != tileOwner[ti] || destRank[ph * plan.V + sl] != tileOwner[ti];
          }
        }
    }
    std::vector<uint32_t> phaseOfTile(plan.tiles.size(), 0);
    for (size_t ph = 0; ph < plan.phases.size(); ++ph)
      for (uint32_t ti = plan.phases[ph].tileBegin; ti < plan.phases[ph].tileBegin + plan.phases[ph].tileCount; ++ti) phaseOfTile[ti] = (uint32_t)ph;
    for (size_t ti = 0; ti < plan.tiles.size(); ++ti) {
      const Tile& t = plan.tiles[ti];
      const bool asRange = t.contiguous && !pushModel;   // the kernel walks the slot range [vertBegin, +vertCount) itself
      const uint32_t nVG = asRange ? 0u : t.vertCount;
      TileHdr h{};
      const uint32_t nPred = useFlags_ ? (uint32_t)preds[ti].size() : 0u;
      h.vertCount = t.vertCount; h.flags = (asRange ? 1u : 0u) | (remote[ti] ? 2u : 0u) | (t.mixed ? 4u : 0u) | (t.ride ? 8u : 0u) | (nPred << 8); h.vertBegin = asRange ? t.vertBegin : 0u;
      h.nEdgeGroups = t.edgeGroupCount; h.nTetGroups = t.tetGroupCount; h.nEdges = t.edgeCount; h.nTets = t.tetCount;
      // planner invariants the sweeps rely on (pbd_sweep.cuh projects a colour group in ONE pass of the
      // block and would silently drop the rest): fail loudly instead
      if (t.mixed && t.edgeGroupCount != t.tetGroupCount) return cudaErrorInvalidConfiguration;
      for (uint32_t g = 0; g < t.edgeGroupCount; ++g) {
        const uint32_t ne = plan.groups[t.edgeGroupBegin + g].count;
        const uint32_t nt = t.mixed ? plan.groups[t.tetGroupBegin + g].count : 0u;
        if (ne + nt > block_) return cudaErrorInvalidConfiguration;
      }
      for (uint32_t g = 0; g < t.tetGroupCount; ++g)
        if (plan.groups[t.tetGroupBegin + g].count * lanes_ > block_) return cudaErrorInvalidConfiguration;
      uint32_t off = 64 + 4u * kMaxPreds;
      h.offVertIdx = off; off += 4u * pad4(nVG);
      h.offEdgeGroups = off; off += 8u * (pad4(t.edgeGroupCount * 2) / 2);
      h.offTetGroups = off; off += 8u * (pad4(t.tetGroupCount * 2) / 2);
      h.offEdgeIdx = off; off += 4u * pad4(t.edgeCount);
      h.offEdgeRest = off; off += 4u * pad4(t.edgeCount);
      h.offTetIdx = off; off += 8u * (pad4(t.tetCount * 2) / 2);
      h.offTetRest = off; off += 4u * pad4(t.tetCount);
      const uint32_t offTetRide = off;   // PBD_ORDER_RIDING: u32 per tet, right behind the rest values (the kernel derives the offset)
      if (t.ride) off += 4u * pad4(t.tetCount);
      const uint32_t staticBytes = off;
      h.offEdgeLam = off; off += 4u * pad4(t.edgeCount);
      h.offTetLam = off; off += 4u * pad4(t.tetCount);
      if (staticBytes != tile_static_bytes(nVG, t.edgeGroupCount, t.tetGroupCount, t.edgeCount, t.tetCount, t.ride != 0)) return cudaErrorUnknown;
      if (t.ride && (t.edgeCount >= 0xffffu || plan.tetRide.size() != 2 * (size_t)plan.T)) return cudaErrorInvalidConfiguration;
      recMax = std::max(recMax, off);
      recMaxNoTetLam = std::max(recMaxNoTetLam, h.offTetLam);

      const size_t base = blob.size();
      blob.resize(base + staticBytes, 0);
      unsigned char* b = blob.data() + base;
      memcpy(b, &h, sizeof(h));
      if (nPred) memcpy(b + 64, preds[ti].data(), 4u * nPred);
      if (nVG) {
        uint32_t* vi = reinterpret_cast<uint32_t*>(b + h.offVertIdx);
        for (uint32_t q = 0; q < nVG; ++q) {
          const uint32_t sl = t.contiguous ? t.vertBegin + q : plan.tileVerts[t.vertBegin + q];
          vi[q] = sl | ((pushModel ? (uint32_t)destRank[(size_t)phaseOfTile[ti] * plan.V + sl] : owner_of_slot(sl)) << 28);
        }
      }
      uint32_t* eg = reinterpret_cast<uint32_t*>(b + h.offEdgeGroups);
      for (uint32_t g = 0; g < t.edgeGroupCount; ++g) {
        eg[2 * g] = plan.groups[t.edgeGroupBegin + g].begin - t.edgeBegin;
        eg[2 * g + 1] = plan.groups[t.edgeGroupBegin + g].count;
      }
      uint32_t* tg = reinterpret_cast<uint32_t*>(b + h.offTetGroups);
      for (uint32_t g = 0; g < t.tetGroupCount; ++g) {
        tg[2 * g] = plan.groups[t.tetGroupBegin + g].begin - t.tetBegin;
        tg[2 * g + 1] = plan.groups[t.tetGroupBegin + g].count;
      }
      if (t.mixed) {
        // mixed steps read ONE table of {edge begin, edge count, tet begin, first tet thread} entries: the two sections
        // are adjacent, equally long (same group count) and together hold exactly 16 bytes per step
        if (h.offTetGroups != h.offEdgeGroups + 8u * (pad4(t.edgeGroupCount * 2) / 2) || t.edgeGroupCount != t.tetGroupCount) return cudaErrorUnknown;
        for (uint32_t g = 0; g < t.edgeGroupCount; ++g) {
          eg[4 * g] = plan.groups[t.edgeGroupBegin + g].begin - t.edgeBegin;
          eg[4 * g + 1] = plan.groups[t.edgeGroupBegin + g].count;
          eg[4 * g + 2] = plan.groups[t.tetGroupBegin + g].begin - t.tetBegin;
          eg[4 * g + 3] = block_ - plan.groups[t.tetGroupBegin + g].count;   // the first tet thread (tets sit on the block's last threads)
        }
      }
      uint32_t* ei = reinterpret_cast<uint32_t*>(b + h.offEdgeIdx);
      float* er = reinterpret_cast<float*>(b + h.offEdgeRest);
      for (uint32_t q = 0; q < t.edgeCount; ++q) {
        const size_t kk = (size_t)t.edgeBegin + q;
        ei[q] = (uint32_t)plan.edgeLocal[2 * kk] | ((uint32_t)plan.edgeLocal[2 * kk + 1] << 16);
        er[q] = eRest[plan.edgeDev[kk]];
      }
      uint32_t* tix = reinterpret_cast<uint32_t*>(b + h.offTetIdx);
      float* trs = reinterpret_cast<float*>(b + h.offTetRest);
      for (uint32_t q = 0; q < t.tetCount; ++q) {
        const size_t kk = (size_t)t.tetBegin + q;
        uint16_t l[4] = {plan.tetLocal[4 * kk], plan.tetLocal[4 * kk + 1], plan.tetLocal[4 * kk + 2], plan.tetLocal[4 * kk + 3]};
        if (fast_ && kRegRiders && t.ride && plan.tetRide[2 * kk] != 0xffffffffu) {
          // fast arithmetic keeps a tet's riders in registers: relabel the tet's vertices by an EVEN
          // permutation (same signed volume, same rest value / lambda) so that rider 0 is the edge
          // (v0, v1) and rider 1 -- the opposite edge, if there is one -- is (v2, v3).  The edge's
          // own orientation does not matter (its correction is antisymmetric).  Exact arithmetic never
          // relabels: the reference's rounding sequence depends on the caller's vertex order.
          const size_t ek = plan.tetRide[2 * kk];
          const uint16_t a = plan.edgeLocal[2 * ek], b = plan.edgeLocal[2 * ek + 1];
          int ia = -1, ib = -1;
          for (int j = 0; j < 4; ++j) { if (l[j] == a && ia < 0) ia = j; else if (l[j] == b && ib < 0) ib = j; }
          if (ia < 0 || ib < 0) return cudaErrorInvalidConfiguration;   // a rider's vertices are its host's
          int rest2[2], nr = 0;
          for (int j = 0; j < 4; ++j) if (j != ia && j != ib) rest2[nr++] = j;
          int perm[4] = {ia, ib, rest2[0], rest2[1]};
          int inv = 0;
          for (int x = 0; x < 4; ++x) for (int y = x + 1; y < 4; ++y) inv += perm[x] > perm[y];
          if (inv & 1) std::swap(perm[2], perm[3]);
          const uint16_t o4[4] = {l[0], l[1], l[2], l[3]};
          for (int j = 0; j < 4; ++j) l[j] = o4[perm[j]];
        }
        tix[2 * q] = (uint32_t)l[0] | ((uint32_t)l[1] << 16);
        tix[2 * q + 1] = (uint32_t)l[2] | ((uint32_t)l[3] << 16);
        // (a tet the planner relabelled by an odd permutation -- fast arithmetic only -- sees its signed volume negated:
        // its record carries the negated rest volume, and its multiplier lives with the opposite sign)
        const bool flipped = !plan.tetPerm.empty() && tet_perm_is_odd(plan.tetPerm[kk]);
        if (flipped && !fast_) return cudaErrorInvalidConfiguration;   // the exact arithmetic follows the reference's evaluation order
        trs[q] = flipped ? -tRest[plan.tetDev[kk]] : tRest[plan.tetDev[kk]];
      }
      if (t.ride) {
        uint32_t* rd = reinterpret_cast<uint32_t*>(b + offTetRide);
        for (uint32_t q = 0; q < t.tetCount; ++q) {
          uint32_t pos[2];
          for (uint32_t sl = 0; sl < 2; ++sl) {
            const uint32_t gp = plan.tetRide[2 * ((size_t)t.tetBegin + q) + sl];   // schedule position of the rider
            if (gp != 0xffffffffu && (gp < t.edgeBegin || gp - t.edgeBegin >= t.edgeCount)) return cudaErrorInvalidConfiguration;
            pos[sl] = gp == 0xffffffffu ? 0xffffu : gp - t.edgeBegin;
          }
          rd[q] = pos[0] | (pos[1] << 16);
        }
      }
      TileCopy& c = copies[ti];
      c.blobOff = base;
      c.staticBytes = staticBytes;
      c.edgeDevBegin = t.edgeDevBegin; c.edgeLamBytes = 4u * pad4(t.edgeCount);
      c.tetDevBegin = t.tetDevBegin; c.tetLamBytes = 4u * pad4(t.tetCount);
      c.pad = 0;
    }
    if (const char* dumpPath = getenv("PBD_DUMP_TILE")) {
      // tools/mb_sweep.cu input: one tile's record block + its vertices (debug aid)
      size_t ti = plan.phases.size() > 1 ? plan.phases[1].tileBegin : 0;
      if (const char* e = getenv("PBD_DUMP_TILE_INDEX")) ti = (size_t)atoi(e);
      if (ti < plan.tiles.size()) {
        const Tile& t = plan.tiles[ti];
        const TileCopy& c = copies[ti];
        std::vector<float4> all(d.V), vv(t.vertCount);
        cudaMemcpy(all.data(), d.pos, sizeof(float4) * d.V, cudaMemcpyDeviceToHost);
        for (uint32_t i = 0; i < t.vertCount; ++i) vv[i] = all[t.contiguous ? t.vertBegin + i : plan.tileVerts[t.vertBegin + i]];
        const uint32_t lamBytes = c.edgeLamBytes + c.tetLamBytes;
        const uint32_t head[6] = {0x54494c45u, c.staticBytes + lamBytes, c.staticBytes, t.vertCount, t.edgeCount, t.tetCount};
        if (FILE* f = fopen(dumpPath, "wb")) {
          fwrite(head, sizeof(head), 1, f);
          fwrite(blob.data() + c.blobOff, 1, c.staticBytes, f);
          std::vector<unsigned char> z(lamBytes, 0);
          fwrite(z.data(), 1, lamBytes, f);
          fwrite(vv.data(), sizeof(float4), vv.size(), f);
          fclose(f);
        }
      }
    }
    recStride_ = (recMax + 127u) & ~127u;
    smemBytes_ = 2 * (size_t)recStride_ + sizeof(float4) * (size_t)std::max(plan.tileVertexCapacity, 1u);
    // the resident-block kernel runs without the tet multipliers: its three buffers end before that section
    recStrideRes_ = (recMaxNoTetLam + 127u) & ~127u;
    smemBytesRes_ = 3 * (size_t)recStrideRes_ + sizeof(float4) * (size_t)std::max(plan.tileVertexCapacity, 1u);

    std::vector<PhaseDesc> pd(plan.phases.size());
    std::vector<uint32_t> tileList, homeList;
    maxTilesPerPhase_ = 0;
    for (size_t i = 0; i < pd.size(); ++i) {
      pd[i].tileBegin = (uint32_t)tileList.size();
      for (int pass = 1; pass >= 0; --pass)
        for (uint32_t ti = plan.phases[i].tileBegin; ti < plan.phases[i].tileBegin + plan.phases[i].tileCount; ++ti)
          if (tileOwner[ti] == rank_ && remote[ti] == pass) tileList.push_back(ti);
      pd[i].tileCount = (uint32_t)tileList.size() - pd[i].tileBegin;
      maxTilesPerPhase_ = std::max(maxTilesPerPhase_, pd[i].tileCount);
    }
    if (!pd.empty()) {
      // phase 0 = my home tiles, in the order just chosen (the final commit walks homeList)
      for (uint32_t q = 0; q < pd[0].tileCount; ++q) homeList.push_back(tileList[pd[0].tileBegin + q] - plan.phases[0].tileBegin);
    } else {
      for (uint32_t t = 0; t < nTile0_; ++t)
        if (homeOwner_[t] == rank_) homeList.push_back(t);
    }
    nHome_ = (uint32_t)homeList.size();
    maxTilesPerPhase_ = std::max(maxTilesPerPhase_, nHome_);
    auto up = [&](auto** dst, const auto& src) -> cudaError_t {
      using T = typename std::remove_reference<decltype(src)>::type::value_type;
      cudaError_t e = cudaMalloc((void**)dst, sizeof(T) * src.size() + 256);
      if (e != cudaSuccess) return e;
      bytes_ += sizeof(T) * src.size();
      if (!src.empty()) e = cudaMemcpy(*dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice);
      return e;
    };
    if ((err = up(&blob_, blob)) != cudaSuccess) return err;
    if ((err = up(&copies_, copies)) != cudaSuccess) return err;
    if ((err = up(&phases_, pd)) != cudaSuccess) return err;
    if ((err = up(&tile0Begin_, plan.tile0Begin)) != cudaSuccess) return err;
    if ((err = up(&tileList_, tileList)) != cudaSuccess) return err;
    if ((err = up(&homeList_, homeList)) != cudaSuccess) return err;
    if ((err = cudaMalloc((void**)&barrier_, 2048)) != cudaSuccess) return err;
    // bounded spins: a mapped host word the kernel sets when a wait gives up (checked by pbd_sync for free)
    if ((err = cudaHostAlloc((void**)&abortHost_, 64, cudaHostAllocMapped)) != cudaSuccess) return err;
    *abortHost_ = 0u;
    if ((err = cudaHostGetDevicePointer((void**)&abortHostDev_, abortHost_, 0)) != cudaSuccess) return err;
    {
      int khz = 0;
      cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device_);
      const char* e = getenv("PBD_SPIN_LIMIT_MS");
      const double ms = e ? atof(e) : 10000.0;
      spinLimit_ = (long long)(ms * (double)std::max(khz, 1000000));
    }
    doneBytes_ = sizeof(unsigned) * (plan.tiles.size() + 1);
    // debug/test hook: start the iteration count (done counters, tags) near the 32-bit wrap
    if (const char* e = getenv("PBD_DEBUG_ITERBASE")) iterBase_ = (uint32_t)strtoull(e, nullptr, 10);
    if (useFlags_) {
      if ((err = cudaMalloc((void**)&done_, doneBytes_)) != cudaSuccess) return err;
      // counters only ever grow (compared by signed distance, so the 32-bit wrap is harmless)
      std::vector<unsigned> init(plan.tiles.size() + 1, iterBase_);
      if ((err = cudaMemcpy(done_, init.data(), doneBytes_, cudaMemcpyHostToDevice)) != cudaSuccess) return err;
    }
    for (uint32_t r = 0; r < kMaxRanks; ++r) { posPeers_[r] = nullptr; donePeers_[r] = nullptr; }
    posPeers_[rank_] = d.pos;
    donePeers_[rank_] = done_;
    attached_ = world_ == 1;
    stagger_ = getenv("PBD_TILE_STAGGER") ? (uint32_t)atoi(getenv("PBD_TILE_STAGGER")) : 0u;
    dropInert_ = getenv("PBD_TILE_KEEP_LAMBDA") ? 0u : 1u;   // debug / A-B: carry inert multipliers in the fast mode too
    if (getenv("PBD_TILE_TRACE")) {
      traceN_ = 2 * (size_t)(nPhases_ + 1) * 4096;
      if ((err = cudaMalloc((void**)&trace_, sizeof(unsigned long long) * traceN_)) != cudaSuccess) return err;
      cudaMemset(trace_, 0, sizeof(unsigned long long) * traceN_);
      if ((err = cudaMalloc((void**)&ftrace_, sizeof(long long) * 256 * (nPhases_ + 1))) != cudaSuccess) return err;
      cudaMemset(ftrace_, 0, sizeof(long long) * 256 * (nPhases_ + 1));
    }

    if (tagged_) {
      if ((err = cudaMalloc((void**)&posT_, sizeof(uint4) * ((size_t)plan.V + 1))) != cudaSuccess) return err;
      if ((err = cudaMalloc((void**)&invMass_, sizeof(float) * ((size_t)plan.V + 1))) != cudaSuccess) return err;
      bytes_ += (sizeof(uint4) + sizeof(float)) * (size_t)plan.V;
      to_tagged_kernel<<<(plan.V + 255) / 256, 256>>>(d.pos, posT_, invMass_, plan.V, 2u * (iterBase_ * nPhases_) + 1u);   // the frame-start tag
      if ((err = cudaGetLastError()) != cudaSuccess) return err;
      posPeers_[rank_] = reinterpret_cast<float4*>(posT_);   // what peers map and address: the tagged words (same 16 B per vertex)
    }

    int perSM = 0, nSM = 0, coop = 0;
    cudaDeviceGetAttribute(&nSM, cudaDevAttrMultiProcessorCount, device_);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device_);
    const uint32_t wantTiles = std::max(1u, opts_.tiles_per_sm ? opts_.tiles_per_sm : tilesPerSm_);
    {
      const void* fn = kernel();
      if ((err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes_)) != cudaSuccess) return err;
      if (kernel(true) != fn && (err = cudaFuncSetAttribute(kernel(true), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes_)) != cudaSuccess) return err;
      if ((err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, fn, (int)block_, smemBytes_)) != cudaSuccess) return err;
    }
    if (!coop || perSM < 1) return cudaErrorCooperativeLaunchTooLarge;
    const uint32_t wantPerSm = std::max(1u, std::min(wantTiles, (uint32_t)perSM));
    grid_ = std::max(1u, std::min(maxTilesPerPhase_, wantPerSm * (uint32_t)nSM));
    // Resident record blocks (the RES kernel; fast arithmetic without inert tet multipliers only -- with them three
    // buffers do not leave room for two CTAs per SM): four phases of at most one tile per CTA, and the same grid must
    // stay co-resident with the third buffer.
    if (tagged_ && fast_ && dropInert_ && world_ == 1 && pd.size() == 4 && maxTilesPerPhase_ <= grid_ && !getenv("PBD_TILE_NORESIDENT")) {
      resident_ = true;
      int perRes = 0;
      const void* fr = kernel(true);
      const bool ok = cudaFuncSetAttribute(fr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytesRes_) == cudaSuccess &&
                      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perRes, fr, (int)block_, smemBytesRes_) == cudaSuccess &&
                      (uint32_t)perRes * (uint32_t)nSM >= grid_;
      if (!ok) { cudaGetLastError(); resident_ = false; }
      if (getenv("PBD_TILE_TRACE")) fprintf(stderr, "[pbd-tile] resident record blocks: %s (3 x %u + %zu vertex bytes per CTA, %d CTAs per SM)\n",
                                            resident_ ? "on" : "off", recStrideRes_, smemBytesRes_ - 3 * (size_t)recStrideRes_, perRes);
    }
    // every CTA's per-iteration tile list must fit the kernel's item table
    uint64_t items = 0;
    for (const PhaseDesc& p : pd) items += (p.tileCount + grid_ - 1) / grid_;
    if (items > kMaxItems) return cudaErrorInvalidConfiguration;
    return cudaSuccess;
  }

  cudaError_t enqueue_frame(const DeviceArrays& d, const FrameShape& f, cudaStream_t s) override {
    TileParams P{};
    P.pos = d.pos; P.posT = tagged_ ? posT_ : nullptr; P.invMass = invMass_; P.prev = d.prev; P.vel = d.vel;
    P.blob = blob_; P.copies = copies_; P.edgeLam = d.edgeLam; P.tetLam = d.tetLam;
    P.phases = phases_; P.tile0Begin = tile0Begin_; P.consts = d.consts; P.barrier = barrier_;
    P.colliders = d.colliders; P.nColliders = d.nColliders;
    P.trace = trace_; P.ftrace = ftrace_;
    P.nTile0 = nTile0_; P.nPhases = nPhases_; P.substeps = f.substeps; P.iterations = f.iterations;
    const bool resKernel = resident_ && f.tetInert;   // == what kernel(f.tetInert) returns
    P.recStride = resKernel ? recStrideRes_ : recStride_;
    P.stagger = stagger_;
    P.done = useFlags_ ? done_ : nullptr;
    if (!attached_) return cudaErrorNotReady;   // pbd_shard_attach_* first
    for (uint32_t r = 0; r < kMaxRanks; ++r) { P.posPeers[r] = posPeers_[r]; P.donePeers[r] = donePeers_[r]; }
    P.tileList = tileList_; P.homeList = homeList_; P.nHome = nHome_; P.world = world_; P.rank = rank_;
    P.iterBase = iterBase_;
    iterBase_ += f.substeps * f.iterations;
    P.spinLimit = spinLimit_; P.abortWord = barrier_ + 64; P.abortHost = abortHostDev_;
    if (!stageHost_) {
      if (cudaHostAlloc((void**)&stageHost_, sizeof(uint32_t) * 4 * kStageCtas, cudaHostAllocMapped) != cudaSuccess ||
          cudaHostGetDevicePointer((void**)&stageHostDev_, stageHost_, 0) != cudaSuccess) return cudaErrorMemoryAllocation;
      std::memset(stageHost_, 0, sizeof(uint32_t) * 4 * kStageCtas);
    }
    P.stageHost = grid_ <= kStageCtas ? stageHostDev_ : nullptr;
    stageGrid_ = grid_;
    cudaError_t err = cudaMemsetAsync(barrier_, 0, 2048, s);
    if (err != cudaSuccess) return err;
    void* args[] = {&P};
    return cudaLaunchCooperativeKernel(kernel(f.tetInert), dim3(grid_), dim3(block_), args, resKernel ? smemBytesRes_ : smemBytes_, s);
  }

  // pbd_step_stats (a16): the frame is ONE kernel, so the stages are shares of its device time -- cycles thread 0 of
  // every CTA spent in the vertex stages that carry the reference's predict (PBDServer.h:75-119 predictMs) and
  // commit work, summed over the CTAs, over the cycles of the whole frame.  A fused commit+predict stage
  // (substeps 2..S) is charged half to each.  Valid after the frame's sync; the last frame's figures.
  bool stage_share(double& predict, double& commit) override {
    if (!stageHost_ || stageGrid_ == 0 || stageGrid_ > kStageCtas) return false;
    double a0 = 0, a1 = 0, a2 = 0, tot = 0;
    for (uint32_t c = 0; c < stageGrid_; ++c) {
      const volatile uint32_t* o = stageHost_ + 4u * c;
      a0 += o[0]; a1 += o[1]; a2 += o[2]; tot += o[3];
    }
    if (tot <= 0) return false;
    predict = (a0 + 0.5 * a1) / tot;
    commit = (0.5 * a1 + a2) / tot;
    return true;
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
    std::vector<long long> f(256 * (size_t)nPhases_);
    cudaMemcpy(f.data(), ftrace_, sizeof(long long) * f.size(), cudaMemcpyDeviceToHost);
    for (uint32_t ph = 0; ph < nPhases_; ++ph) {
      const long long* q = &f[256 * (size_t)ph];
      fprintf(stderr, "[pbd-ftrace] phase %u CTA0: verts %lld edges %lld tets %lld | groups %lld+%lld | wait rec %lld cyc | poll %lld | vertex load %lld | edge sweep %lld | tet sweep %lld | "
                      "store issue %lld | fence+barrier %lld | release %lld | lambda bulk %lld\n",
              ph, q[8], q[9], q[10], q[6], q[7], q[1] - q[0], q[14] - q[1], q[2] - q[14], q[3] - q[2], q[4] - q[3], q[11] - q[4], q[12] - q[11],
              q[13] - q[12], q[5] - q[13]);
      fprintf(stderr, "[pbd-steps] phase %u edges (cycles/size):", ph);
      long long prev = q[2];
      for (int g = 0; g < 20 && g < q[6]; ++g) { fprintf(stderr, " %lld/%lld", q[16 + g] - prev, q[56 + g]); prev = q[16 + g]; }
      fprintf(stderr, " | tets:");
      prev = q[3];
      for (int g = 0; g < 40 && g < q[7]; ++g) { fprintf(stderr, " %lld/%lld", q[36 + g] - prev, q[76 + g]); prev = q[36 + g]; }
      fprintf(stderr, "\n");
    }
  }
  // ---- one body across several GPUs
  uint32_t shard_world() const override { return world_; }
  uint32_t shard_rank() const override { return rank_; }
  void shard_slot_ranges(std::vector<uint32_t>& begin) const override { begin = slotOwnerBegin_; }
  cudaError_t shard_export(void* out128) override {
    cudaIpcMemHandle_t hs[2];
    memset(hs, 0, sizeof(hs));
    cudaError_t e = cudaIpcGetMemHandle(&hs[0], posPeers_[rank_]);
    if (e == cudaSuccess && done_) e = cudaIpcGetMemHandle(&hs[1], done_);
    memcpy(out128, hs, sizeof(hs));
    return e;
  }
  cudaError_t shard_attach_ipc(const void* all) override {
    const cudaIpcMemHandle_t* hs = static_cast<const cudaIpcMemHandle_t*>(all);
    for (uint32_t r = 0; r < world_; ++r) {
      if (r == rank_) continue;
      void *pp = nullptr, *dp = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&pp, hs[2 * r], cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) return e;
      ipcOpened_.push_back(pp);
      e = cudaIpcOpenMemHandle(&dp, hs[2 * r + 1], cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) return e;
      ipcOpened_.push_back(dp);
      posPeers_[r] = static_cast<float4*>(pp);
      donePeers_[r] = static_cast<unsigned*>(dp);
    }
    attached_ = true;
    return cudaSuccess;
  }
  void shard_local_pointers(void** pos, void** done) override { *pos = posPeers_[rank_]; *done = done_; }
  cudaError_t shard_attach_pointers(uint32_t r, void* pos, void* done, int peerDevice) override {
    if (r >= world_) return cudaErrorInvalidValue;
    if (r != rank_) {
      cudaError_t e = cudaSuccess;
      if (peerDevice != device_) e = cudaDeviceEnablePeerAccess(peerDevice, 0);   // (two ranks on one device: plain pointers)
      if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
      if (e != cudaSuccess) return e;
      posPeers_[r] = static_cast<float4*>(pos);
      donePeers_[r] = static_cast<unsigned*>(done);
    }
    bool all = true;
    for (uint32_t q = 0; q < world_; ++q) all &= posPeers_[q] != nullptr;
    attached_ = all;
    return cudaSuccess;
  }
  cudaError_t export_pos(const DeviceArrays& d, cudaStream_t s) override {
    if (!tagged_ || d.V == 0) return cudaSuccess;
    from_tagged_kernel<<<(d.V + 255) / 256, 256, 0, s>>>(posT_, d.pos, d.V);
    return cudaGetLastError();
  }
  bool take_abort() override {
    if (!abortHost_ || *reinterpret_cast<volatile unsigned*>(abortHost_) == 0u) return false;
    *abortHost_ = 0u;
    return true;
  }
  uint32_t launches_per_frame(const FrameShape&) const override { return 1; }
  uint64_t device_bytes() const override { return bytes_; }
  void fill_info(pbd_info& info) const override {
    info.grid_blocks = grid_;
    info.block_threads = block_;
    info.lanes_per_tet = lanes_;
  }

 private:
  // tetInert: alpha of the tets is 0 this frame (FrameShape)
  const void* kernel(bool tetInert = false) const {
    if (tagged_ && fast_ && tetInert && dropInert_)
      return resident_ ? (const void*)tile_frame_kernel<1, true, true, false, true> : (const void*)tile_frame_kernel<1, true, true, false>;
    if (tagged_) return fast_ ? (const void*)tile_frame_kernel<1, true, true> : (const void*)tile_frame_kernel<1, true, false>;
    if (fast_) return (const void*)tile_frame_kernel<1, false, true>;
    return lanes_ == 1 ? (const void*)tile_frame_kernel<1, false, false>
           : lanes_ == 2 ? (const void*)tile_frame_kernel<2, false, false> : (const void*)tile_frame_kernel<4, false, false>;
  }
  pbd_options opts_;
  int device_;
  uint4* posT_ = nullptr;        // tagged hand-over: positions as {x, y, z, tag} words
  float* invMass_ = nullptr;     // ... and the inverse masses
  unsigned* abortHost_ = nullptr;      // mapped host word + its device alias
  unsigned* abortHostDev_ = nullptr;
  static constexpr uint32_t kStageCtas = 1024;
  uint32_t* stageHost_ = nullptr;      // mapped host memory, 4 words per CTA (stage_share)
  uint32_t* stageHostDev_ = nullptr;
  uint32_t stageGrid_ = 0;
  long long spinLimit_ = 0;
  bool tagged_ = false;
  bool fast_ = false;            // PBD_FLAG_FAST_ARITH
  bool resident_ = false;        // RES kernels: three record buffers, visits 0 and 2 of every CTA keep theirs for the frame
  uint32_t tilesPerSm_ = 1;
  unsigned char* blob_ = nullptr;
  TileCopy* copies_ = nullptr;
  PhaseDesc* phases_ = nullptr;
  uint32_t* tile0Begin_ = nullptr;
  unsigned* barrier_ = nullptr;
  unsigned long long* trace_ = nullptr;
  long long* ftrace_ = nullptr;
  size_t traceN_ = 0;
  uint32_t stagger_ = 0;
  uint32_t dropInert_ = 1;
  uint32_t world_ = 1, rank_ = 0, nHome_ = 0, iterBase_ = 0;
  bool attached_ = true;
  std::vector<uint8_t> homeOwner_;
  std::vector<uint32_t> slotOwnerBegin_;
  float4* posPeers_[kMaxRanks] = {};
  unsigned* donePeers_[kMaxRanks] = {};
  uint32_t* tileList_ = nullptr;
  uint32_t* homeList_ = nullptr;
  std::vector<void*> ipcOpened_;
  unsigned* done_ = nullptr;
  size_t doneBytes_ = 0;
  bool useFlags_ = false;
  uint32_t block_ = 512, grid_ = 1, nPhases_ = 0, nTile0_ = 0, maxTilesPerPhase_ = 0, lanes_ = 4, recStride_ = 128;
  size_t smemBytes_ = 0, smemBytesRes_ = 0;
  uint32_t recStrideRes_ = 128;
  uint64_t bytes_ = 0;
};

}  // namespace

Backend* make_tile_backend(const pbd_options& opts, int device) { return new TileBackend(opts, device); }

}  // namespace pbd

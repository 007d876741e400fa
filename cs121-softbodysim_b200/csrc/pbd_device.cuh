// pbd_device.cuh -- device helpers shared by the tile and batch kernels: PTX wrappers (L2-scope
// release/acquire, mbarrier, TMA bulk copies) and the fused vertex stages.
#pragma once
#include <cstdint>

#include "pbd_math.cuh"

namespace pbd {

enum LoadMode { LOAD_PLAIN = 0, LOAD_GROUND = 1, LOAD_PREDICT = 2, LOAD_COMMIT_PREDICT = 3 };

// ---------------------------------------------------------------- PTX helpers

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// system scope: the flag lives on another GPU of the node (peer memory over NVLink)
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// TMA bulk copy global -> shared memory of this CTA, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_load(void* dstSmem, const void* srcGlobal, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dstSmem)),
               "l"(srcGlobal), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// TMA bulk copy shared -> global, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_store(void* dstGlobal, const void* srcSmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dstGlobal), "r"(smem_u32(srcSmem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// ... until at most N of this thread's bulk async-groups are still pending
template <int N>
__device__ __forceinline__ void bulk_wait_pending() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// make this thread's generic-proxy shared-memory writes visible to the async proxy (bulk copies)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------- vertex stages

// Vertex stage applied while a phase-0 tile is loaded, split into its loads and its arithmetic +
// stores so that a thread can have the loads of several vertices in flight before the first store
// (the compiler must assume that stores to prev/vel alias later loads).
struct VertexIn {
  float4 p, x, v;   // (xStar, w) | committed position | velocity
};
template <class Arrays>
__device__ __forceinline__ VertexIn load_vertex(const Arrays& P, uint32_t s, int mode) {
  VertexIn in;
  in.p = __ldcg(P.pos + s);
  in.x = in.v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (mode == LOAD_PREDICT || mode == LOAD_COMMIT_PREDICT) in.x = __ldcg(P.prev + s);
  if (mode == LOAD_PREDICT) in.v = __ldcg(P.vel + s);
  return in;
}
template <class Arrays>
__device__ __forceinline__ float4 finish_vertex(const Arrays& P, const StepConsts& k, uint32_t s, int mode,
                                                bool clampFirst, VertexIn in) {
  float4 p = in.p;
  if (mode == LOAD_GROUND) {
    ground_vertex(p, k);
    if (P.nColliders) collide_vertex(p, P.colliders, P.nColliders);
  } else if (mode == LOAD_PREDICT) {
    float4 v = in.v;
    p = predict_vertex(in.x, v, p.w, k);
    __stcg(P.vel + s, v);
  } else if (mode == LOAD_COMMIT_PREDICT) {
    float4 x = in.x, v;
    if (clampFirst) { ground_vertex(p, k); if (P.nColliders) collide_vertex(p, P.colliders, P.nColliders); }
    commit_vertex(p, x, v, k);
    v.w = 0.0f;
    p = predict_vertex(x, v, p.w, k);
    __stcg(P.prev + s, x);
    __stcg(P.vel + s, v);
  }
  return p;
}
template <class Arrays>
__device__ __forceinline__ float4 load_transform(const Arrays& P, const StepConsts& k, uint32_t s, int mode,
                                                 bool clampFirst) {
  return finish_vertex(P, k, s, mode, clampFirst, load_vertex(P, s, mode));
}

}  // namespace pbd

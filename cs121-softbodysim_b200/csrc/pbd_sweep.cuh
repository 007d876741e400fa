// pbd_sweep.cuh -- the shared-memory colour sweeps of the tile backend (device code only).
//
// A tile's RECORD BLOCK (built by pbd_tile.cu, layout sized by pbd_plan.h::tile_record_bytes) and
// its vertices live in dynamic shared memory; these functions project the tile's edge colour
// groups and then its tet colour groups on them.  Kept in a header so that tools/mb_sweep.cu can
// time exactly the code the frame kernel runs.
//
// Reference lines restated: CProgram/src/Sim.cpp:104-129 (edge), :136-172 (tet), via pbd_math.cuh.
#pragma once
#include <cstdint>
#include <type_traits>

#include "pbd_math.cuh"

namespace pbd {

// Inlined into the frame kernels (measured A/B on one box, tools/gpu_ab.sh: +3.7 % over calling
// them out of line once the colour loops had been reduced to a single pass without prefetch).
#ifndef PBD_SWEEP_INLINE
#define PBD_SWEEP_INLINE static __device__ __forceinline__
#endif
// per-step clock stamps for PBD_TILE_TRACE (compiled in only with -DPBD_TRACE_STEPS: even a
// predicated-off stamp per colour step costs measurable time in these loops)
#ifdef PBD_TRACE_STEPS
#define PBD_STEP_TRACE(ft, g, n) do { if ((ft) && (g) < 40) { (ft)[16 + (g)] = clock64(); (ft)[56 + (g)] = (n); } } while (0)
#else
#define PBD_STEP_TRACE(ft, g, n) do { (void)(ft); } while (0)
#endif
#ifndef PBD_SWEEP_INLINE
#define PBD_SWEEP_INLINE static __device__ __noinline__
#endif

// first 64 bytes of a tile's record block (shared memory); offsets in bytes from the block start
struct TileHdr {
  uint32_t vertCount, flags, vertBegin, nEdgeGroups;   // flags: bit 0 = contiguous slot range, bit 1 = system-scope sync, bit 2 = mixed steps, bit 3 = riders (u32 per tet behind the tet rest values), bits 8..15 = predecessor count
  uint32_t nTetGroups, nEdges, nTets, offVertIdx;
  uint32_t offEdgeGroups, offTetGroups, offEdgeIdx, offEdgeRest;
  uint32_t offTetIdx, offTetRest, offEdgeLam, offTetLam;
};
static_assert(sizeof(TileHdr) == 64, "TileHdr is the 64-byte block header");

// ---------------------------------------------------------------- raw shared-memory accesses
//
// The one-thread-per-constraint sweeps address shared memory with 32-bit shared-window addresses
// computed ONCE per tile visit.  Going through generic pointers (smem + offset) made the compiler
// rebuild the window base in every colour step (S2UR SR_CgaCtaId / UMOV / ULEA / LDC: ~10 of the
// ~110 instructions an edge thread issued per step, and on its dependent path).  `volatile` +
// "memory" keep the accesses on their side of the block barriers.
PBD_DEV uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
PBD_DEV float lds_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}
PBD_DEV uint2 lds_v2(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
  return v;
}
PBD_DEV uint4 lds_v4u(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
PBD_DEV float4 lds_v4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
PBD_DEV void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
PBD_DEV void sts_v4(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// (the volatile move keeps the window address in a register: the compiler otherwise rematerialises
// it -- S2UR SR_CgaCtaId + ULEA -- inside every colour step)
PBD_DEV uint32_t smem_window(const void* p) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p), r;
  asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(a));
  return r;
}

// Shared address of the vertex a packed 16-bit tile-local index names: sv + 16 * index.  Written as AND / SHR + multiply-add
// in PTX (ptxas: LOP3 or SHF, then LEA) -- the C form (x & 0xffff) << 4 is canonicalised to shift, mask, add: three dependent
// instructions between the index load and the vertex gather instead of two: +2.0 % on the headline (fast), measured.
PBD_DEV uint32_t vert_lo(uint32_t sv, uint32_t id) {
  uint32_t a;
  asm("{\n\t.reg .u32 t;\n\tand.b32 t, %1, 0xffff;\n\tmad.lo.u32 %0, t, 16, %2;\n\t}" : "=r"(a) : "r"(id), "r"(sv));
  return a;
}
PBD_DEV uint32_t vert_hi(uint32_t sv, uint32_t id) {
  uint32_t a;
  asm("{\n\t.reg .u32 t;\n\tshr.u32 t, %1, 16;\n\tmad.lo.u32 %0, t, 16, %2;\n\t}" : "=r"(a) : "r"(id), "r"(sv));
  return a;
}

// one edge / one tet whose record sits at the given shared addresses; sv = address of the tile's vertex 0
template <bool FAST>
PBD_DEV void project_edge_rec(uint32_t sv, uint32_t id, float r, float l, uint32_t lamA, float alpha) {
  const uint32_t a = vert_lo(sv, id), b = vert_hi(sv, id);
  const float4 p0 = lds_v4(a), p1 = lds_v4(b);
  float4 q0, q1;
  float nl;
  if (FAST ? edge_delta_fast(p0, p1, r, l, alpha, q0, q1, nl) : edge_delta(p0, p1, r, l, alpha, q0, q1, nl)) {
    sts_v4(a, q0);
    sts_v4(b, q1);
    sts_f32(lamA, nl);
  }
}
template <bool FAST>
PBD_DEV void project_edge_at(uint32_t sv, uint32_t idA, uint32_t restA, uint32_t lamA, float alpha) {
  const uint32_t id = lds_u32(idA);
  const float r = lds_f32(restA), l = lds_f32(lamA);
  project_edge_rec<FAST>(sv, id, r, l, lamA, alpha);
}
// useLam == false (fast arithmetic, alpha == 0): the multiplier is inert -- it never enters a correction -- and is
// neither read nor accumulated (see tile_frame_kernel: its range is not moved between global and shared memory either)
template <bool FAST>
PBD_DEV void project_tet_rec(uint32_t sv, uint2 id, float r, float l, uint32_t lamA, float alpha, bool useLam = true) {
  const uint32_t a = vert_lo(sv, id.x), b = vert_hi(sv, id.x), c = vert_lo(sv, id.y), d = vert_hi(sv, id.y);
  float4 pa = lds_v4(a), pb = lds_v4(b), pc = lds_v4(c), pd = lds_v4(d);
  float nl;
  if (FAST ? tet_delta_fast(pa, pb, pc, pd, r, l, alpha, nl, !useLam) : tet_delta(pa, pb, pc, pd, r, l, alpha, nl)) {
    sts_v4(a, pa); sts_v4(b, pb); sts_v4(c, pc); sts_v4(d, pd);
    if (useLam) sts_f32(lamA, nl);
  }
}
template <bool FAST>
PBD_DEV void project_tet_at(uint32_t sv, uint32_t idA, uint32_t restA, uint32_t lamA, float alpha, bool useLam = true) {
  const uint2 id = lds_v2(idA);
  const float r = lds_f32(restA), l = useLam ? lds_f32(lamA) : 0.0f;
  project_tet_rec<FAST>(sv, id, r, l, lamA, alpha, useLam);
}

// PBD_ORDER_RIDING: the (at most two) edges that ride on the tet at `o4` = 4 * its position are
// projected by the tet's own thread right after the tet, on the vertices it has just written (same
// thread, same addresses: program order, no barrier).  rideA = address of the tile's ride array,
// eIdx/eRest/eLam = addresses of entry 0 of the tile's edge arrays.
template <bool FAST>
PBD_DEV void project_riders(uint32_t sv, uint32_t rideA, uint32_t o4, uint32_t eIdx, uint32_t eRest, uint32_t eLam, float alphaE) {
  const uint32_t rp = lds_u32(rideA + o4);
  const uint32_t p0 = (rp & 0xffffu) << 2, p1 = (rp >> 16) << 2;
  if (p0 != (0xffffu << 2)) project_edge_at<FAST>(sv, eIdx + p0, eRest + p0, eLam + p0, alphaE);
  if (p1 != (0xffffu << 2)) project_edge_at<FAST>(sv, eIdx + p1, eRest + p1, eLam + p1, alphaE);
}

#ifdef PBD_NO_REG_RIDERS   // A/B switch (tools/build_variant.sh): fast riders through shared memory like the exact ones
constexpr bool kRegRiders = false;
#else
constexpr bool kRegRiders = true;
#endif
// Fast arithmetic + riders: the whole unit on registers.  The upload relabelled the tet (even
// permutation) so that rider 0 is the edge (a, b) and rider 1 the edge (c, d); the riders are
// projected on the tet's freshly updated registers and the four vertices are stored once.
PBD_DEV void project_tet_unit_fast(uint32_t sv, uint32_t idA, uint32_t restA, uint32_t lamA, float alphaT, uint32_t rideWordA,
                                   uint32_t eRest, uint32_t eLam, float alphaE, bool useLam = true) {
  const uint2 id = lds_v2(idA);
  const float r = lds_f32(restA), l = useLam ? lds_f32(lamA) : 0.0f;
  const uint32_t rp = lds_u32(rideWordA);
  const uint32_t a = vert_lo(sv, id.x), b = vert_hi(sv, id.x), c = vert_lo(sv, id.y), d = vert_hi(sv, id.y);
  float4 pa = lds_v4(a), pb = lds_v4(b), pc = lds_v4(c), pd = lds_v4(d);
  const uint32_t p0 = (rp & 0xffffu) << 2, p1 = (rp >> 16) << 2;
  const bool has0 = p0 != (0xffffu << 2), has1 = p1 != (0xffffu << 2);
  float r0 = 0.f, l0 = 0.f, r1 = 0.f, l1 = 0.f;
  if (has0) { r0 = lds_f32(eRest + p0); l0 = lds_f32(eLam + p0); }
  if (has1) { r1 = lds_f32(eRest + p1); l1 = lds_f32(eLam + p1); }
  {
    float4 qa = pa, qb = pb, qc = pc, qd = pd;
    float nl;
    if (tet_delta_fast(qa, qb, qc, qd, r, l, alphaT, nl, !useLam)) { pa = qa; pb = qb; pc = qc; pd = qd; if (useLam) sts_f32(lamA, nl); }
  }
  if (has0) {
    float4 q0, q1;
    float nl;
    if (edge_delta_fast(pa, pb, r0, l0, alphaE, q0, q1, nl)) { pa = q0; pb = q1; sts_f32(eLam + p0, nl); }
  }
  if (has1) {
    float4 q0, q1;
    float nl;
    if (edge_delta_fast(pc, pd, r1, l1, alphaE, q0, q1, nl)) { pc = q0; pd = q1; sts_f32(eLam + p1, nl); }
  }
  sts_v4(a, pa); sts_v4(b, pb); sts_v4(c, pc); sts_v4(d, pd);
}

// ---------------------------------------------------------------- sweeps (shared memory only)
//
// Everything a colour step touches lives in shared memory: `rec` / `svOff` are byte offsets into
// the dynamic shared array, so every access below is an LDS/STS.  The dependent chain of a
// colour step is LDS record -> LDS.128 vertices -> arithmetic -> STS.128 -> block barrier.

// FAST: the PBD_FLAG_FAST_ARITH forms of pbd_math.cuh (FFMA + SFU reciprocal / rsqrt) instead of the
// bit-exact ones; same records, same schedule.
template <bool FAST>
PBD_SWEEP_INLINE void sweep_edges(const TileHdr& h, uint32_t rec, uint32_t svOff, float alpha, long long* ft) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t n = h.nEdgeGroups;
  if (n == 0) return;
  const uint32_t base = smem_window(smem), sv = base + svOff, tid = threadIdx.x;
  uint32_t grp = base + rec + h.offEdgeGroups;
  const uint32_t idA = base + rec + h.offEdgeIdx + 4u * tid, restA = base + rec + h.offEdgeRest + 4u * tid,
                 lamA = base + rec + h.offEdgeLam + 4u * tid;
  // (Fetching a thread's next record before the barrier was measured: no gain, more registers.)
  for (uint32_t g = 0; g < n; ++g, grp += 8u) {
    const uint2 gd = lds_v2(grp);
    if (tid < gd.y) {   // the planner keeps every group within one pass of the block (checked at upload)
      const uint32_t o = gd.x << 2;
      project_edge_at<FAST>(sv, idA + o, restA + o, lamA + o, alpha);
    }
    __syncthreads();
    PBD_STEP_TRACE(ft, g, gd.y);
  }
}

// LANES == 1: one thread per tet.
// LANES == 4: one tet per 4 adjacent lanes, lane `role` owns vertex `role` (a,b,c,d).  All four
// gradients have the form cross(x - o, y - o)/6 (Sim.cpp:146-149):
//   ga: o=b x=d y=c | gb: o=a x=c y=d | gc: o=a x=d y=b | gd: o=a x=b y=c
// so every lane runs the same instructions on role-selected operands; the reduction terms are
// exchanged with quad shuffles and summed in the reference's order, which keeps the result
// bit-identical while shortening the dependent instruction stream of a colour step.
template <int LANES, bool FAST>
PBD_SWEEP_INLINE void sweep_tets(const TileHdr& h, uint32_t rec, uint32_t svOff, float alpha, long long* ft, float alphaE = 0.0f,
                                 bool useLam = true) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t n = h.nTetGroups;
  if (n == 0) return;
  const uint2* groups = reinterpret_cast<const uint2*>(smem + rec + h.offTetGroups);
  const uint2* idx = reinterpret_cast<const uint2*>(smem + rec + h.offTetIdx);
  const float* rest = reinterpret_cast<const float*>(smem + rec + h.offTetRest);
  float* lam = reinterpret_cast<float*>(smem + rec + h.offTetLam);
  float4* sv = reinterpret_cast<float4*>(smem + svOff);
  const uint32_t tid = threadIdx.x;
  if (LANES == 1) {
    const uint32_t base = smem_window(smem), svA = base + svOff;
    uint32_t grp = base + rec + h.offTetGroups;
    const uint32_t idA = base + rec + h.offTetIdx + 8u * tid, restA = base + rec + h.offTetRest + 4u * tid,
                   lamA = base + rec + h.offTetLam + 4u * tid;
    const bool ride = (h.flags & 8u) != 0u;
    const uint32_t rideA = restA + 4u * ((h.nTets + 3u) & ~3u);   // the ride array follows the rest values
    const uint32_t eIdx = base + rec + h.offEdgeIdx, eRest = base + rec + h.offEdgeRest, eLam = base + rec + h.offEdgeLam;
    for (uint32_t g = 0; g < n; ++g, grp += 8u) {
      const uint2 gd = lds_v2(grp);
      if (tid < gd.y) {
        const uint32_t o = gd.x << 2;
        if (FAST && ride && kRegRiders) {
          project_tet_unit_fast(svA, idA + 2u * o, restA + o, lamA + o, alpha, rideA + o, eRest, eLam, alphaE, useLam);
        } else {
          project_tet_at<FAST>(svA, idA + 2u * o, restA + o, lamA + o, alpha, useLam);
          if (ride) project_riders<FAST>(svA, rideA, o, eIdx, eRest, eLam, alphaE);
        }
      }
      __syncthreads();
      PBD_STEP_TRACE(ft, g, gd.y);
    }
  } else if (LANES == 2) {
    // Two adjacent lanes per tet.  Lane A (even) owns vertices a, b and computes ga, gb; lane B
    // (odd) owns c, d and computes gc, gd.  The lanes load the four vertices in the orders
    //   A: (V0,V1,V2,V3) = (a,b,c,d)      B: (V0,V1,V2,V3) = (c,a,b,d)
    // so that   g1 = cross(V3-V1, V2-V1)/6   is ga on A and gc on B      (Sim.cpp:146,148)
    // and       g2 = cross(V2-o2, y2-o2)/6   with (o2,y2) = (V0,V3) on A -> gb, (V1,V0) on B -> gd.
    // B's g2 cross is cross(pb-pa, pc-pa), the volume normal; its V3-V1 is pd-pa.  The four
    // w|g|^2 terms are exchanged with one shuffle pair and summed in the reference's order on
    // both lanes, so both derive the identical delta-lambda.
    const uint32_t half = tid & 1u, lane = tid & 31u;
    const uint32_t pair = tid >> 1;
#ifdef PBD_SWEEP_OVERFLOW
    const uint32_t pairs = blockDim.x >> 1;
#endif
    const unsigned m = 0xffffffffu;
    const bool isB = half != 0u;
    auto project = [&](uint32_t t, uint2 id, float r, float l0, bool live) {
      const uint32_t ia = id.x & 0xffffu, ib = id.x >> 16, ic = id.y & 0xffffu, idd = id.y >> 16;
      const uint32_t i0 = isB ? ic : ia, i1 = isB ? ia : ib, i2 = isB ? ib : ic, i3 = idd;
      const float4 V0 = sv[i0], V1 = sv[i1], V2 = sv[i2], V3 = sv[i3];
      const float wa = isB ? V1.w : V0.w, wb = isB ? V2.w : V1.w, wc = isB ? V0.w : V2.w, wd = V3.w;
      const bool massive = fadd(fadd(fadd(wa, wb), wc), wd) != 0.0f;
      const float k6 = 1.0f / 6.0f;
      // g1
      const float ux = fsub(V3.x, V1.x), uy = fsub(V3.y, V1.y), uz = fsub(V3.z, V1.z);
      const float vx = fsub(V2.x, V1.x), vy = fsub(V2.y, V1.y), vz = fsub(V2.z, V1.z);
      const float g1x = fmul(cross_c(uy, vz, uz, vy), k6), g1y = fmul(cross_c(uz, vx, ux, vz), k6),
                  g1z = fmul(cross_c(ux, vy, uy, vx), k6);
      // g2
      const float ox = isB ? V1.x : V0.x, oy = isB ? V1.y : V0.y, oz = isB ? V1.z : V0.z;
      const float yx = isB ? V0.x : V3.x, yy = isB ? V0.y : V3.y, yz = isB ? V0.z : V3.z;
      const float px = fsub(V2.x, ox), py = fsub(V2.y, oy), pz = fsub(V2.z, oz);
      const float qx = fsub(yx, ox), qy = fsub(yy, oy), qz = fsub(yz, oz);
      const float nx = cross_c(py, qz, pz, qy), ny = cross_c(pz, qx, px, qz), nz = cross_c(px, qy, py, qx);
      const float g2x = fmul(nx, k6), g2y = fmul(ny, k6), g2z = fmul(nz, k6);
      // own vertices: u1 = V0, u2 = V1 (A) / V3 (B)
      const uint32_t iu2 = isB ? i3 : i1;
      float4 u1 = V0, u2;
      u2.x = isB ? V3.x : V1.x; u2.y = isB ? V3.y : V1.y; u2.z = isB ? V3.z : V1.z; u2.w = isB ? V3.w : V1.w;
      const float t1 = fmul(u1.w, dot3(g1x, g1y, g1z, g1x, g1y, g1z));
      const float t2 = fmul(u2.w, dot3(g2x, g2y, g2z, g2x, g2y, g2z));
      const float o1 = __shfl_xor_sync(m, t1, 1), o2 = __shfl_xor_sync(m, t2, 1);
      // ((ta + tb) + tc) + td
      const float sab = isB ? fadd(o1, o2) : fadd(t1, t2);
      const float wSum = fadd(fadd(sab, isB ? t1 : o1), isB ? t2 : o2);
      // volume numerator dot(cross(pb-pa, pc-pa), pd-pa): lane B's (n, V3-V1)
      const float vnLocal = dot3(nx, ny, nz, ux, uy, uz);
      const float vol = fdiv(__shfl_sync(m, vnLocal, lane | 1u), 6.0f);
      const float C = fsub(vol, r);
      const float dl = fdiv_pos(fsub(-C, fmul(alpha, l0)), fadd(wSum, alpha));
      const float s1 = fmul(u1.w, dl), s2 = fmul(u2.w, dl);
      u1.x = fadd(u1.x, fmul(g1x, s1)); u1.y = fadd(u1.y, fmul(g1y, s1)); u1.z = fadd(u1.z, fmul(g1z, s1));
      u2.x = fadd(u2.x, fmul(g2x, s2)); u2.y = fadd(u2.y, fmul(g2y, s2)); u2.z = fadd(u2.z, fmul(g2z, s2));
      if (live && massive && !(wSum < 1e-20f)) {
        sv[i0] = u1;
        sv[iu2] = u2;
        if (!isB) lam[t] = fadd(l0, dl);
      }
    };
    uint2 gd = groups[0];
    bool live = pair < gd.y;
    uint32_t t = gd.x + (live ? pair : gd.y - 1u);
    uint2 id = idx[t];
    float r = rest[t], l = lam[t];
    for (uint32_t g = 0; g < n; ++g) {
      const uint2 gn = (g + 1 < n) ? groups[g + 1] : make_uint2(0u, 1u);
      if (((tid >> 5) << 4) < gd.y) project(t, id, r, l, live);                // warp-uniform: idle warps skip the step
#ifdef PBD_SWEEP_OVERFLOW
      for (uint32_t i0 = pairs + ((tid >> 5) << 4); i0 < gd.y; i0 += pairs) {   // warp-uniform trip count
        const uint32_t i = i0 + (lane >> 1);
        const bool lv = i < gd.y;
        const uint32_t t2 = gd.x + (lv ? i : gd.y - 1u);
        project(t2, idx[t2], rest[t2], lam[t2], lv);
      }
#endif
      live = pair < gn.y && g + 1 < n;
      t = gn.x + ((pair < gn.y) ? pair : gn.y - 1u);
      if (g + 1 < n) { id = idx[t]; r = rest[t]; l = lam[t]; }
      __syncthreads();
      PBD_STEP_TRACE(ft, g, gd.y);
      gd = gn;
    }
  } else {
    const uint32_t role = tid & 3u, lane = tid & 31u, qbase = lane & ~3u;
    const uint32_t fo = (0x00000001u >> (role * 8u)) & 3u;        // {1,0,0,0}
    const uint32_t fx = (0x01030203u >> (role * 8u)) & 3u;        // {3,2,3,1}
    const uint32_t fy = (0x02010302u >> (role * 8u)) & 3u;        // {2,3,1,2}
    const uint32_t quad = tid >> 2;
#ifdef PBD_SWEEP_OVERFLOW
    const uint32_t quads = blockDim.x >> 2;
#endif
    const unsigned m = 0xffffffffu;
    // `live` is quad-uniform; idle quads run the same instructions on the group's last tet and
    // write nothing, so every lane of a warp takes part in the shuffles
    auto project = [&](uint32_t t, uint2 id, float r, float l0, bool live) {
      auto pick = [&](uint32_t f) -> uint32_t { return (((f & 2u) ? id.y : id.x) >> ((f & 1u) * 16u)) & 0xffffu; };
      const uint32_t iown = pick(role);
      float4 own = sv[iown];
      const float4 o = sv[pick(fo)], x = sv[pick(fx)], y = sv[pick(fy)];
      const float wa = __shfl_sync(m, own.w, qbase), wb = __shfl_sync(m, own.w, qbase + 1),
                  wc = __shfl_sync(m, own.w, qbase + 2), wd = __shfl_sync(m, own.w, qbase + 3);
      const bool massive = fadd(fadd(fadd(wa, wb), wc), wd) != 0.0f;   // quad-uniform
      const float k6 = 1.0f / 6.0f;
      const float ux = fsub(x.x, o.x), uy = fsub(x.y, o.y), uz = fsub(x.z, o.z);
      const float vx = fsub(y.x, o.x), vy = fsub(y.y, o.y), vz = fsub(y.z, o.z);
      const float nx = cross_c(uy, vz, uz, vy), ny = cross_c(uz, vx, ux, vz), nz = cross_c(ux, vy, uy, vx);
      const float gx = fmul(nx, k6), gy = fmul(ny, k6), gz = fmul(nz, k6);
      const float tt = fmul(own.w, dot3(gx, gy, gz, gx, gy, gz));
      const float ta = __shfl_sync(m, tt, qbase), tb = __shfl_sync(m, tt, qbase + 1), tc = __shfl_sync(m, tt, qbase + 2),
                  td = __shfl_sync(m, tt, qbase + 3);
      const float wSum = fadd(fadd(fadd(ta, tb), tc), td);
      // role 3 holds n = cross(pb-pa, pc-pa) and own - o = pd - pa: the volume numerator
      const float vn = dot3(nx, ny, nz, fsub(own.x, o.x), fsub(own.y, o.y), fsub(own.z, o.z));
      const float vol = fdiv(__shfl_sync(m, vn, qbase + 3), 6.0f);
      const float C = fsub(vol, r);
      const float dl = fdiv_pos(fsub(-C, fmul(alpha, l0)), fadd(wSum, alpha));
      const float sc = fmul(own.w, dl);
      own.x = fadd(own.x, fmul(gx, sc)); own.y = fadd(own.y, fmul(gy, sc)); own.z = fadd(own.z, fmul(gz, sc));
      if (live && massive && !(wSum < 1e-20f)) {
        sv[iown] = own;
        if (role == 0) lam[t] = fadd(l0, dl);
      }
    };
    uint2 gd = groups[0];
    bool live = quad < gd.y;
    uint32_t t = gd.x + (live ? quad : gd.y - 1u);
    uint2 id = idx[t];
    float r = rest[t], l = lam[t];
    for (uint32_t g = 0; g < n; ++g) {
      const uint2 gn = (g + 1 < n) ? groups[g + 1] : make_uint2(0u, 1u);
      if (((tid >> 5) << 3) < gd.y) project(t, id, r, l, live);   // warp-uniform: idle warps skip the step
#ifdef PBD_SWEEP_OVERFLOW
      // colour groups larger than a block's worth of quads: warp-uniform trip count
      for (uint32_t i0 = quads + ((tid >> 5) << 3); i0 < gd.y; i0 += quads) {
        const uint32_t i = i0 + (lane >> 2);
        const bool lv = i < gd.y;
        const uint32_t t2 = gd.x + (lv ? i : gd.y - 1u);
        project(t2, idx[t2], rest[t2], lam[t2], lv);
      }
#endif
      live = quad < gn.y && g + 1 < n;
      t = gn.x + ((quad < gn.y) ? quad : gn.y - 1u);
      if (g + 1 < n) { id = idx[t]; r = rest[t]; l = lam[t]; }
      __syncthreads();
      PBD_STEP_TRACE(ft, g, gd.y);
      gd = gn;
    }
  }
}

// Mixed colour steps (interleaved order, one thread per constraint): the planner coloured the
// tile's edges and tets TOGETHER, so step s projects edge group s on the block's first threads and
// tet group s on its last threads -- all vertex-disjoint -- before one block barrier.  The two
// groups sit in different warps and their sum fits the block (pbd_tileplan.cpp::colour_joint).
// A visit then needs about max-joint-vertex-load steps instead of edge colours + tet colours, and
// the warps a tet-only step would leave idle carry the edge work.
template <bool FAST>
PBD_SWEEP_INLINE void sweep_mixed(const TileHdr& h, uint32_t rec, uint32_t svOff, float alphaE, float alphaT,
                                  long long* ft, bool useTetLam = true) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t n = h.nEdgeGroups;   // == h.nTetGroups
  if (n == 0) return;
  const uint32_t base = smem_window(smem), sv = base + svOff;
  const uint32_t tid = threadIdx.x, rtid = blockDim.x - 1u - tid;   // tets are counted from the block's last thread
  uint32_t eGrp = base + rec + h.offEdgeGroups, tGrp = base + rec + h.offTetGroups;
  const uint32_t eIdA = base + rec + h.offEdgeIdx + 4u * tid, eRestA = base + rec + h.offEdgeRest + 4u * tid,
                 eLamA = base + rec + h.offEdgeLam + 4u * tid;
  const uint32_t tIdA = base + rec + h.offTetIdx + 8u * rtid, tRestA = base + rec + h.offTetRest + 4u * rtid,
                 tLamA = base + rec + h.offTetLam + 4u * rtid;
  const bool ride = (h.flags & 8u) != 0u;
  const uint32_t rideA = tRestA + 4u * ((h.nTets + 3u) & ~3u);   // the ride array follows the tet rest values
  const uint32_t eIdx0 = base + rec + h.offEdgeIdx, eRest0 = base + rec + h.offEdgeRest, eLam0 = base + rec + h.offEdgeLam;
#ifdef PBD_SWEEP_PREFETCH
  // The record of a thread's NEXT constraint (group entry, indices, rest value, lambda -- none of
  // which another constraint ever writes) is fetched before the block barrier, so that after the
  // barrier the dependent chain starts at the vertex gathers.
  uint32_t role, lamA = 0u;
  uint2 id = make_uint2(0u, 0u);
  float r = 0.f, l = 0.f;
  auto fetch = [&]() {
    const uint2 ge = lds_v2(eGrp), gt = lds_v2(tGrp);
    role = 0u;
    if (tid < ge.y) {
      const uint32_t o = ge.x << 2;
      role = 1u; id.x = lds_u32(eIdA + o); r = lds_f32(eRestA + o); lamA = eLamA + o; l = lds_f32(lamA);
    } else if (rtid < gt.y) {
      const uint32_t o = gt.x << 2;
      role = 2u; id = lds_v2(tIdA + 2u * o); r = lds_f32(tRestA + o); lamA = tLamA + o; l = lds_f32(lamA);
    }
  };
  fetch();
  for (uint32_t g = 0; g < n; ++g) {
    const uint32_t roleC = role, lamC = lamA;
    const uint2 idC = id;
    const float rC = r, lC = l;
    eGrp += 8u; tGrp += 8u;
    if (roleC == 1u) {
      const uint32_t a = sv + ((idC.x & 0xffffu) << 4), b = sv + ((idC.x >> 16) << 4);
      const float4 p0 = lds_v4(a), p1 = lds_v4(b);
      if (g + 1 < n) fetch();
      float4 q0, q1;
      float nl;
      if (FAST ? edge_delta_fast(p0, p1, rC, lC, alphaE, q0, q1, nl) : edge_delta(p0, p1, rC, lC, alphaE, q0, q1, nl)) {
        sts_v4(a, q0); sts_v4(b, q1); sts_f32(lamC, nl);
      }
    } else if (roleC == 2u) {
      const uint32_t a = sv + ((idC.x & 0xffffu) << 4), b = sv + ((idC.x >> 16) << 4);
      const uint32_t c = sv + ((idC.y & 0xffffu) << 4), d = sv + ((idC.y >> 16) << 4);
      float4 pa = lds_v4(a), pb = lds_v4(b), pc = lds_v4(c), pd = lds_v4(d);
      if (g + 1 < n) fetch();
      float nl;
      if (FAST ? tet_delta_fast(pa, pb, pc, pd, rC, lC, alphaT, nl) : tet_delta(pa, pb, pc, pd, rC, lC, alphaT, nl)) {
        sts_v4(a, pa); sts_v4(b, pb); sts_v4(c, pc); sts_v4(d, pd); sts_f32(lamC, nl);
      }
    } else if (g + 1 < n) {
      fetch();
    }
    __syncthreads();
    PBD_STEP_TRACE(ft, g, 0);
  }
#else
  // a mixed tile's two group tables are stored as ONE table of {edge begin, edge count, tet begin, first tet thread}
  // entries (pbd_tile.cu upload; the two sections are adjacent and together exactly that large): one LDS.128
  // per step and warp instead of two LDS.64
  (void)tGrp; (void)rtid;
  // The step's barrier waits for the tet warps (~100 dependent instructions against the edges' ~40), so whatever sits
  // between the entry load and a tet thread's first shared load is on the step's critical path: the tet test compares
  // with the entry's fourth word = the block's first tet thread (block size - tet count, written by the upload --
  // comparing a count with blockDim.x - 1 - tid made the compiler rebuild that number in every step), and the
  // per-tile `ride` switch is taken once per visit, not once per step.
  // (Measured and not kept: the tet test ahead of the edge test, -1.4 %; fetching the NEXT step's entry before this
  // step's barrier, -2.2 % fast, -3.9 % exact.)
  auto steps = [&](auto withRiders) {
    constexpr bool RIDE = decltype(withRiders)::value;
    for (uint32_t g = 0; g < n; ++g, eGrp += 16u) {
      const uint4 gq = lds_v4u(eGrp);
      if (tid < gq.y) {
        const uint32_t o = gq.x << 2;
        project_edge_at<FAST>(sv, eIdA + o, eRestA + o, eLamA + o, alphaE);
      } else if (tid >= gq.w) {
        const uint32_t o = gq.z << 2;
        if (FAST && RIDE && kRegRiders) {
          project_tet_unit_fast(sv, tIdA + 2u * o, tRestA + o, tLamA + o, alphaT, rideA + o, eRest0, eLam0, alphaE, useTetLam);
        } else {
          project_tet_at<FAST>(sv, tIdA + 2u * o, tRestA + o, tLamA + o, alphaT, useTetLam);
          if (RIDE) project_riders<FAST>(sv, rideA, o, eIdx0, eRest0, eLam0, alphaE);
        }
      }
      __syncthreads();
      PBD_STEP_TRACE(ft, g, gq.y + ((blockDim.x - gq.w) << 16));
    }
  };
  if (ride) steps(std::true_type{}); else steps(std::false_type{});
#endif
}

}  // namespace pbd

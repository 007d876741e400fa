// pbd_math.cuh -- the per-constraint XPBD arithmetic, shared by every backend.
//
// IEEE binary32, round-to-nearest, in the reference's evaluation order and WITHOUT fused
// multiply-add: the reference is built for baseline x86-64 (SSE2 scalar), so a*b+c rounds twice.
// Every operation below is an explicit __f{add,sub,mul,div,sqrt}_rn intrinsic, which nvcc never
// contracts into FFMA and which keeps denormals (no -ftz), so the result is bit-identical to
// the reference given identical inputs and order.  (The path is latency/bandwidth bound; the
// extra FADD/FMUL issue slots are not the limiter.)
//
// Reference lines restated:
//   project_edge  CProgram/src/Sim.cpp:104-129   (body of solve_edges_xpbd_gs)
//   project_tet   CProgram/src/Sim.cpp:136-172   (body of solve_tets_xpbd_gs)
//   tet volume    CProgram/include/PBDServer.h:140-145
#pragma once
#include <cuda_runtime.h>

namespace pbd {

#define PBD_DEV __device__ __forceinline__

PBD_DEV float fmul(float a, float b) { return __fmul_rn(a, b); }
PBD_DEV float fadd(float a, float b) { return __fadd_rn(a, b); }
PBD_DEV float fsub(float a, float b) { return __fsub_rn(a, b); }
PBD_DEV float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// a.y*b.z - a.z*b.y etc. (PBDServer.h:134-136), each product rounded, then the difference
PBD_DEV float cross_c(float ay, float bz, float az, float by) { return fsub(fmul(ay, bz), fmul(az, by)); }
// x*x' + y*y' + z*z' left to right (PBDServer.h:133)
PBD_DEV float dot3(float ax, float ay, float az, float bx, float by, float bz) {
  return fadd(fadd(fmul(ax, bx), fmul(ay, by)), fmul(az, bz));
}

// a / b for a finite b > 0, IEEE round-to-nearest like fdiv, but a zero numerator (a constraint
// that is exactly satisfied: the whole body while it is in free fall) does not send the warp
// through the division's slow path: +-0 / b = +-0 = a.
PBD_DEV float fdiv_pos(float a, float b) {
  const bool z = a == 0.0f;
  const float q = fdiv(z ? 1.0f : a, b);
  return z ? a : q;
}

// Edge-distance projection.  p0/p1 carry (x, y, z, invMass).  Returns false when the reference
// `continue`s (no write).  lambda is updated in place.
PBD_DEV bool project_edge(float4& p0, float4& p1, float rest, float& lambda, float alpha) {
  const float w0 = p0.w, w1 = p1.w;
  const float wSum = fadd(w0, w1);
  if (wSum == 0.0f) return false;
  const float dx = fsub(p0.x, p1.x), dy = fsub(p0.y, p1.y), dz = fsub(p0.z, p1.z);
  const float len = __fsqrt_rn(dot3(dx, dy, dz, dx, dy, dz));
  if (len < 1e-12f) return false;
  const float C = fsub(len, rest);
  const float dl = fdiv(fsub(-C, fmul(alpha, lambda)), fadd(wSum, alpha));
  lambda = fadd(lambda, dl);
  const float inv = fdiv(1.0f, len);
  const float cx = fmul(fmul(dx, inv), dl), cy = fmul(fmul(dy, inv), dl), cz = fmul(fmul(dz, inv), dl);
  p0.x = fadd(p0.x, fmul(cx, w0)); p0.y = fadd(p0.y, fmul(cy, w0)); p0.z = fadd(p0.z, fmul(cz, w0));
  p1.x = fsub(p1.x, fmul(cx, w1)); p1.y = fsub(p1.y, fmul(cy, w1)); p1.z = fsub(p1.z, fmul(cz, w1));
  return true;
}

// Tet-volume projection.
PBD_DEV bool project_tet(float4& pa, float4& pb, float4& pc, float4& pd, float rest, float& lambda,
                         float alpha) {
  const float k6 = 1.0f / 6.0f;  // the gradients MULTIPLY by 1/6 (Sim.cpp:146-149) ...
  const float wa = pa.w, wb = pb.w, wc = pc.w, wd = pd.w;
  if (fadd(fadd(fadd(wa, wb), wc), wd) == 0.0f) return false;

  const float dbx = fsub(pd.x, pb.x), dby = fsub(pd.y, pb.y), dbz = fsub(pd.z, pb.z);  // pd - pb
  const float cbx = fsub(pc.x, pb.x), cby = fsub(pc.y, pb.y), cbz = fsub(pc.z, pb.z);  // pc - pb
  const float cax = fsub(pc.x, pa.x), cay = fsub(pc.y, pa.y), caz = fsub(pc.z, pa.z);  // pc - pa
  const float dax = fsub(pd.x, pa.x), day = fsub(pd.y, pa.y), daz = fsub(pd.z, pa.z);  // pd - pa
  const float bax = fsub(pb.x, pa.x), bay = fsub(pb.y, pa.y), baz = fsub(pb.z, pa.z);  // pb - pa

  const float gax = fmul(cross_c(dby, cbz, dbz, cby), k6), gay = fmul(cross_c(dbz, cbx, dbx, cbz), k6),
              gaz = fmul(cross_c(dbx, cby, dby, cbx), k6);
  const float gbx = fmul(cross_c(cay, daz, caz, day), k6), gby = fmul(cross_c(caz, dax, cax, daz), k6),
              gbz = fmul(cross_c(cax, day, cay, dax), k6);
  const float gcx = fmul(cross_c(day, baz, daz, bay), k6), gcy = fmul(cross_c(daz, bax, dax, baz), k6),
              gcz = fmul(cross_c(dax, bay, day, bax), k6);
  // cross(pb-pa, pc-pa): also the normal used by the volume below
  const float nx = cross_c(bay, caz, baz, cay), ny = cross_c(baz, cax, bax, caz), nz = cross_c(bax, cay, bay, cax);
  const float gdx = fmul(nx, k6), gdy = fmul(ny, k6), gdz = fmul(nz, k6);

  const float wSum = fadd(fadd(fadd(fmul(wa, dot3(gax, gay, gaz, gax, gay, gaz)),
                                    fmul(wb, dot3(gbx, gby, gbz, gbx, gby, gbz))),
                               fmul(wc, dot3(gcx, gcy, gcz, gcx, gcy, gcz))),
                          fmul(wd, dot3(gdx, gdy, gdz, gdx, gdy, gdz)));
  if (wSum < 1e-20f) return false;

  // ... while the volume DIVIDES by 6.0f (PBDServer.h:144)
  const float vol = fdiv(dot3(nx, ny, nz, dax, day, daz), 6.0f);
  const float C = fsub(vol, rest);
  const float dl = fdiv(fsub(-C, fmul(alpha, lambda)), fadd(wSum, alpha));
  lambda = fadd(lambda, dl);

  const float sa = fmul(wa, dl), sb = fmul(wb, dl), sc = fmul(wc, dl), sd = fmul(wd, dl);
  pa.x = fadd(pa.x, fmul(gax, sa)); pa.y = fadd(pa.y, fmul(gay, sa)); pa.z = fadd(pa.z, fmul(gaz, sa));
  pb.x = fadd(pb.x, fmul(gbx, sb)); pb.y = fadd(pb.y, fmul(gby, sb)); pb.z = fadd(pb.z, fmul(gbz, sb));
  pc.x = fadd(pc.x, fmul(gcx, sc)); pc.y = fadd(pc.y, fmul(gcy, sc)); pc.z = fadd(pc.z, fmul(gcz, sc));
  pd.x = fadd(pd.x, fmul(gdx, sd)); pd.y = fadd(pd.y, fmul(gdy, sd)); pd.z = fadd(pd.z, fmul(gdz, sd));
  return true;
}

// ---- branch-free forms for the shared-memory sweeps -------------------------------------------
// Same operations in the same order as project_edge / project_tet, but the reference's early
// `continue`s become one predicate that the caller uses to suppress the write-back: the dependent
// chain of a colour step then has no divergence/reconvergence points and every operand load can
// be issued up front.  When the predicate is false the computed values are garbage (possibly
// NaN/inf) and are discarded, exactly as if the reference had skipped the constraint.
PBD_DEV bool edge_delta(const float4 p0, const float4 p1, float rest, float lambda, float alpha, float4& q0,
                        float4& q1, float& newLambda) {
  const float w0 = p0.w, w1 = p1.w;
  const float wSum = fadd(w0, w1);
  const float dx = fsub(p0.x, p1.x), dy = fsub(p0.y, p1.y), dz = fsub(p0.z, p1.z);
  const float len = __fsqrt_rn(dot3(dx, dy, dz, dx, dy, dz));
  const bool ok = (wSum != 0.0f) && !(len < 1e-12f);
  const float C = fsub(len, rest);
  const float dl = fdiv_pos(fsub(-C, fmul(alpha, lambda)), fadd(wSum, alpha));   // wSum + alpha > 0 whenever ok
  newLambda = fadd(lambda, dl);
  const float inv = __frcp_rn(len);   // == 1.0f / len, both correctly rounded
  const float cx = fmul(fmul(dx, inv), dl), cy = fmul(fmul(dy, inv), dl), cz = fmul(fmul(dz, inv), dl);
  q0.x = fadd(p0.x, fmul(cx, w0)); q0.y = fadd(p0.y, fmul(cy, w0)); q0.z = fadd(p0.z, fmul(cz, w0)); q0.w = w0;
  q1.x = fsub(p1.x, fmul(cx, w1)); q1.y = fsub(p1.y, fmul(cy, w1)); q1.z = fsub(p1.z, fmul(cz, w1)); q1.w = w1;
  return ok;
}

PBD_DEV bool tet_delta(float4& pa, float4& pb, float4& pc, float4& pd, float rest, float lambda, float alpha,
                       float& newLambda) {
  const float k6 = 1.0f / 6.0f;
  const float wa = pa.w, wb = pb.w, wc = pc.w, wd = pd.w;
  const bool massive = fadd(fadd(fadd(wa, wb), wc), wd) != 0.0f;
  const float dbx = fsub(pd.x, pb.x), dby = fsub(pd.y, pb.y), dbz = fsub(pd.z, pb.z);
  const float cbx = fsub(pc.x, pb.x), cby = fsub(pc.y, pb.y), cbz = fsub(pc.z, pb.z);
  const float cax = fsub(pc.x, pa.x), cay = fsub(pc.y, pa.y), caz = fsub(pc.z, pa.z);
  const float dax = fsub(pd.x, pa.x), day = fsub(pd.y, pa.y), daz = fsub(pd.z, pa.z);
  const float bax = fsub(pb.x, pa.x), bay = fsub(pb.y, pa.y), baz = fsub(pb.z, pa.z);
  const float gax = fmul(cross_c(dby, cbz, dbz, cby), k6), gay = fmul(cross_c(dbz, cbx, dbx, cbz), k6),
              gaz = fmul(cross_c(dbx, cby, dby, cbx), k6);
  const float gbx = fmul(cross_c(cay, daz, caz, day), k6), gby = fmul(cross_c(caz, dax, cax, daz), k6),
              gbz = fmul(cross_c(cax, day, cay, dax), k6);
  const float gcx = fmul(cross_c(day, baz, daz, bay), k6), gcy = fmul(cross_c(daz, bax, dax, baz), k6),
              gcz = fmul(cross_c(dax, bay, day, bax), k6);
  const float nx = cross_c(bay, caz, baz, cay), ny = cross_c(baz, cax, bax, caz), nz = cross_c(bax, cay, bay, cax);
  const float gdx = fmul(nx, k6), gdy = fmul(ny, k6), gdz = fmul(nz, k6);
  const float wSum = fadd(fadd(fadd(fmul(wa, dot3(gax, gay, gaz, gax, gay, gaz)),
                                    fmul(wb, dot3(gbx, gby, gbz, gbx, gby, gbz))),
                               fmul(wc, dot3(gcx, gcy, gcz, gcx, gcy, gcz))),
                          fmul(wd, dot3(gdx, gdy, gdz, gdx, gdy, gdz)));
  const bool ok = massive && !(wSum < 1e-20f);
  const float vol = fdiv(dot3(nx, ny, nz, dax, day, daz), 6.0f);
  const float C = fsub(vol, rest);
  const float dl = fdiv_pos(fsub(-C, fmul(alpha, lambda)), fadd(wSum, alpha));   // wSum + alpha > 0 whenever ok
  newLambda = fadd(lambda, dl);
  const float sa = fmul(wa, dl), sb = fmul(wb, dl), sc = fmul(wc, dl), sd = fmul(wd, dl);
  pa.x = fadd(pa.x, fmul(gax, sa)); pa.y = fadd(pa.y, fmul(gay, sa)); pa.z = fadd(pa.z, fmul(gaz, sa));
  pb.x = fadd(pb.x, fmul(gbx, sb)); pb.y = fadd(pb.y, fmul(gby, sb)); pb.z = fadd(pb.z, fmul(gbz, sb));
  pc.x = fadd(pc.x, fmul(gcx, sc)); pc.y = fadd(pc.y, fmul(gcy, sc)); pc.z = fadd(pc.z, fmul(gcz, sc));
  pd.x = fadd(pd.x, fmul(gdx, sd)); pd.y = fadd(pd.y, fmul(gdy, sd)); pd.z = fadd(pd.z, fmul(gdz, sd));
  return ok;
}

// ---- PBD_FLAG_FAST_ARITH forms -----------------------------------------------------------------
// The same projections (same constraint, same formula, same skip conditions up to rounding) written
// for the GPU's instruction set instead of the reference's SSE2 rounding sequence: products feed
// FFMA, the two scale factors 1/6 are folded into scalars (the gradients are used unscaled:
// n_k = 6 g_k), ga follows from ga + gb + gc + gd = 0, and the divisions / the square root use the
// SFU approximations (rcp / rsqrt, ~1-2 ulp).  About 100 instructions per tet and 40 per edge instead
// of ~220 / ~110.  Results differ from the exact forms at rounding level only; this mode is validated
// by the tolerance legs of the parity protocol (P2: relative RMS <= 1e-4 after 10 frames, P3:
// residuals after 1000 frames), not bit for bit.
PBD_DEV float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
PBD_DEV float rcp_fast(float a) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}
PBD_DEV float rsqrt_fast(float a) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}
// a.y*b.z - a.z*b.y with one rounding on the difference
PBD_DEV float cross_f(float ay, float bz, float az, float by) { return ffma(ay, bz, -fmul(az, by)); }
PBD_DEV float dot3_f(float ax, float ay, float az, float bx, float by, float bz) {
  return ffma(az, bz, ffma(ay, by, fmul(ax, bx)));
}

// (The same reordering as in tet_delta_fast below -- rs / (wSum + alpha) times w0, w1 formed beside the `num` chain -- was
// measured on the edges too: -0.5 %, tools/gpu_r2_exp32.sh.  The edge warps are not the ones a mixed step waits for.)
PBD_DEV bool edge_delta_fast(const float4 p0, const float4 p1, float rest, float lambda, float alpha, float4& q0,
                             float4& q1, float& newLambda) {
  const float w0 = p0.w, w1 = p1.w;
  const float wSum = fadd(w0, w1);
  const float dx = fsub(p0.x, p1.x), dy = fsub(p0.y, p1.y), dz = fsub(p0.z, p1.z);
  const float d2 = dot3_f(dx, dy, dz, dx, dy, dz);
  const float rs = rsqrt_fast(d2);                         // 1 / len
  const bool ok = (wSum != 0.0f) && !(d2 < 1e-24f);        // len < 1e-12  (Sim.cpp:112)
  const float C = ffma(d2, rs, -rest);                     // len - rest
  const float num = ffma(alpha, lambda, C);                // -(−C − alpha lambda)
  const float dl = fmul(-num, rcp_fast(fadd(wSum, alpha)));
  newLambda = fadd(lambda, dl);
  const float s = fmul(dl, rs);                            // d * s = n * dl
  const float s0 = fmul(s, w0), s1 = fmul(s, w1);
  q0.x = ffma(dx, s0, p0.x); q0.y = ffma(dy, s0, p0.y); q0.z = ffma(dz, s0, p0.z); q0.w = w0;
  q1.x = ffma(dx, -s1, p1.x); q1.y = ffma(dy, -s1, p1.y); q1.z = ffma(dz, -s1, p1.z); q1.w = w1;
  return ok;
}

// `inert` (alpha == 0 and the multiplier not carried, see pbd_sweep.cuh): alpha * lambda and the + alpha of the
// denominator drop out -- bit-neutral, one instruction less on the dependent path.
// The dependent path is what a colour step waits for (a block has ~3 tet warps per step and its barrier waits for
// them), so its length counts, not the instruction count: the three cross-product terms of wS come first and the
// -(nb + nc + nd) term, ready two additions later, last; and the four w_k * (-C / 6) are formed while the SFU computes
// the reciprocal, so that ONE multiplication and the FFMAs follow MUFU.RCP instead of three and the FFMAs: +1.2 % on the
// headline.  Measured and not kept: forming the twelve n_k w_k t products under the reciprocal as well (one FFMA after
// MUFU.RCP: -1.5 %, twelve more instructions on the warps everybody waits for); folding the 1/36 into the scalar when the
// multiplier is inert (would end the bit-identity with the kernel that carries it).
PBD_DEV bool tet_delta_fast(float4& pa, float4& pb, float4& pc, float4& pd, float rest, float lambda, float alpha,
                            float& newLambda, bool inert = false) {
  const float k6 = 1.0f / 6.0f, k36 = 1.0f / 36.0f;
  const float wa = pa.w, wb = pb.w, wc = pc.w, wd = pd.w;
  const float bax = fsub(pb.x, pa.x), bay = fsub(pb.y, pa.y), baz = fsub(pb.z, pa.z);
  const float cax = fsub(pc.x, pa.x), cay = fsub(pc.y, pa.y), caz = fsub(pc.z, pa.z);
  const float dax = fsub(pd.x, pa.x), day = fsub(pd.y, pa.y), daz = fsub(pd.z, pa.z);
  // unscaled gradients n_k = 6 g_k (Sim.cpp:147-149): nb = ca x da, nc = da x ba, nd = ba x ca
  const float nbx = cross_f(cay, daz, caz, day), nby = cross_f(caz, dax, cax, daz), nbz = cross_f(cax, day, cay, dax);
  const float ncx = cross_f(day, baz, daz, bay), ncy = cross_f(daz, bax, dax, baz), ncz = cross_f(dax, bay, day, bax);
  const float ndx = cross_f(bay, caz, baz, cay), ndy = cross_f(baz, cax, bax, caz), ndz = cross_f(bax, cay, bay, cax);
  // na = (pd - pb) x (pc - pb) = -(nb + nc + nd)   (Sim.cpp:146); ma = -na
  const float max_ = fadd(fadd(nbx, ncx), ndx), may = fadd(fadd(nby, ncy), ndy), maz = fadd(fadd(nbz, ncz), ndz);
  float wS = fmul(wb, dot3_f(nbx, nby, nbz, nbx, nby, nbz));
  wS = ffma(wc, dot3_f(ncx, ncy, ncz, ncx, ncy, ncz), wS);
  wS = ffma(wd, dot3_f(ndx, ndy, ndz, ndx, ndy, ndz), wS);
  wS = ffma(wa, dot3_f(max_, may, maz, max_, may, maz), wS);   // 36 * sum w_k |g_k|^2
  const bool ok = !(wS < 36.0e-20f);                            // also false when every w is 0 (Sim.cpp:143,155)
  const float C = ffma(dot3_f(ndx, ndy, ndz, dax, day, daz), k6, -rest);   // volume - rest
  // dl = -num / (wS / 36 + alpha); s_k = w_k dl / 6
  const float num = inert ? C : ffma(alpha, lambda, C);
  const float rc = rcp_fast(inert ? fmul(wS, k36) : ffma(wS, k36, alpha));
  const float t = fmul(num, -k6);
  const float ta = fmul(wa, t), tb = fmul(wb, t), tc = fmul(wc, t), td = fmul(wd, t);
  newLambda = fadd(lambda, fmul(-num, rc));
  const float sa = fmul(ta, rc), sb = fmul(tb, rc), sc = fmul(tc, rc), sd = fmul(td, rc);
  pa.x = ffma(max_, -sa, pa.x); pa.y = ffma(may, -sa, pa.y); pa.z = ffma(maz, -sa, pa.z);
  pb.x = ffma(nbx, sb, pb.x); pb.y = ffma(nby, sb, pb.y); pb.z = ffma(nbz, sb, pb.z);
  pc.x = ffma(ncx, sc, pc.x); pc.y = ffma(ncy, sc, pc.y); pc.z = ffma(ncz, sc, pc.z);
  pd.x = ffma(ndx, sd, pd.x); pd.y = ffma(ndy, sd, pd.y); pd.z = ffma(ndz, sd, pd.z);
  return ok;
}

// Per-frame scalars derived on the host in float exactly as the reference does.
struct StepConsts {
  float sdt;          // dt / float(substeps)                      Sim.cpp:286
  float invDt;        // sdt > 1e-12 ? 1/sdt : 0                   Sim.cpp:198
  float alphaEdge;    // max(0,edgeCompliance) * invDt2            Sim.cpp:101-102,116
  float alphaTet;     // max(0,volumeCompliance) * invDt2          Sim.cpp:133-134,162
  float gdx, gdy, gdz;  // g * sdt (the product the reference forms per vertex, Sim.cpp:182)
  float groundY;
  float groundYEps;   // groundY + 1e-6f                            Sim.cpp:212
  float fricScale;    // 1 - clamp(friction,0,1)                    Sim.cpp:200,213-214
  int groundEnabled;
};

// predict for one vertex (Sim.cpp:180-184).  x: committed position, v: velocity (updated).
PBD_DEV float4 predict_vertex(const float4 x, float4& v, float w, const StepConsts& k) {
  float4 p;
  p.w = w;
  if (w == 0.0f) { p.x = x.x; p.y = x.y; p.z = x.z; return p; }
  v.x = fadd(v.x, k.gdx); v.y = fadd(v.y, k.gdy); v.z = fadd(v.z, k.gdz);
  p.x = fadd(x.x, fmul(v.x, k.sdt)); p.y = fadd(x.y, fmul(v.y, k.sdt)); p.z = fadd(x.z, fmul(v.z, k.sdt));
  return p;
}

// ground clamp for one vertex (Sim.cpp:190-194)
PBD_DEV void ground_vertex(float4& p, const StepConsts& k) {
  if (k.groundEnabled && p.w != 0.0f && p.y < k.groundY) p.y = k.groundY;
}

// ---- primitive colliders in the clamp stage ----------------------------------
// Sphere / oriented box / capsule push-out of the reference's in-engine solver, restated from
// Assets/Scripts/Softbody/SoftBodyCollisionMath.cs:8-110 (ComputePushOut, PushOutSphere :24-41,
// PushOutBox :45-90, PushOutCapsule :93-110) and applied where that solver applies it
// (SoftBodySolver.cs:554-561 == SoftBodyCompute.compute:410-430: after the ground plane, colliders in
// order, `if (hit) p += push`).  float32, C# evaluation order, no FMA; the quaternion product is
// Unity's `Quaternion * Vector3`.  Quaternion.Inverse is taken as the conjugate (unit rotations).
// PBDServer itself has only the y-plane (Sim.cpp:187-195): with no collider set this is dead code.
struct Collider {              // == pbd_collider (include/pbd_b200.h), 48 bytes
  uint32_t type;               // 0 sphere, 1 box, 2 capsule (SoftBodyPrimitiveCollider.PrimitiveType)
  float px, py, pz;            // positionW
  float qx, qy, qz, qw;        // rotationW
  float dx, dy, dz;            // sphere: radius | box: half extents | capsule: radius, half height
};
constexpr uint32_t kMaxColliders = 16;
struct ColliderSet {
  uint32_t n;
  float particleRadius;
  uint32_t pad[2];
  Collider c[kMaxColliders];
};

PBD_DEV void quat_rotate(float qx, float qy, float qz, float qw, float vx, float vy, float vz, float& ox, float& oy, float& oz) {
  const float x = fmul(qx, 2.0f), y = fmul(qy, 2.0f), z = fmul(qz, 2.0f);
  const float xx = fmul(qx, x), yy = fmul(qy, y), zz = fmul(qz, z);
  const float xy = fmul(qx, y), xz = fmul(qx, z), yz = fmul(qy, z);
  const float wx = fmul(qw, x), wy = fmul(qw, y), wz = fmul(qw, z);
  ox = fadd(fadd(fmul(fsub(1.0f, fadd(yy, zz)), vx), fmul(fsub(xy, wz), vy)), fmul(fadd(xz, wy), vz));
  oy = fadd(fadd(fmul(fadd(xy, wz), vx), fmul(fsub(1.0f, fadd(xx, zz)), vy)), fmul(fsub(yz, wx), vz));
  oz = fadd(fadd(fmul(fsub(xz, wy), vx), fmul(fadd(yz, wx), vy)), fmul(fsub(1.0f, fadd(xx, yy)), vz));
}
// SoftBodyCollisionMath.cs:24-41
PBD_DEV bool push_out_sphere(float cx, float cy, float cz, float radius, float px, float py, float pz, float& ox, float& oy, float& oz) {
  const float vx = fsub(px, cx), vy = fsub(py, cy), vz = fsub(pz, cz);
  const float d2 = dot3(vx, vy, vz, vx, vy, vz);
  const float r = fmaxf(1e-6f, radius);
  if (d2 >= fmul(r, r)) return false;
  const float d = __fsqrt_rn(fmaxf(d2, 1e-20f));
  float nx = 0.0f, ny = 1.0f, nz = 0.0f;   // Vector3.up
  if (d > 1e-10f) { nx = fdiv(vx, d); ny = fdiv(vy, d); nz = fdiv(vz, d); }
  const float k = fsub(r, d);
  ox = fmul(nx, k); oy = fmul(ny, k); oz = fmul(nz, k);
  return true;
}
PBD_DEV bool push_out(const Collider& c, float pr, float px, float py, float pz, float& ox, float& oy, float& oz) {
  if (c.type == 0u) return push_out_sphere(c.px, c.py, c.pz, fadd(c.dx, pr), px, py, pz, ox, oy, oz);
  if (c.type == 1u) {   // SoftBodyCollisionMath.cs:45-90
    float lx, ly, lz;
    quat_rotate(-c.qx, -c.qy, -c.qz, c.qw, fsub(px, c.px), fsub(py, c.py), fsub(pz, c.pz), lx, ly, lz);
    const float ex = fadd(c.dx, pr), ey = fadd(c.dy, pr), ez = fadd(c.dz, pr);
    if (!(fabsf(lx) <= ex && fabsf(ly) <= ey && fabsf(lz) <= ez)) return false;
    const float dx = fsub(ex, fabsf(lx)), dy = fsub(ey, fabsf(ly)), dz = fsub(ez, fabsf(lz));
    float ux = 0.0f, uy = 0.0f, uz = 0.0f;
    if (dx <= dy && dx <= dz) ux = fmul(dx, lx >= 0.0f ? 1.0f : -1.0f);
    else if (dy <= dz) uy = fmul(dy, ly >= 0.0f ? 1.0f : -1.0f);
    else uz = fmul(dz, lz >= 0.0f ? 1.0f : -1.0f);
    quat_rotate(c.qx, c.qy, c.qz, c.qw, ux, uy, uz, ox, oy, oz);
    return true;
  }
  // capsule, axis = local Y: SoftBodyCollisionMath.cs:93-110
  const float r = fmaxf(1e-6f, fadd(c.dx, pr)), h = fmaxf(0.0f, c.dy);
  float ux, uy, uz;
  quat_rotate(c.qx, c.qy, c.qz, c.qw, 0.0f, 1.0f, 0.0f, ux, uy, uz);
  const float ax = fsub(c.px, fmul(ux, h)), ay = fsub(c.py, fmul(uy, h)), az = fsub(c.pz, fmul(uz, h));
  const float bx = fadd(c.px, fmul(ux, h)), by = fadd(c.py, fmul(uy, h)), bz = fadd(c.pz, fmul(uz, h));
  const float abx = fsub(bx, ax), aby = fsub(by, ay), abz = fsub(bz, az);
  const float ab2 = dot3(abx, aby, abz, abx, aby, abz);
  float t = 0.0f;
  if (ab2 > 1e-20f) {
    t = fdiv(dot3(fsub(px, ax), fsub(py, ay), fsub(pz, az), abx, aby, abz), ab2);
    t = fminf(fmaxf(t, 0.0f), 1.0f);
  }
  return push_out_sphere(fadd(ax, fmul(abx, t)), fadd(ay, fmul(aby, t)), fadd(az, fmul(abz, t)), r, px, py, pz, ox, oy, oz);
}
// every collider in order, on a vertex with mass (SoftBodyCompute.compute:396-397: invMass == 0 -> untouched)
// (out of line: the frame kernels reach it from several unrolled vertex stages; inlined it was a fifth of their instructions
// for a stage that runs only when a collider is set)
static __device__ __noinline__ void collide_vertex(float4& p, const ColliderSet* cs, uint32_t n) {
  if (p.w == 0.0f) return;
  const float pr = fmaxf(1e-6f, cs->particleRadius);   // SoftBodySolver.cs:546
  for (uint32_t i = 0; i < n; ++i) {
    const Collider c = cs->c[i];
    float ox, oy, oz;
    if (push_out(c, pr, p.x, p.y, p.z, ox, oy, oz)) { p.x = fadd(p.x, ox); p.y = fadd(p.y, oy); p.z = fadd(p.z, oz); }
  }
}

// commit for one vertex (Sim.cpp:202-221).  p: xStar (+w); x: committed position, updated; v out.
PBD_DEV void commit_vertex(float4& p, float4& x, float4& v, const StepConsts& k) {
  if (p.w == 0.0f) {
    v.x = v.y = v.z = 0.0f;
    p.x = x.x; p.y = x.y; p.z = x.z;
    return;
  }
  float vx = fmul(fsub(p.x, x.x), k.invDt), vy = fmul(fsub(p.y, x.y), k.invDt), vz = fmul(fsub(p.z, x.z), k.invDt);
  if (k.groundEnabled && p.y <= k.groundYEps) {
    vx = fmul(vx, k.fricScale);
    vz = fmul(vz, k.fricScale);
    if (vy < 0.0f) vy = 0.0f;
  }
  v.x = vx; v.y = vy; v.z = vz;
  x.x = p.x; x.y = p.y; x.z = p.z;
}

}  // namespace pbd

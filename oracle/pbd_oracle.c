/*
 * pbd_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference PBDServer XPBD substep
 * (Captain-Noble/CS121-softbodysim, CProgram/).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product path (cs121-softbodysim_b200/csrc) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_cpu.py checks this port bit-for-bit
 * against (a) oracle/_ref/libpbdref.so = the unmodified reference Sim.cpp compiled
 * where it lies (when present) and (b) tests/golden/ref_*.npz, outputs of that same
 * reference generated in the build container by tests/golden/make_golden.py.
 *
 * Every function names the reference lines it restates (paths relative to
 * /root/reference).  All arithmetic is IEEE binary32 in the reference's evaluation
 * order; build with -ffp-contract=off so no FMA is formed (the reference is built for
 * baseline x86-64, SSE2 scalar, no FMA).
 *
 * Layout differs from the reference on purpose (flat float arrays, x/y/z interleaved)
 * -- it is a restatement of the algorithm, not of the source text.
 */
#define _POSIX_C_SOURCE 199309L
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct pbdo_params {
  /* field order == MSG_INIT wire order, CProgram/src/Server.cpp:38-50 */
  uint32_t substeps, iterations;
  float dtHint, omega;
  float edgeCompliance, volumeCompliance;
  float gx, gy, gz;
  uint32_t groundEnabled;
  float groundY, friction;
} pbdo_params;

typedef struct pbdo_state {
  uint32_t V, E, T;
  pbdo_params prm;
  float *x, *v, *xs;        /* 3V each: committed position, velocity, predicted (xStar) */
  float *w;                 /* V inverse masses                                          */
  uint32_t *e0, *e1;        /* E                                                         */
  float *eRest, *eLam;      /* E                                                         */
  uint32_t *ta, *tb, *tc, *td; /* T                                                      */
  float *tRest, *tLam;      /* T                                                         */
  double ms_predict, ms_solve, ms_commit, ms_total; /* accumulated like perf::StepStats  */
  /* primitive colliders of the clamp stage; 0 = PBDServer's behaviour */
  uint32_t nColliders;
  float particleRadius;
  struct pbdo_collider { uint32_t type; float p[3]; float q[4]; float d[3]; } colliders[16];
} pbdo_state;

static double now_ms(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec * 1e3 + (double)ts.tv_nsec * 1e-6;
}

/* CProgram/include/PBDServer.h:140-145 -- signed volume, true division by 6.0f. */
static float signed_volume(const float *p0, const float *p1, const float *p2, const float *p3) {
  float ax = p1[0] - p0[0], ay = p1[1] - p0[1], az = p1[2] - p0[2];
  float bx = p2[0] - p0[0], by = p2[1] - p0[1], bz = p2[2] - p0[2];
  float cx = p3[0] - p0[0], cy = p3[1] - p0[1], cz = p3[2] - p0[2];
  /* cross(a,b) as in PBDServer.h:134-136, then dot with c (PBDServer.h:133) */
  float nx = ay * bz - az * by;
  float ny = az * bx - ax * bz;
  float nz = ax * by - ay * bx;
  return (nx * cx + ny * cy + nz * cz) / 6.0f;
}

/* CProgram/src/Sim.cpp:63-79 -- w_i = sum over tets (in array order) of 4/|vol|, pinned -> 0. */
static void inv_mass(pbdo_state *s, const uint32_t *pinned, uint32_t nPinned) {
  for (uint32_t i = 0; i < s->V; ++i) s->w[i] = 0.0f;
  for (uint32_t t = 0; t < s->T; ++t) {
    uint32_t a = s->ta[t], b = s->tb[t], c = s->tc[t], d = s->td[t];
    float vol = signed_volume(s->x + 3 * a, s->x + 3 * b, s->x + 3 * c, s->x + 3 * d);
    float m = fabsf(vol);
    if (m > 1e-12f) {
      float inv = 4.0f / m;
      s->w[a] += inv; s->w[b] += inv; s->w[c] += inv; s->w[d] += inv;
    }
  }
  for (uint32_t k = 0; k < nPinned; ++k)
    if (pinned[k] < s->V) s->w[pinned[k]] = 0.0f;
}

/* CProgram/src/Sim.cpp:81-95 -- rest lengths / signed rest volumes, lambdas := 0. */
static void rest_state(pbdo_state *s) {
  for (uint32_t e = 0; e < s->E; ++e) {
    const float *p0 = s->x + 3 * s->e0[e], *p1 = s->x + 3 * s->e1[e];
    float dx = p1[0] - p0[0], dy = p1[1] - p0[1], dz = p1[2] - p0[2];
    s->eRest[e] = sqrtf(dx * dx + dy * dy + dz * dz);
    s->eLam[e] = 0.0f;
  }
  for (uint32_t t = 0; t < s->T; ++t) {
    s->tRest[t] = signed_volume(s->x + 3 * s->ta[t], s->x + 3 * s->tb[t],
                                s->x + 3 * s->tc[t], s->x + 3 * s->td[t]);
    s->tLam[t] = 0.0f;
  }
}

/* One edge projection: body of the loop at CProgram/src/Sim.cpp:104-129. */
static void project_edge(pbdo_state *s, uint32_t e, float alpha) {
  uint32_t i0 = s->e0[e], i1 = s->e1[e];
  float w0 = s->w[i0], w1 = s->w[i1];
  float wSum = w0 + w1;
  if (wSum == 0.0f) return;
  float *p0 = s->xs + 3 * i0, *p1 = s->xs + 3 * i1;
  float dx = p0[0] - p1[0], dy = p0[1] - p1[1], dz = p0[2] - p1[2];
  float len = sqrtf(dx * dx + dy * dy + dz * dz);
  if (len < 1e-12f) return;
  float C = len - s->eRest[e];
  float lam = s->eLam[e];
  float dl = (-C - alpha * lam) / (wSum + alpha);
  s->eLam[e] = lam + dl;
  float inv = 1.0f / len;
  float cx = (dx * inv) * dl, cy = (dy * inv) * dl, cz = (dz * inv) * dl;
  p0[0] = p0[0] + cx * w0; p0[1] = p0[1] + cy * w0; p0[2] = p0[2] + cz * w0;
  p1[0] = p1[0] - cx * w1; p1[1] = p1[1] - cy * w1; p1[2] = p1[2] - cz * w1;
}

/* One tet-volume projection: body of the loop at CProgram/src/Sim.cpp:136-172. */
static void project_tet(pbdo_state *s, uint32_t t, float alpha) {
  const float k6 = 1.0f / 6.0f;             /* multiply, unlike signed_volume (Sim.cpp:146-149) */
  uint32_t a = s->ta[t], b = s->tb[t], c = s->tc[t], d = s->td[t];
  float wa = s->w[a], wb = s->w[b], wc = s->w[c], wd = s->w[d];
  if (wa + wb + wc + wd == 0.0f) return;
  float *pa = s->xs + 3 * a, *pb = s->xs + 3 * b, *pc = s->xs + 3 * c, *pd = s->xs + 3 * d;

  float dbx = pd[0] - pb[0], dby = pd[1] - pb[1], dbz = pd[2] - pb[2]; /* pd - pb */
  float cbx = pc[0] - pb[0], cby = pc[1] - pb[1], cbz = pc[2] - pb[2]; /* pc - pb */
  float cax = pc[0] - pa[0], cay = pc[1] - pa[1], caz = pc[2] - pa[2]; /* pc - pa */
  float dax = pd[0] - pa[0], day = pd[1] - pa[1], daz = pd[2] - pa[2]; /* pd - pa */
  float bax = pb[0] - pa[0], bay = pb[1] - pa[1], baz = pb[2] - pa[2]; /* pb - pa */

  /* ga = cross(pd-pb, pc-pb)/6 ; gb = cross(pc-pa, pd-pa)/6 ;
     gc = cross(pd-pa, pb-pa)/6 ; gd = cross(pb-pa, pc-pa)/6   (Sim.cpp:146-149) */
  float gax = (dby * cbz - dbz * cby) * k6, gay = (dbz * cbx - dbx * cbz) * k6, gaz = (dbx * cby - dby * cbx) * k6;
  float gbx = (cay * daz - caz * day) * k6, gby = (caz * dax - cax * daz) * k6, gbz = (cax * day - cay * dax) * k6;
  float gcx = (day * baz - daz * bay) * k6, gcy = (daz * bax - dax * baz) * k6, gcz = (dax * bay - day * bax) * k6;
  float gdx = (bay * caz - baz * cay) * k6, gdy = (baz * cax - bax * caz) * k6, gdz = (bax * cay - bay * cax) * k6;

  float wSum = wa * (gax * gax + gay * gay + gaz * gaz) +
               wb * (gbx * gbx + gby * gby + gbz * gbz) +
               wc * (gcx * gcx + gcy * gcy + gcz * gcz) +
               wd * (gdx * gdx + gdy * gdy + gdz * gdz);
  if (wSum < 1e-20f) return;

  /* tet_volume(pa,pb,pc,pd): dot(cross(pb-pa, pc-pa), pd-pa) / 6.0f  (Sim.cpp:159) */
  float nx = bay * caz - baz * cay;
  float ny = baz * cax - bax * caz;
  float nz = bax * cay - bay * cax;
  float vol = (nx * dax + ny * day + nz * daz) / 6.0f;
  float C = vol - s->tRest[t];

  float lam = s->tLam[t];
  float dl = (-C - alpha * lam) / (wSum + alpha);
  s->tLam[t] = lam + dl;

  float sa = wa * dl, sb = wb * dl, sc = wc * dl, sd = wd * dl;
  pa[0] = pa[0] + gax * sa; pa[1] = pa[1] + gay * sa; pa[2] = pa[2] + gaz * sa;
  pb[0] = pb[0] + gbx * sb; pb[1] = pb[1] + gby * sb; pb[2] = pb[2] + gbz * sb;
  pc[0] = pc[0] + gcx * sc; pc[1] = pc[1] + gcy * sc; pc[2] = pc[2] + gcz * sc;
  pd[0] = pd[0] + gdx * sd; pd[1] = pd[1] + gdy * sd; pd[2] = pd[2] + gdz * sd;
}

/* alpha = max(0,compliance) * (dt>1e-12 ? 1/(dt*dt) : 0)   (Sim.cpp:101-102,116 / 133-134,162) */
static float xpbd_alpha(float compliance, float dt) {
  float invDt2 = (dt > 1e-12f) ? (1.0f / (dt * dt)) : 0.0f;
  float comp = (compliance > 0.0f) ? compliance : 0.0f;
  return comp * invDt2;
}

/* CProgram/src/Sim.cpp:178-185 */
static void predict(pbdo_state *s, float dt) {
  float gx = s->prm.gx, gy = s->prm.gy, gz = s->prm.gz;
  for (uint32_t i = 0; i < s->V; ++i) {
    float *x = s->x + 3 * i, *v = s->v + 3 * i, *p = s->xs + 3 * i;
    if (s->w[i] == 0.0f) { p[0] = x[0]; p[1] = x[1]; p[2] = x[2]; continue; }
    v[0] = v[0] + gx * dt; v[1] = v[1] + gy * dt; v[2] = v[2] + gz * dt;
    p[0] = x[0] + v[0] * dt; p[1] = x[1] + v[1] * dt; p[2] = x[2] + v[2] * dt;
  }
}

/* CProgram/src/Sim.cpp:187-195 */
static void ground(pbdo_state *s) {
  if (!s->prm.groundEnabled) return;
  float y0 = s->prm.groundY;
  for (uint32_t i = 0; i < s->V; ++i) {
    if (s->w[i] == 0.0f) continue;
    if (s->xs[3 * i + 1] < y0) s->xs[3 * i + 1] = y0;
  }
}

/* ---- primitive colliders.  Parity status of THIS part: UNPINNED -- the formulas
 * live in the reference's Unity C# (Assets/Scripts/Softbody/SoftBodyCollisionMath.cs:8-110; HLSL twin
 * Assets/Shaders/SoftBodyCompute.compute:108-204) and no C# / HLSL toolchain exists in this image, so
 * this restatement cannot be run against them; it is checked against hand-computed known answers
 * (tests/test_oracle_cpu.py) and is what the GPU clamp stage is compared with bit for bit.
 * float32, C# evaluation order; Quaternion * Vector3 as Unity implements it; Quaternion.Inverse of a
 * unit rotation = its conjugate. */
static void quat_rotate(const float q[4], const float v[3], float o[3]) {
  float x = q[0] * 2.0f, y = q[1] * 2.0f, z = q[2] * 2.0f;
  float xx = q[0] * x, yy = q[1] * y, zz = q[2] * z;
  float xy = q[0] * y, xz = q[0] * z, yz = q[1] * z;
  float wx = q[3] * x, wy = q[3] * y, wz = q[3] * z;
  o[0] = (1.0f - (yy + zz)) * v[0] + (xy - wz) * v[1] + (xz + wy) * v[2];
  o[1] = (xy + wz) * v[0] + (1.0f - (xx + zz)) * v[1] + (yz - wx) * v[2];
  o[2] = (xz - wy) * v[0] + (yz + wx) * v[1] + (1.0f - (xx + yy)) * v[2];
}
/* PushOutSphere, SoftBodyCollisionMath.cs:24-41 */
static int push_out_sphere(const float c[3], float radius, const float p[3], float push[3]) {
  float v[3] = {p[0] - c[0], p[1] - c[1], p[2] - c[2]};
  float d2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  float r = fmaxf(1e-6f, radius);
  if (d2 >= r * r) return 0;
  float d = sqrtf(fmaxf(d2, 1e-20f));
  float n[3] = {0.0f, 1.0f, 0.0f};
  if (d > 1e-10f) { n[0] = v[0] / d; n[1] = v[1] / d; n[2] = v[2] / d; }
  float k = r - d;
  push[0] = n[0] * k; push[1] = n[1] * k; push[2] = n[2] * k;
  return 1;
}
/* ComputePushOut :8-21, PushOutBox :45-90, PushOutCapsule :93-110 */
static int push_out(const struct pbdo_collider *c, float pr, const float p[3], float push[3]) {
  if (c->type == 0u) return push_out_sphere(c->p, c->d[0] + pr, p, push);
  if (c->type == 1u) {
    float inv[4] = {-c->q[0], -c->q[1], -c->q[2], c->q[3]};
    float rel[3] = {p[0] - c->p[0], p[1] - c->p[1], p[2] - c->p[2]}, l[3];
    quat_rotate(inv, rel, l);
    float ex = c->d[0] + pr, ey = c->d[1] + pr, ez = c->d[2] + pr;
    if (!(fabsf(l[0]) <= ex && fabsf(l[1]) <= ey && fabsf(l[2]) <= ez)) return 0;
    float dx = ex - fabsf(l[0]), dy = ey - fabsf(l[1]), dz = ez - fabsf(l[2]);
    float u[3] = {0.0f, 0.0f, 0.0f};
    if (dx <= dy && dx <= dz) u[0] = dx * (l[0] >= 0.0f ? 1.0f : -1.0f);
    else if (dy <= dz) u[1] = dy * (l[1] >= 0.0f ? 1.0f : -1.0f);
    else u[2] = dz * (l[2] >= 0.0f ? 1.0f : -1.0f);
    quat_rotate(c->q, u, push);
    return 1;
  }
  float r = fmaxf(1e-6f, c->d[0] + pr), h = fmaxf(0.0f, c->d[1]);
  float yv[3] = {0.0f, 1.0f, 0.0f}, up[3];
  quat_rotate(c->q, yv, up);
  float a[3] = {c->p[0] - up[0] * h, c->p[1] - up[1] * h, c->p[2] - up[2] * h};
  float b[3] = {c->p[0] + up[0] * h, c->p[1] + up[1] * h, c->p[2] + up[2] * h};
  float ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
  float ab2 = ab[0] * ab[0] + ab[1] * ab[1] + ab[2] * ab[2];
  float t = 0.0f;
  if (ab2 > 1e-20f) {
    t = ((p[0] - a[0]) * ab[0] + (p[1] - a[1]) * ab[1] + (p[2] - a[2]) * ab[2]) / ab2;
    t = fminf(fmaxf(t, 0.0f), 1.0f);
  }
  float cc[3] = {a[0] + ab[0] * t, a[1] + ab[1] * t, a[2] + ab[2] * t};
  return push_out_sphere(cc, r, p, push);
}
/* SoftBodySolver.cs:554-561: after the ground plane, every collider in order, `if (hit) p += push` */
static void collide(pbdo_state *s) {
  if (!s->nColliders) return;
  float pr = fmaxf(1e-6f, s->particleRadius);
  for (uint32_t i = 0; i < s->V; ++i) {
    if (s->w[i] == 0.0f) continue;
    float *p = s->xs + 3 * i;
    for (uint32_t c = 0; c < s->nColliders; ++c) {
      float push[3];
      if (push_out(&s->colliders[c], pr, p, push)) { p[0] = p[0] + push[0]; p[1] = p[1] + push[1]; p[2] = p[2] + push[2]; }
    }
  }
}
void pbdo_set_colliders(pbdo_state *s, const void *cols, uint32_t n, float particleRadius) {
  s->nColliders = n > 16u ? 16u : n;
  s->particleRadius = particleRadius;
  if (s->nColliders) memcpy(s->colliders, cols, sizeof(s->colliders[0]) * s->nColliders);
}
/* one point against one collider (known-answer tests) */
int pbdo_push_out(const void *col, float particleRadius, const float *p, float *push) {
  return push_out((const struct pbdo_collider *)col, fmaxf(1e-6f, particleRadius), p, push);
}

/* CProgram/src/Sim.cpp:197-222 */
static void commit(pbdo_state *s, float dt) {
  float invDt = (dt > 1e-12f) ? (1.0f / dt) : 0.0f;
  float y0 = s->prm.groundY;
  float fr = fmaxf(0.0f, fminf(1.0f, s->prm.friction));
  for (uint32_t i = 0; i < s->V; ++i) {
    float *x = s->x + 3 * i, *v = s->v + 3 * i, *p = s->xs + 3 * i;
    if (s->w[i] == 0.0f) {
      v[0] = v[1] = v[2] = 0.0f;
      p[0] = x[0]; p[1] = x[1]; p[2] = x[2];
      continue;
    }
    float vx = (p[0] - x[0]) * invDt, vy = (p[1] - x[1]) * invDt, vz = (p[2] - x[2]) * invDt;
    if (s->prm.groundEnabled && p[1] <= y0 + 1e-6f) {
      vx *= (1.0f - fr);
      vz *= (1.0f - fr);
      if (vy < 0.0f) vy = 0.0f;
    }
    v[0] = vx; v[1] = vy; v[2] = vz;
    x[0] = p[0]; x[1] = p[1]; x[2] = p[2];
  }
}

/* ------------------------------------------------------------------ public API */

pbdo_state *pbdo_create(const pbdo_params *prm, uint32_t V, uint32_t E, uint32_t T,
                        const float *x0, const uint32_t *edgeIds, const uint32_t *tetIds,
                        const uint32_t *pinned, uint32_t nPinned) {
  /* mirrors the MSG_INIT decode, CProgram/src/Server.cpp:72-104 */
  pbdo_state *s = (pbdo_state *)calloc(1, sizeof(*s));
  s->V = V; s->E = E; s->T = T; s->prm = *prm;
  size_t v3 = (size_t)V * 3;
  s->x = (float *)malloc(sizeof(float) * (v3 + 1));
  s->v = (float *)calloc(v3 + 1, sizeof(float));
  s->xs = (float *)malloc(sizeof(float) * (v3 + 1));
  s->w = (float *)calloc((size_t)V + 1, sizeof(float));
  memcpy(s->x, x0, sizeof(float) * v3);
  memcpy(s->xs, x0, sizeof(float) * v3);
  s->e0 = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)E + 1));
  s->e1 = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)E + 1));
  s->eRest = (float *)malloc(sizeof(float) * ((size_t)E + 1));
  s->eLam = (float *)malloc(sizeof(float) * ((size_t)E + 1));
  for (uint32_t e = 0; e < E; ++e) { s->e0[e] = edgeIds[2 * e]; s->e1[e] = edgeIds[2 * e + 1]; }
  s->ta = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)T + 1));
  s->tb = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)T + 1));
  s->tc = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)T + 1));
  s->td = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)T + 1));
  s->tRest = (float *)malloc(sizeof(float) * ((size_t)T + 1));
  s->tLam = (float *)malloc(sizeof(float) * ((size_t)T + 1));
  for (uint32_t t = 0; t < T; ++t) {
    s->ta[t] = tetIds[4 * t]; s->tb[t] = tetIds[4 * t + 1];
    s->tc[t] = tetIds[4 * t + 2]; s->td[t] = tetIds[4 * t + 3];
  }
  inv_mass(s, pinned, nPinned);
  rest_state(s);
  return s;
}

void pbdo_destroy(pbdo_state *s) {
  if (!s) return;
  free(s->x); free(s->v); free(s->xs); free(s->w);
  free(s->e0); free(s->e1); free(s->eRest); free(s->eLam);
  free(s->ta); free(s->tb); free(s->tc); free(s->td); free(s->tRest); free(s->tLam);
  free(s);
}

/* Re-order the constraint arrays (indices, rest values, lambdas) AFTER w and the rest
 * state were derived in the caller's original order: new[k] = old[order[k]].  This is the
 * "same-order" device of SURVEY.md 8(c): the sequential sweep over a colour/tile-sorted
 * list computes exactly what a conflict-free parallel schedule must compute.            */
static void permute_u32(uint32_t *a, const uint32_t *order, uint32_t n) {
  uint32_t *t = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)n + 1));
  for (uint32_t k = 0; k < n; ++k) t[k] = a[order[k]];
  memcpy(a, t, sizeof(uint32_t) * n); free(t);
}
static void permute_f32(float *a, const uint32_t *order, uint32_t n) {
  float *t = (float *)malloc(sizeof(float) * ((size_t)n + 1));
  for (uint32_t k = 0; k < n; ++k) t[k] = a[order[k]];
  memcpy(a, t, sizeof(float) * n); free(t);
}
void pbdo_permute_constraints(pbdo_state *s, const uint32_t *edgeOrder, const uint32_t *tetOrder) {
  if (edgeOrder) {
    permute_u32(s->e0, edgeOrder, s->E); permute_u32(s->e1, edgeOrder, s->E);
    permute_f32(s->eRest, edgeOrder, s->E); permute_f32(s->eLam, edgeOrder, s->E);
  }
  if (tetOrder) {
    permute_u32(s->ta, tetOrder, s->T); permute_u32(s->tb, tetOrder, s->T);
    permute_u32(s->tc, tetOrder, s->T); permute_u32(s->td, tetOrder, s->T);
    permute_f32(s->tRest, tetOrder, s->T); permute_f32(s->tLam, tetOrder, s->T);
  }
}

/* The frame step: loop nest of SerialStepper::step, CProgram/src/Sim.cpp:280-305. */
void pbdo_step(pbdo_state *s, float dt) {
  double tAll = now_ms();
  uint32_t ss = s->prm.substeps > 1u ? s->prm.substeps : 1u;
  float sdt = dt / (float)ss;
  for (uint32_t k = 0; k < ss; ++k) {
    double t0 = now_ms();
    predict(s, sdt);
    double t1 = now_ms();
    for (uint32_t it = 0; it < s->prm.iterations; ++it) {
      float aE = xpbd_alpha(s->prm.edgeCompliance, sdt);
      for (uint32_t e = 0; e < s->E; ++e) project_edge(s, e, aE);
      float aT = xpbd_alpha(s->prm.volumeCompliance, sdt);
      for (uint32_t t = 0; t < s->T; ++t) project_tet(s, t, aT);
      ground(s);
      collide(s);
    }
    double t2 = now_ms();
    commit(s, sdt);
    double t3 = now_ms();
    s->ms_predict += t1 - t0; s->ms_solve += t2 - t1; s->ms_commit += t3 - t2;
  }
  s->ms_total += now_ms() - tAll;
}

/* Generalised sweep used only to check the GPU's "interleaved" schedule: per iteration the
 * items are projected in the given sequence (bit 31 set = tet index, clear = edge index,
 * indices refer to the CURRENT array order), then the ground clamp.  With items = all edges
 * then all tets this is pbdo_step.                                                       */
void pbdo_step_sequence(pbdo_state *s, float dt, const uint32_t *items, uint64_t nItems) {
  uint32_t ss = s->prm.substeps > 1u ? s->prm.substeps : 1u;
  float sdt = dt / (float)ss;
  float aE = xpbd_alpha(s->prm.edgeCompliance, sdt);
  float aT = xpbd_alpha(s->prm.volumeCompliance, sdt);
  for (uint32_t k = 0; k < ss; ++k) {
    predict(s, sdt);
    for (uint32_t it = 0; it < s->prm.iterations; ++it) {
      for (uint64_t j = 0; j < nItems; ++j) {
        uint32_t id = items[j];
        if (id & 0x80000000u) project_tet(s, id & 0x7fffffffu, aT);
        else project_edge(s, id, aE);
      }
      ground(s);
      collide(s);
    }
    commit(s, sdt);
  }
}

/* pack_positions, CProgram/src/Sim.cpp:307-316: committed x of ALL vertices, input order. */
void pbdo_pack(const pbdo_state *s, float *out) { memcpy(out, s->x, sizeof(float) * 3 * (size_t)s->V); }

void pbdo_get(const pbdo_state *s, int what, void *out) {
  switch (what) {
    case 0: memcpy(out, s->w, sizeof(float) * s->V); break;
    case 1: memcpy(out, s->eRest, sizeof(float) * s->E); break;
    case 2: memcpy(out, s->tRest, sizeof(float) * s->T); break;
    case 3: memcpy(out, s->eLam, sizeof(float) * s->E); break;
    case 4: memcpy(out, s->tLam, sizeof(float) * s->T); break;
    case 5: memcpy(out, s->v, sizeof(float) * 3 * (size_t)s->V); break;
    case 6: memcpy(out, s->xs, sizeof(float) * 3 * (size_t)s->V); break;
    default: break;
  }
}

/* a later MSG_INIT would rebuild the state; the C ABI also lets a caller change SolverParams in
 * place (pbd_set_params) -- mirrored here so that path can be checked */
void pbdo_set_params(pbdo_state *s, const pbdo_params *prm) { s->prm = *prm; }

void pbdo_set_inv_mass(pbdo_state *s, const float *w) { memcpy(s->w, w, sizeof(float) * s->V); }

void pbdo_stats(pbdo_state *s, double *out5, int reset) {
  out5[0] = s->ms_predict; out5[1] = s->ms_solve; out5[2] = s->ms_commit; out5[3] = 0.0; out5[4] = s->ms_total;
  if (reset) s->ms_predict = s->ms_solve = s->ms_commit = s->ms_total = 0.0;
}

const char *pbdo_name(void) { return "oracle-port"; }

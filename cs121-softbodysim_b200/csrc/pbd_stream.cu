// pbd_stream.cu -- "stream" backend: globally graph-coloured Gauss-Seidel, one launch per colour.
//
// The simple, obviously-correct schedule: constraints of one global colour
// share no vertex, so one thread per constraint gathers its float4 vertices from HBM/L2,
// projects, and scatters them back without atomics.  A whole frame
//     S x [ predict ; I x ( nEdgeColours launches ; nTetColours launches ; ground ) ; commit ]
// is captured once into a CUDA graph (the per-frame scalars live in device memory, so a change
// of dt does not invalidate it).  This backend is the right tool when each colour moves
// tens of MB (very large meshes) and the bring-up / fallback path otherwise; the L2-resident
// 1M-tet headline config is phase-count bound here and is served by the tile backend.
//
// Reference functions these kernels replace (CProgram/src/Sim.cpp): predict_serial :178-185,
// solve_edges_xpbd_gs :100-130, solve_tets_xpbd_gs :132-173, project_ground_serial :187-195,
// commit_serial :197-222, SerialStepper::pack_positions :307-316.
#include <vector>

#include "pbd_body.h"

namespace pbd {

namespace {

constexpr int kBlock = 256;

__global__ void __launch_bounds__(kBlock) predict_kernel(float4* __restrict__ pos, const float4* __restrict__ prev,
                                                         float4* __restrict__ vel, uint32_t V,
                                                         const StepConsts* __restrict__ kc) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V) return;
  const StepConsts k = *kc;
  const float w = pos[i].w;
  const float4 x = prev[i];
  if (w == 0.0f) {
    pos[i] = make_float4(x.x, x.y, x.z, w);
    return;
  }
  float4 v = vel[i];
  const float4 p = predict_vertex(x, v, w, k);
  vel[i] = v;
  pos[i] = p;
}

__global__ void __launch_bounds__(kBlock) edge_colour_kernel(float4* __restrict__ pos, const uint2* __restrict__ idx,
                                                             const float* __restrict__ rest, float* __restrict__ lam,
                                                             uint32_t begin, uint32_t count,
                                                             const StepConsts* __restrict__ kc) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= count) return;
  const uint32_t e = begin + j;
  const uint2 id = idx[e];
  float4 p0 = pos[id.x], p1 = pos[id.y];
  float l = lam[e];
  if (project_edge(p0, p1, rest[e], l, kc->alphaEdge)) {
    lam[e] = l;
    pos[id.x] = p0;
    pos[id.y] = p1;
  }
}

__global__ void __launch_bounds__(kBlock) tet_colour_kernel(float4* __restrict__ pos, const uint4* __restrict__ idx,
                                                            const float* __restrict__ rest, float* __restrict__ lam,
                                                            uint32_t begin, uint32_t count,
                                                            const StepConsts* __restrict__ kc) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= count) return;
  const uint32_t t = begin + j;
  const uint4 id = idx[t];
  float4 pa = pos[id.x], pb = pos[id.y], pc = pos[id.z], pd = pos[id.w];
  float l = lam[t];
  if (project_tet(pa, pb, pc, pd, rest[t], l, kc->alphaTet)) {
    lam[t] = l;
    pos[id.x] = pa;
    pos[id.y] = pb;
    pos[id.z] = pc;
    pos[id.w] = pd;
  }
}

__global__ void __launch_bounds__(kBlock) ground_kernel(float4* __restrict__ pos, uint32_t V,
                                                        const StepConsts* __restrict__ kc, const ColliderSet* __restrict__ cs,
                                                        uint32_t nColliders) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V) return;
  float4 p = pos[i];
  const float4 q = p;
  if (kc->groundEnabled && p.w != 0.0f && p.y < kc->groundY) p.y = kc->groundY;
  if (nColliders) collide_vertex(p, cs, nColliders);
  if (p.x != q.x || p.y != q.y || p.z != q.z) pos[i] = p;
}

__global__ void __launch_bounds__(kBlock) commit_kernel(float4* __restrict__ pos, float4* __restrict__ prev,
                                                        float4* __restrict__ vel, uint32_t V,
                                                        const StepConsts* __restrict__ kc) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V) return;
  const StepConsts k = *kc;
  float4 p = pos[i], x = prev[i], v;
  commit_vertex(p, x, v, k);
  v.w = 0.0f;
  vel[i] = v;
  if (p.w == 0.0f) pos[i] = p; else prev[i] = x;
}

__global__ void __launch_bounds__(kBlock) pack_kernel(const float4* __restrict__ prev, const uint32_t* __restrict__ slotOf,
                                                      float* __restrict__ out, uint32_t V) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V) return;
  const float4 x = prev[slotOf ? slotOf[i] : i];
  out[3 * (size_t)i + 0] = x.x;
  out[3 * (size_t)i + 1] = x.y;
  out[3 * (size_t)i + 2] = x.z;
}

// Per-vertex normals for client-side lighting: K_UpdateNormals of the reference's compute
// solver (Assets/Shaders/SoftBodyCompute.compute:459-491): sum of cross(pb - pa, pc - pa) over the surface
// triangles incident to the vertex, in BuildTriAdjacency's order (SoftBodySolver.cs:1173-1213: ascending
// triangle index), normalised; (0, 1, 0) when the sum vanishes.  Caller vertex order in and out.
__global__ void __launch_bounds__(kBlock) normals_kernel(const float4* __restrict__ prev, const uint32_t* __restrict__ slotOf,
                                                         const uint32_t* __restrict__ tris, const uint32_t* __restrict__ adjOff,
                                                         const uint32_t* __restrict__ adjTri, float* __restrict__ out, uint32_t V) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V) return;
  float sx = 0.0f, sy = 0.0f, sz = 0.0f;
  for (uint32_t k = adjOff[i]; k < adjOff[i + 1]; ++k) {
    const uint32_t t = adjTri[k];
    const uint32_t a = tris[3 * t], b = tris[3 * t + 1], c = tris[3 * t + 2];
    const float4 pa = prev[slotOf ? slotOf[a] : a], pb = prev[slotOf ? slotOf[b] : b], pc = prev[slotOf ? slotOf[c] : c];
    const float ux = fsub(pb.x, pa.x), uy = fsub(pb.y, pa.y), uz = fsub(pb.z, pa.z);
    const float vx = fsub(pc.x, pa.x), vy = fsub(pc.y, pa.y), vz = fsub(pc.z, pa.z);
    sx = fadd(sx, cross_c(uy, vz, uz, vy)); sy = fadd(sy, cross_c(uz, vx, ux, vz)); sz = fadd(sz, cross_c(ux, vy, uy, vx));
  }
  const float n2 = dot3(sx, sy, sz, sx, sy, sz);
  if (n2 < 1e-20f) { sx = 0.0f; sy = 1.0f; sz = 0.0f; }
  else { const float r = fdiv(1.0f, __fsqrt_rn(n2)); sx = fmul(sx, r); sy = fmul(sy, r); sz = fmul(sz, r); }
  out[3 * (size_t)i] = sx; out[3 * (size_t)i + 1] = sy; out[3 * (size_t)i + 2] = sz;
}

// ---- Jacobi + SOR comparison schedule ------------------------------------------------------
// The reference's in-engine solver (Assets/Scripts/Softbody/SoftBodySolver.cs:379-527 == the compute kernels
// K_EdgeGather / K_VolumeGather / K_ApplyDelta, Assets/Shaders/SoftBodyCompute.compute:229-389): every vertex
// GATHERS the corrections of its incident constraints from the same position snapshot, then all vertices apply
// (omega / count) * sum.  Two grid-wide phases per constraint type and iteration, no colouring -- the comparison
// point for the phase-count problem (it converges more slowly than Gauss-Seidel and uses stiffness in [0, 1], not
// XPBD compliance: a different algorithm, NOT a parity target of the PBDServer path).  One thread per vertex,
// CSR adjacency in ascending constraint order (BuildEdgeAdjacency / BuildTetAdjacency, SoftBodySolver.cs:1082-1171).
// float32, reference evaluation order, no FMA.
__global__ void __launch_bounds__(kBlock) jacobi_edge_gather_kernel(const float4* __restrict__ pos, const uint32_t* __restrict__ adjOff,
                                                                    const uint32_t* __restrict__ adjOther, const uint32_t* __restrict__ adjEdge,
                                                                    const float* __restrict__ rest, float4* __restrict__ scratch,
                                                                    uint32_t V, float stiffness) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V) return;
  const float4 xi = pos[i];
  float sx = 0.0f, sy = 0.0f, sz = 0.0f;
  uint32_t cnt = 0;
  if (xi.w != 0.0f) {
    for (uint32_t k = adjOff[i]; k < adjOff[i + 1]; ++k) {
      const float4 xj = pos[adjOther[k]];
      const float w = fadd(xi.w, xj.w);
      if (w == 0.0f) continue;
      const float dx = fsub(xi.x, xj.x), dy = fsub(xi.y, xj.y), dz = fsub(xi.z, xj.z);
      const float len2 = dot3(dx, dy, dz, dx, dy, dz);
      if (len2 < 1e-18f) continue;
      const float len = __fsqrt_rn(len2);
      const float C = fsub(len, rest[adjEdge[k]]);
      const float lambda = fmul(-stiffness, fdiv(C, w));
      const float sc = fmul(lambda, xi.w);
      sx = fadd(sx, fmul(fdiv(dx, len), sc)); sy = fadd(sy, fmul(fdiv(dy, len), sc)); sz = fadd(sz, fmul(fdiv(dz, len), sc));
      ++cnt;
    }
  }
  scratch[i] = make_float4(sx, sy, sz, __uint_as_float(cnt));
}

__global__ void __launch_bounds__(kBlock) jacobi_tet_gather_kernel(const float4* __restrict__ pos, const uint32_t* __restrict__ adjOff,
                                                                   const uint32_t* __restrict__ adjTet, const uint4* __restrict__ tets,
                                                                   const float* __restrict__ rest, float4* __restrict__ scratch,
                                                                   uint32_t V, float stiffness) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V) return;
  float sx = 0.0f, sy = 0.0f, sz = 0.0f;
  uint32_t cnt = 0;
  if (pos[i].w != 0.0f) {
    for (uint32_t kk = adjOff[i]; kk < adjOff[i + 1]; ++kk) {
      const uint32_t t = adjTet[kk] >> 2, role = adjTet[kk] & 3u;
      const uint4 id = tets[t];
      const float4 pa = pos[id.x], pb = pos[id.y], pc = pos[id.z], pd = pos[id.w];
      if (fadd(fadd(fadd(pa.w, pb.w), pc.w), pd.w) == 0.0f) continue;
      const float dbx = fsub(pd.x, pb.x), dby = fsub(pd.y, pb.y), dbz = fsub(pd.z, pb.z);
      const float cbx = fsub(pc.x, pb.x), cby = fsub(pc.y, pb.y), cbz = fsub(pc.z, pb.z);
      const float cax = fsub(pc.x, pa.x), cay = fsub(pc.y, pa.y), caz = fsub(pc.z, pa.z);
      const float dax = fsub(pd.x, pa.x), day = fsub(pd.y, pa.y), daz = fsub(pd.z, pa.z);
      const float bax = fsub(pb.x, pa.x), bay = fsub(pb.y, pa.y), baz = fsub(pb.z, pa.z);
      // gradients DIVIDE by 6 here (SoftBodySolver.cs:483-486), unlike Sim.cpp:146-149
      const float gax = fdiv(cross_c(dby, cbz, dbz, cby), 6.0f), gay = fdiv(cross_c(dbz, cbx, dbx, cbz), 6.0f), gaz = fdiv(cross_c(dbx, cby, dby, cbx), 6.0f);
      const float gbx = fdiv(cross_c(cay, daz, caz, day), 6.0f), gby = fdiv(cross_c(caz, dax, cax, daz), 6.0f), gbz = fdiv(cross_c(cax, day, cay, dax), 6.0f);
      const float gcx = fdiv(cross_c(day, baz, daz, bay), 6.0f), gcy = fdiv(cross_c(daz, bax, dax, baz), 6.0f), gcz = fdiv(cross_c(dax, bay, day, bax), 6.0f);
      const float nx = cross_c(bay, caz, baz, cay), ny = cross_c(baz, cax, bax, caz), nz = cross_c(bax, cay, bay, cax);
      const float gdx = fdiv(nx, 6.0f), gdy = fdiv(ny, 6.0f), gdz = fdiv(nz, 6.0f);
      const float wsum = fadd(fadd(fadd(fmul(pa.w, dot3(gax, gay, gaz, gax, gay, gaz)), fmul(pb.w, dot3(gbx, gby, gbz, gbx, gby, gbz))),
                                   fmul(pc.w, dot3(gcx, gcy, gcz, gcx, gcy, gcz))), fmul(pd.w, dot3(gdx, gdy, gdz, gdx, gdy, gdz)));
      if (wsum < 1e-20f) continue;
      const float vol = fdiv(dot3(nx, ny, nz, dax, day, daz), 6.0f);
      const float lambda = fmul(-stiffness, fdiv(fsub(vol, rest[t]), wsum));
      const float gx = role == 0 ? gax : role == 1 ? gbx : role == 2 ? gcx : gdx;
      const float gy = role == 0 ? gay : role == 1 ? gby : role == 2 ? gcy : gdy;
      const float gz = role == 0 ? gaz : role == 1 ? gbz : role == 2 ? gcz : gdz;
      const float wi = role == 0 ? pa.w : role == 1 ? pb.w : role == 2 ? pc.w : pd.w;
      if (wi == 0.0f) continue;
      const float sc = fmul(lambda, wi);
      sx = fadd(sx, fmul(gx, sc)); sy = fadd(sy, fmul(gy, sc)); sz = fadd(sz, fmul(gz, sc));
      ++cnt;
    }
  }
  scratch[i] = make_float4(sx, sy, sz, __uint_as_float(cnt));
}

// K_ApplyDelta: pos += (omega / count) * sum
__global__ void __launch_bounds__(kBlock) jacobi_apply_kernel(float4* __restrict__ pos, const float4* __restrict__ scratch, uint32_t V, float omega) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V) return;
  const float4 s = scratch[i];
  const uint32_t cnt = __float_as_uint(s.w);
  float4 p = pos[i];
  if (cnt == 0u || p.w == 0.0f) return;
  const float f = fdiv(omega, (float)cnt);
  p.x = fadd(p.x, fmul(f, s.x)); p.y = fadd(p.y, fmul(f, s.y)); p.z = fadd(p.z, fmul(f, s.z));
  pos[i] = p;
}

inline uint32_t blocks_for(uint32_t n) { return (n + kBlock - 1) / kBlock; }

class StreamBackend final : public Backend {
 public:
  StreamBackend(uint32_t flags) : flags_(flags) {}
  ~StreamBackend() override {
    drop_graph();
    cudaFree(edgeIdx_);
    cudaFree(tetIdx_);
    for (auto& e : ev_) if (e) cudaEventDestroy(e);
  }
  const char* name() const override { return "b200-stream"; }

  cudaError_t upload(const Plan& plan, const MeshView& m, DeviceArrays& d) override {
    edgeOff_ = plan.edgeColorOff;
    tetOff_ = plan.tetColorOff;
    std::vector<uint2> e(m.E);
    for (uint32_t k = 0; k < m.E; ++k) {
      const uint32_t c = plan.edgeOrder[k];
      e[k] = make_uint2(plan.vertexToSlot[m.edges[2 * (size_t)c]], plan.vertexToSlot[m.edges[2 * (size_t)c + 1]]);
    }
    std::vector<uint4> t(m.T);
    for (uint32_t k = 0; k < m.T; ++k) {
      const uint32_t* id = m.tets + 4 * (size_t)plan.tetOrder[k];
      t[k] = make_uint4(plan.vertexToSlot[id[0]], plan.vertexToSlot[id[1]], plan.vertexToSlot[id[2]],
                        plan.vertexToSlot[id[3]]);
    }
    cudaError_t err;
    if ((err = cudaMalloc(&edgeIdx_, sizeof(uint2) * (size_t)(m.E + 1))) != cudaSuccess) return err;
    if ((err = cudaMalloc(&tetIdx_, sizeof(uint4) * (size_t)(m.T + 1))) != cudaSuccess) return err;
    bytes_ = sizeof(uint2) * (size_t)m.E + sizeof(uint4) * (size_t)m.T;
    if (m.E && (err = cudaMemcpy(edgeIdx_, e.data(), sizeof(uint2) * (size_t)m.E, cudaMemcpyHostToDevice)) != cudaSuccess) return err;
    if (m.T && (err = cudaMemcpy(tetIdx_, t.data(), sizeof(uint4) * (size_t)m.T, cudaMemcpyHostToDevice)) != cudaSuccess) return err;
    (void)d;
    return cudaSuccess;
  }

  uint32_t launches_per_frame(const FrameShape& f) const override {
    const uint32_t nE = nonempty(edgeOff_), nT = nonempty(tetOff_);
    return f.substeps * (2 + f.iterations * (nE + nT + ((f.groundEnabled || f.nColliders) ? 1 : 0)));
  }
  uint64_t device_bytes() const override { return bytes_; }
  void invalidate() override { drop_graph(); }

  void fill_info(pbd_info& info) const override {
    info.edge_colors = (uint32_t)(edgeOff_.empty() ? 0 : edgeOff_.size() - 1);
    info.tet_colors = (uint32_t)(tetOff_.empty() ? 0 : tetOff_.size() - 1);
    info.edge_phases = info.edge_colors;
    info.tet_phases = info.tet_colors;
    info.block_threads = kBlock;
  }

  cudaError_t enqueue_frame(const DeviceArrays& d, const FrameShape& f, cudaStream_t s) override {
    if (flags_ & PBD_FLAG_STAGE_TIMING) return record(d, f, s, true);
    if (flags_ & PBD_FLAG_NO_GRAPH) return record(d, f, s, false);
    if (!exec_ || !(f.substeps == shape_.substeps && f.iterations == shape_.iterations &&
                    f.groundEnabled == shape_.groundEnabled && f.nColliders == shape_.nColliders)) {
      drop_graph();
      cudaError_t err = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
      if (err != cudaSuccess) return err;
      err = record(d, f, s, false);
      cudaGraph_t g = nullptr;
      cudaError_t err2 = cudaStreamEndCapture(s, &g);
      if (err != cudaSuccess) { if (g) cudaGraphDestroy(g); return err; }
      if (err2 != cudaSuccess) return err2;
      err = cudaGraphInstantiate(&exec_, g, 0);
      cudaGraphDestroy(g);
      if (err != cudaSuccess) { exec_ = nullptr; return err; }
      shape_ = f;
    }
    return cudaGraphLaunch(exec_, s);
  }

  bool stage_ms(double& predict, double& solve, double& commit) override {
    if (!(flags_ & PBD_FLAG_STAGE_TIMING) || marks_.empty()) return false;
    predict = solve = commit = 0.0;
    for (size_t i = 0; i + 1 < marks_.size(); ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev_[i], ev_[i + 1]);
      (marks_[i] == 0 ? predict : marks_[i] == 1 ? solve : commit) += ms;
    }
    return true;
  }

 private:
  static uint32_t nonempty(const std::vector<uint32_t>& off) {
    uint32_t n = 0;
    for (size_t c = 0; c + 1 < off.size(); ++c) n += off[c + 1] > off[c];
    return n;
  }
  void drop_graph() {
    if (exec_) cudaGraphExecDestroy(exec_);
    exec_ = nullptr;
  }
  void mark(int stage, cudaStream_t s) {
    if (ev_.size() <= marks_.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      ev_.push_back(e);
    }
    cudaEventRecord(ev_[marks_.size()], s);
    marks_.push_back(stage);
  }

  // stage codes for timing marks: 0 predict, 1 solve, 2 commit, 3 end
  cudaError_t record(const DeviceArrays& d, const FrameShape& f, cudaStream_t s, bool timed) {
    if (timed) marks_.clear();
    const uint32_t vb = blocks_for(d.V);
    for (uint32_t k = 0; k < f.substeps; ++k) {
      if (timed) mark(0, s);
      if (d.V) predict_kernel<<<vb, kBlock, 0, s>>>(d.pos, d.prev, d.vel, d.V, d.consts);
      if (timed) mark(1, s);
      for (uint32_t it = 0; it < f.iterations; ++it) {
        for (size_t c = 0; c + 1 < edgeOff_.size(); ++c) {
          const uint32_t n = edgeOff_[c + 1] - edgeOff_[c];
          if (n) edge_colour_kernel<<<blocks_for(n), kBlock, 0, s>>>(d.pos, edgeIdx_, d.edgeRest, d.edgeLam, edgeOff_[c], n, d.consts);
        }
        for (size_t c = 0; c + 1 < tetOff_.size(); ++c) {
          const uint32_t n = tetOff_[c + 1] - tetOff_[c];
          if (n) tet_colour_kernel<<<blocks_for(n), kBlock, 0, s>>>(d.pos, tetIdx_, d.tetRest, d.tetLam, tetOff_[c], n, d.consts);
        }
        if ((f.groundEnabled || f.nColliders) && d.V) ground_kernel<<<vb, kBlock, 0, s>>>(d.pos, d.V, d.consts, d.colliders, f.nColliders);
      }
      if (timed) mark(2, s);
      if (d.V) commit_kernel<<<vb, kBlock, 0, s>>>(d.pos, d.prev, d.vel, d.V, d.consts);
    }
    if (timed) mark(3, s);
    return cudaGetLastError();
  }

  uint32_t flags_;
  uint2* edgeIdx_ = nullptr;
  uint4* tetIdx_ = nullptr;
  std::vector<uint32_t> edgeOff_, tetOff_;
  uint64_t bytes_ = 0;
  cudaGraphExec_t exec_ = nullptr;
  FrameShape shape_{};
  std::vector<cudaEvent_t> ev_;
  std::vector<int> marks_;
};


// Jacobi + SOR backend (PBD_BACKEND_JACOBI): S x [ predict ; I x ( edge gather ; apply ; tet gather ; apply ; ground ) ; commit ]
class JacobiBackend final : public Backend {
 public:
  explicit JacobiBackend(const pbd_options& o) : edgeK_(o.jacobi_edge_stiffness > 0.0f ? o.jacobi_edge_stiffness : 0.9f),
                                                 tetK_(o.jacobi_volume_stiffness > 0.0f ? o.jacobi_volume_stiffness : 0.98f) {}
  ~JacobiBackend() override {
    cudaFree(eOff_); cudaFree(eOther_); cudaFree(eEdge_); cudaFree(tOff_); cudaFree(tTet_); cudaFree(tets_); cudaFree(scratch_);
  }
  const char* name() const override { return "b200-jacobi-sor"; }
  void set_omega(float w) override { omega_ = w; }

  cudaError_t upload(const Plan& plan, const MeshView& m, DeviceArrays& d) override {
    (void)plan; (void)d;
    // adjacency in ascending constraint order (SoftBodySolver.cs:1082-1171); identity vertex order
    std::vector<uint32_t> eOff((size_t)m.V + 1, 0), tOff((size_t)m.V + 1, 0);
    for (size_t i = 0; i < (size_t)m.E * 2; ++i) eOff[m.edges[i] + 1]++;
    for (size_t i = 0; i < (size_t)m.T * 4; ++i) tOff[m.tets[i] + 1]++;
    for (uint32_t v = 0; v < m.V; ++v) { eOff[v + 1] += eOff[v]; tOff[v + 1] += tOff[v]; }
    std::vector<uint32_t> eOther((size_t)m.E * 2), eEdge((size_t)m.E * 2), tTet((size_t)m.T * 4);
    {
      std::vector<uint32_t> cur(eOff.begin(), eOff.end() - 1);
      for (uint32_t e = 0; e < m.E; ++e) {
        const uint32_t a = m.edges[2 * (size_t)e], b = m.edges[2 * (size_t)e + 1];
        uint32_t k = cur[a]++; eOther[k] = b; eEdge[k] = e;
        k = cur[b]++; eOther[k] = a; eEdge[k] = e;
      }
      std::vector<uint32_t> cu2(tOff.begin(), tOff.end() - 1);
      for (uint32_t t = 0; t < m.T; ++t)
        for (uint32_t r = 0; r < 4; ++r) tTet[cu2[m.tets[4 * (size_t)t + r]]++] = (t << 2) | r;
    }
    if (m.T >= (1u << 30)) return cudaErrorInvalidValue;
    auto up = [&](uint32_t** dst, const std::vector<uint32_t>& src) -> cudaError_t {
      cudaError_t e = cudaMalloc((void**)dst, sizeof(uint32_t) * (src.size() + 4));
      if (e == cudaSuccess && !src.empty()) e = cudaMemcpy(*dst, src.data(), sizeof(uint32_t) * src.size(), cudaMemcpyHostToDevice);
      bytes_ += sizeof(uint32_t) * src.size();
      return e;
    };
    cudaError_t err;
    if ((err = up(&eOff_, eOff)) != cudaSuccess || (err = up(&eOther_, eOther)) != cudaSuccess || (err = up(&eEdge_, eEdge)) != cudaSuccess ||
        (err = up(&tOff_, tOff)) != cudaSuccess || (err = up(&tTet_, tTet)) != cudaSuccess) return err;
    if ((err = cudaMalloc((void**)&tets_, sizeof(uint4) * ((size_t)m.T + 1))) != cudaSuccess) return err;
    if (m.T && (err = cudaMemcpy(tets_, m.tets, sizeof(uint4) * (size_t)m.T, cudaMemcpyHostToDevice)) != cudaSuccess) return err;
    if ((err = cudaMalloc((void**)&scratch_, sizeof(float4) * ((size_t)m.V + 1))) != cudaSuccess) return err;
    bytes_ += sizeof(uint4) * (size_t)m.T + sizeof(float4) * (size_t)m.V;
    hasE_ = m.E != 0; hasT_ = m.T != 0;
    return cudaSuccess;
  }
  uint32_t launches_per_frame(const FrameShape& f) const override {
    return f.substeps * (2 + f.iterations * ((hasE_ ? 2 : 0) + (hasT_ ? 2 : 0) + ((f.groundEnabled || f.nColliders) ? 1 : 0)));
  }
  uint64_t device_bytes() const override { return bytes_; }
  void fill_info(pbd_info& info) const override { info.block_threads = kBlock; info.edge_phases = hasE_ ? 2 : 0; info.tet_phases = hasT_ ? 2 : 0; }

  cudaError_t enqueue_frame(const DeviceArrays& d, const FrameShape& f, cudaStream_t s) override {
    const uint32_t vb = blocks_for(d.V);
    if (!d.V) return cudaSuccess;
    for (uint32_t k = 0; k < f.substeps; ++k) {
      predict_kernel<<<vb, kBlock, 0, s>>>(d.pos, d.prev, d.vel, d.V, d.consts);
      for (uint32_t it = 0; it < f.iterations; ++it) {
        if (hasE_) {
          jacobi_edge_gather_kernel<<<vb, kBlock, 0, s>>>(d.pos, eOff_, eOther_, eEdge_, d.edgeRest, scratch_, d.V, edgeK_);
          jacobi_apply_kernel<<<vb, kBlock, 0, s>>>(d.pos, scratch_, d.V, omega_);
        }
        if (hasT_) {
          jacobi_tet_gather_kernel<<<vb, kBlock, 0, s>>>(d.pos, tOff_, tTet_, tets_, d.tetRest, scratch_, d.V, tetK_);
          jacobi_apply_kernel<<<vb, kBlock, 0, s>>>(d.pos, scratch_, d.V, omega_);
        }
        if (f.groundEnabled || f.nColliders) ground_kernel<<<vb, kBlock, 0, s>>>(d.pos, d.V, d.consts, d.colliders, f.nColliders);
      }
      commit_kernel<<<vb, kBlock, 0, s>>>(d.pos, d.prev, d.vel, d.V, d.consts);
    }
    return cudaGetLastError();
  }

 private:
  float edgeK_, tetK_, omega_ = 1.4f;
  uint32_t *eOff_ = nullptr, *eOther_ = nullptr, *eEdge_ = nullptr, *tOff_ = nullptr, *tTet_ = nullptr;
  uint4* tets_ = nullptr;
  float4* scratch_ = nullptr;
  bool hasE_ = false, hasT_ = false;
  uint64_t bytes_ = 0;
};

}  // namespace

Backend* make_stream_backend(uint32_t flags, uint32_t) { return new StreamBackend(flags); }
Backend* make_jacobi_backend(const pbd_options& opts) { return new JacobiBackend(opts); }

cudaError_t launch_pack(const DeviceArrays& d, cudaStream_t s) {
  if (d.V) pack_kernel<<<blocks_for(d.V), kBlock, 0, s>>>(d.prev, d.slotOf, d.packed, d.V);
  return cudaGetLastError();
}

cudaError_t launch_normals(const DeviceArrays& d, const uint32_t* tris, const uint32_t* adjOff, const uint32_t* adjTri, cudaStream_t s) {
  if (d.V) normals_kernel<<<blocks_for(d.V), kBlock, 0, s>>>(d.prev, d.slotOf, tris, adjOff, adjTri, d.packed, d.V);
  return cudaGetLastError();
}

StepConsts make_consts(const pbd_params& p, float dt) {
  // host float arithmetic, same expressions as the reference (Sim.cpp:285-286, 101-102, 198-200)
  StepConsts k{};
  const uint32_t ss = p.substeps > 1u ? p.substeps : 1u;
  const float sdt = dt / float(ss);
  const float invDt2 = (sdt > 1e-12f) ? (1.0f / (sdt * sdt)) : 0.0f;
  k.sdt = sdt;
  k.invDt = (sdt > 1e-12f) ? (1.0f / sdt) : 0.0f;
  k.alphaEdge = (p.edgeCompliance > 0.0f ? p.edgeCompliance : 0.0f) * invDt2;
  k.alphaTet = (p.volumeCompliance > 0.0f ? p.volumeCompliance : 0.0f) * invDt2;
  k.gdx = p.gx * sdt; k.gdy = p.gy * sdt; k.gdz = p.gz * sdt;
  k.groundY = p.groundY;
  k.groundYEps = p.groundY + 1e-6f;
  float fr = p.friction < 1.0f ? p.friction : 1.0f;   // fmin(1, friction)
  fr = fr > 0.0f ? fr : 0.0f;                          // fmax(0, .)
  k.fricScale = 1.0f - fr;
  k.groundEnabled = p.groundEnabled ? 1 : 0;
  return k;
}

}  // namespace pbd

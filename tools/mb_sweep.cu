// mb_sweep.cu -- times the tile backend's colour sweeps on ONE real tile (dumped by the library
// with PBD_DUMP_TILE=<file>): cycles per colour step, for the production code and for variants
// that remove one ingredient at a time (to see where a step's latency goes).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I cs121-softbodysim_b200/csrc tools/mb_sweep.cu -o tools/mb_sweep
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "pbd_sweep.cuh"
using namespace pbd;

// Variants of the LANES == 1 tet sweep with one ingredient removed:
//   1: no arithmetic (vertices loaded, stored back unchanged)     2: no vertex LDS/STS (arithmetic on registers)
//   3: no block barrier (racy, timing only)                       4: no record prefetch before the barrier
template <int V>
__device__ __forceinline__ void tet_variant(const TileHdr& h, uint32_t rec, uint32_t svOff, float alpha, long long* ft) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t n = h.nTetGroups;
  const uint2* groups = reinterpret_cast<const uint2*>(smem + rec + h.offTetGroups);
  const uint2* idx = reinterpret_cast<const uint2*>(smem + rec + h.offTetIdx);
  const float* rest = reinterpret_cast<const float*>(smem + rec + h.offTetRest);
  float* lam = reinterpret_cast<float*>(smem + rec + h.offTetLam);
  float4* sv = reinterpret_cast<float4*>(smem + svOff);
  const uint32_t tid = threadIdx.x;
  float4 ka = make_float4(0.1f, 0.2f, 0.3f, 1e8f), kb = make_float4(0.12f, 0.2f, 0.31f, 1e8f),
         kc = make_float4(0.1f, 0.22f, 0.3f, 1e8f), kd = make_float4(0.11f, 0.21f, 0.33f, 1e8f);
  auto project = [&](uint32_t t, uint2 id, float r, float l) {
    const uint32_t a = id.x & 0xffffu, b = id.x >> 16, c = id.y & 0xffffu, d = id.y >> 16;
    float4 pa, pb, pc, pd;
    if (V == 2) { pa = ka; pb = kb; pc = kc; pd = kd; pa.x += r; } else { pa = sv[a]; pb = sv[b]; pc = sv[c]; pd = sv[d]; }
    float nl = l;
    bool ok = true;
    if (V != 1) ok = tet_delta(pa, pb, pc, pd, r, l, alpha, nl);
    if (ok) {
      if (V == 2) { ka = pa; kb = pb; kc = pc; kd = pd; } else { sv[a] = pa; sv[b] = pb; sv[c] = pc; sv[d] = pd; }
      lam[t] = nl;
    }
  };
  uint2 gd = groups[0];
  uint32_t t = gd.x + tid;
  uint2 id = make_uint2(0u, 0u);
  float r = 0.0f, l = 0.0f;
  bool have = tid < gd.y;
  if (have) { id = idx[t]; r = rest[t]; l = lam[t]; }
  for (uint32_t g = 0; g < n; ++g) {
    const uint2 gn = (g + 1 < n) ? groups[g + 1] : make_uint2(0u, 0u);
    if (V == 4 && have) { id = idx[t]; r = rest[t]; l = lam[t]; }
    if (have) project(t, id, r, l);
    t = gn.x + tid;
    have = tid < gn.y;
    if (V != 4 && have) { id = idx[t]; r = rest[t]; l = lam[t]; }
    if (V != 3) __syncthreads();
    if (ft && g < 40) { ft[16 + g] = clock64(); ft[56 + g] = gd.y; }
    gd = gn;
  }
  if (V == 2 && ka.x == 123.f) lam[0] = kb.y + kc.z + kd.x;
}

template <int LANES, int VARIANT>
__global__ void __launch_bounds__(512, 1) mb(const unsigned char* recG, uint32_t recBytes, const float4* vertsG,
                                             uint32_t nVerts, long long* out, int iters, float alphaE, float alphaT) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t svOff = (recBytes + 127u) & ~127u;
  for (uint32_t i = threadIdx.x; i < recBytes / 4; i += blockDim.x) ((uint32_t*)smem)[i] = ((const uint32_t*)recG)[i];
  float4* sv = (float4*)(smem + svOff);
  for (uint32_t i = threadIdx.x; i < nVerts; i += blockDim.x) sv[i] = vertsG[i];
  __syncthreads();
  const TileHdr h = *(const TileHdr*)smem;
  __shared__ long long ft[256];
  long long acc[2] = {0, 0};
  for (int it = 0; it < iters; ++it) {
    long long t0 = clock64();
    if (VARIANT == 0) sweep_edges(h, 0, svOff, alphaE, (threadIdx.x == 0 && it == iters - 1) ? ft : nullptr);
    long long t1 = clock64();
    if (VARIANT == 0) sweep_tets<LANES>(h, 0, svOff, alphaT, (threadIdx.x == 0 && it == iters - 1) ? ft + 20 : nullptr);
    else tet_variant<VARIANT>(h, 0, svOff, alphaT, (threadIdx.x == 0 && it == iters - 1) ? ft + 20 : nullptr);
    long long t2 = clock64();
    acc[0] += t1 - t0; acc[1] += t2 - t1;
    if (threadIdx.x == 0 && it == iters - 1) { ft[0] = t0; ft[1] = t1; }
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    out[0] = acc[0] / iters; out[1] = acc[1] / iters; out[2] = h.nEdgeGroups; out[3] = h.nTetGroups;
    long long prev = ft[0];
    for (uint32_t g = 0; g < h.nEdgeGroups && g < 20; ++g) { out[16 + g] = ft[16 + g] - prev; out[56 + g] = ft[56 + g]; prev = ft[16 + g]; }
    prev = ft[1];
    for (uint32_t g = 0; g < h.nTetGroups && g < 20; ++g) { out[36 + g] = ft[36 + g] - prev; out[76 + g] = ft[76 + g]; prev = ft[36 + g]; }
  }
}

template <int LANES, int VARIANT = 0>
void run(const char* name, const unsigned char* dRec, uint32_t recBytes, const float4* dV, uint32_t nV, int grid, int iters = 50) {
  long long* o; cudaMalloc(&o, 128 * 8); cudaMemset(o, 0, 128 * 8);
  const size_t smem = ((recBytes + 127u) & ~127u) + 16 * (size_t)nV;
  cudaFuncSetAttribute(mb<LANES, VARIANT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  mb<LANES, VARIANT><<<grid, 512, smem>>>(dRec, recBytes, dV, nV, o, iters, 720.0f, 0.0f);
  cudaError_t e = cudaDeviceSynchronize();
  long long r[128]; cudaMemcpy(r, o, sizeof(r), cudaMemcpyDeviceToHost);
  printf("%s (grid %d): %s | edge sweep %lld cyc (%lld groups), tet sweep %lld cyc (%lld groups)\n", name, grid, cudaGetErrorString(e), r[0], r[2], r[1], r[3]);
  printf("  edge steps:"); for (int g = 0; g < r[2] && g < 20; ++g) printf(" %lld/%lld", r[16 + g], r[56 + g]);
  printf("\n  tet steps:"); for (int g = 0; g < r[3] && g < 20; ++g) printf(" %lld/%lld", r[36 + g], r[76 + g]);
  printf("\n");
  cudaFree(o);
}

int main(int argc, char** argv) {
  const char* path = argc > 1 ? argv[1] : "gpurun_out/tile.bin";
  FILE* f = fopen(path, "rb");
  if (!f) { printf("cannot open %s\n", path); return 1; }
  uint32_t head[6];
  if (fread(head, sizeof(head), 1, f) != 1 || head[0] != 0x54494c45u) { printf("bad dump\n"); return 1; }
  std::vector<unsigned char> rec(head[1]);
  std::vector<float4> v(head[3]);
  if (fread(rec.data(), 1, rec.size(), f) != rec.size() || fread(v.data(), 16, v.size(), f) != v.size()) { printf("short dump\n"); return 1; }
  fclose(f);
  printf("tile: %u verts, %u edges, %u tets, record %u bytes\n", head[3], head[4], head[5], head[1]);
  unsigned char* dRec; float4* dV;
  cudaMalloc(&dRec, rec.size()); cudaMalloc(&dV, 16 * v.size());
  cudaMemcpy(dRec, rec.data(), rec.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dV, v.data(), 16 * v.size(), cudaMemcpyHostToDevice);
  run<1>("lanes 1", dRec, head[1], dV, head[3], 1);
  run<2>("lanes 2", dRec, head[1], dV, head[3], 1);
  run<4>("lanes 4", dRec, head[1], dV, head[3], 1);
  run<1>("lanes 1, 2 iterations only", dRec, head[1], dV, head[3], 1, 2);
  run<1, 1>("V1 no arithmetic", dRec, head[1], dV, head[3], 1);
  run<1, 2>("V2 no vertex LDS/STS", dRec, head[1], dV, head[3], 1);
  run<1, 3>("V3 no barrier", dRec, head[1], dV, head[3], 1);
  run<1, 4>("V4 no prefetch", dRec, head[1], dV, head[3], 1);
  return 0;
}

// microbench_step.cu -- what does one colour step of the tile sweep cost on B200?
// Single CTA, 512 threads, shared-memory vertex tile; measures cycles per step for:
//   A  __syncthreads() only
//   B  4x LDS.128 + 4x STS.128 + barrier (no math)
//   C  project_tet on shared data + barrier, with 1 / 6 / 16 active warps
//   D  project_edge likewise
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -I cs121-softbodysim_b200/csrc tools/microbench_step.cu -o /tmp/mb
#include <cstdio>
#include <cstdint>
#include "pbd_math.cuh"
using namespace pbd;

__global__ void __launch_bounds__(512, 1) mb(float4* gpos, long long* out, int iters) {
  __shared__ float4 sv[2048];
  const int tid = threadIdx.x;
  for (int i = tid; i < 2048; i += blockDim.x) sv[i] = gpos[i];
  __syncthreads();
  long long t0, t1;
  int slot = 0;
  // A
  t0 = clock64();
  for (int i = 0; i < iters; ++i) __syncthreads();
  t1 = clock64();
  if (tid == 0) out[slot] = (t1 - t0) / iters;
  slot++;
  // B
  {
    const int a = (tid * 4) & 2047, b = (tid * 4 + 1) & 2047, c = (tid * 4 + 2) & 2047, d = (tid * 4 + 3) & 2047;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      float4 pa = sv[a], pb = sv[b], pc = sv[c], pd = sv[d];
      pa.x += 1.f; pb.x += 1.f; pc.x += 1.f; pd.x += 1.f;
      sv[a] = pa; sv[b] = pb; sv[c] = pc; sv[d] = pd;
      __syncthreads();
    }
    t1 = clock64();
    if (tid == 0) out[slot] = (t1 - t0) / iters;
    slot++;
  }
  // C: tets with nActive warps
  for (int nw : {1, 2, 6, 16}) {
    const int a = (tid * 4) & 2047, b = (tid * 4 + 1) & 2047, c = (tid * 4 + 2) & 2047, d = (tid * 4 + 3) & 2047;
    float lam = 0.f;
    __syncthreads();
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (tid < nw * 32) {
        float4 pa = sv[a], pb = sv[b], pc = sv[c], pd = sv[d];
        if (project_tet(pa, pb, pc, pd, 1.0e-6f, lam, 0.0f)) { sv[a] = pa; sv[b] = pb; sv[c] = pc; sv[d] = pd; }
      }
      __syncthreads();
    }
    t1 = clock64();
    if (tid == 0) out[slot] = (t1 - t0) / iters;
    slot++;
    if (lam == 123.f) out[63] = 1;
  }
  // D: edges
  for (int nw : {1, 2, 6, 16}) {
    const int a = (tid * 2) & 2047, b = (tid * 2 + 1) & 2047;
    float lam = 0.f;
    __syncthreads();
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (tid < nw * 32) {
        float4 p0 = sv[a], p1 = sv[b];
        if (project_edge(p0, p1, 0.02f, lam, 720.0f)) { sv[a] = p0; sv[b] = p1; }
      }
      __syncthreads();
    }
    t1 = clock64();
    if (tid == 0) out[slot] = (t1 - t0) / iters;
    slot++;
    if (lam == 123.f) out[63] = 1;
  }
  for (int i = tid; i < 2048; i += blockDim.x) gpos[i] = sv[i];
}

int main() {
  float4* h = new float4[2048];
  for (int i = 0; i < 2048; ++i) {
    // a jittered lattice so tets (4 consecutive vertices) are non-degenerate
    h[i] = make_float4(0.02f * (i % 13) + 0.001f * (i % 7), 0.3f + 0.02f * ((i / 13) % 11) + 0.0013f * (i % 5),
                       0.02f * (i / 143) + 0.0017f * (i % 3), 1.0e8f);
  }
  float4* d; long long* o;
  cudaMalloc(&d, sizeof(float4) * 2048); cudaMalloc(&o, 64 * 8);
  cudaMemset(o, 0, 64 * 8);
  cudaMemcpy(d, h, sizeof(float4) * 2048, cudaMemcpyHostToDevice);
  mb<<<1, 512>>>(d, o, 2000);
  cudaError_t e = cudaDeviceSynchronize();
  long long r[64];
  cudaMemcpy(r, o, sizeof(r), cudaMemcpyDeviceToHost);
  printf("status %s\n", cudaGetErrorString(e));
  const char* names[] = {"A barrier only", "B 4xLDS+4xSTS+bar", "C tet 1 warp", "C tet 2 warps", "C tet 6 warps", "C tet 16 warps",
                         "D edge 1 warp", "D edge 2 warps", "D edge 6 warps", "D edge 16 warps"};
  for (int i = 0; i < 10; ++i) printf("%-22s %lld cycles/step\n", names[i], r[i]);
  return 0;
}

// Does instruction-level parallelism inside a thread buy what more CTAs per SM cannot?  ILP independent copies
// of the per-level chain (3 LDS, max, 3 ex2, lg2, STS) per thread, ONE barrier per level for all of them.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o chain2 chain2.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lds(unsigned a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts(unsigned a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
template <int ILP, int NT>
__global__ void __launch_bounds__(NT) chain(float* out, int T, const float* em) {
  __shared__ float rows[ILP][2][264];
  const int tid = threadIdx.x;
  for (int k = 0; k < ILP; k++) {
    rows[k][0][tid] = 0.001f * tid + k; rows[k][1][tid] = 0.0f;
    if (tid < 8) { rows[k][0][256 + tid] = -1e30f; rows[k][1][256 + tid] = -1e30f; }
  }
  __syncthreads();
  unsigned pc[ILP], qc[ILP];
  for (int k = 0; k < ILP; k++) {
    pc[k] = (unsigned)__cvta_generic_to_shared(&rows[k][0][0]);
    qc[k] = (unsigned)__cvta_generic_to_shared(&rows[k][1][0]);
  }
  const unsigned s0 = 4u * tid, s1 = 4u * (tid > 0 ? tid - 1 : 256), s2 = 4u * (tid > 1 ? tid - 2 : 256);
  float e = em[tid & 63];
  float acc = 0.f;
  for (int t = 0; t < T; t++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) {
      float a = lds(pc[k] + s0), b = lds(pc[k] + s1), c = lds(pc[k] + s2);
      float m = fmaxf(fmaxf(a, b), c);
      float s = ex2(a - m) + ex2(b - m) + ex2(c - m);
      float v = m + lg2(s) + e;
      sts(qc[k] + s0, v);
      acc += v;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
#pragma unroll
    for (int k = 0; k < ILP; k++) { unsigned tmp = pc[k]; pc[k] = qc[k]; qc[k] = tmp; }
  }
  out[blockIdx.x * NT + tid] = acc;
}
template <int ILP, int NT>
void run(const char* name, float* out, const float* em) {
  const int T = 4000;
  for (int per_sm : {1, 2, 3, 4}) {
    int grid = 148 * per_sm;
    chain<ILP, NT><<<grid, NT>>>(out, 100, em);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    chain<ILP, NT><<<grid, NT>>>(out, T, em);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double cyc = ms * 1e-3 / T * clk * 1e3;
    printf("%-22s CTAs/SM %d  %.0f cycles/level  per chain-level per SM: %.0f cycles (chains/SM %d)\n", name, per_sm, cyc,
           cyc / (per_sm * ILP), per_sm * ILP);
  }
}
int main() {
  float *out, *em; cudaMalloc(&out, 148 * 8 * 256 * 4); cudaMalloc(&em, 256); cudaMemset(em, 0, 256);
  run<1, 224>("ILP 1, 7 warps", out, em);
  run<2, 224>("ILP 2, 7 warps", out, em);
  run<4, 224>("ILP 4, 7 warps", out, em);
  run<2, 128>("ILP 2, 4 warps", out, em);
  run<4, 128>("ILP 4, 4 warps", out, em);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

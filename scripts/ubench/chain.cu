// Microbenchmark of the per-level dependent chain of the CTC sweeps (k_bidir.cu node warps): what bounds
// a level when several CTAs share an SM?  Variants switch off pieces of the level (the results are then
// meaningless; only the time matters).  Build: nvcc -arch=sm_100a -O3 -o chain chain.cu ; run: ./chain
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lds(unsigned a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts(unsigned a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
// NEX: number of ex2 on the chain (0..3); LG: lg2 on the chain; BAR: 0 none, 1 bar.sync, 2 __syncwarp only
template <int NEX, int LG, int BAR, int NLDS>
__global__ void __launch_bounds__(224) chain(float* out, int T, const float* em) {
  __shared__ float rows[2][264];
  const int tid = threadIdx.x;
  rows[0][tid] = 0.001f * tid; rows[1][tid] = 0.0f;
  if (tid < 8) { rows[0][256 + tid] = -1e30f; rows[1][256 + tid] = -1e30f; }
  __syncthreads();
  unsigned pc = (unsigned)__cvta_generic_to_shared(&rows[0][0]), qc = (unsigned)__cvta_generic_to_shared(&rows[1][0]);
  const unsigned s0 = 4u * tid, s1 = 4u * (tid > 0 ? tid - 1 : 256), s2 = 4u * (tid > 1 ? tid - 2 : 256);
  float e = em[tid & 63];
  float acc = 0.f;
  for (int t = 0; t < T; t++) {
    float a = lds(pc + s0), b = NLDS > 1 ? lds(pc + s1) : a - 1.0f, c = NLDS > 2 ? lds(pc + s2) : a - 2.0f;
    float m = fmaxf(fmaxf(a, b), c);
    float s = 1.0f;
    if (NEX > 0) s = ex2(a - m); else s = (a - m) + 1.0f;
    if (NEX > 1) s += ex2(b - m); else s += (b - m) * 0.001f;
    if (NEX > 2) s += ex2(c - m); else s += (c - m) * 0.001f;
    float v = m + (LG ? lg2(s) : s * 0.5f) + e;
    sts(qc + s0, v);
    acc += v;
    if (BAR == 1) asm volatile("bar.sync 1, 224;" ::: "memory");
    else if (BAR == 2) __syncwarp();
    unsigned tmp = pc; pc = qc; qc = tmp;
  }
  out[blockIdx.x * 224 + tid] = acc;
}
template <int NEX, int LG, int BAR, int NLDS>
void run(const char* name, float* out, const float* em) {
  const int T = 4000;
  for (int per_sm : {1, 2, 4, 6}) {
    int grid = 148 * per_sm;
    chain<NEX, LG, BAR, NLDS><<<grid, 224>>>(out, 100, em);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    chain<NEX, LG, BAR, NLDS><<<grid, 224>>>(out, T, em);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-28s CTAs/SM %d  %.1f ns/level  (%.0f cycles at %d MHz)  per CTA-level per SM: %.0f cycles\n", name, per_sm,
           ms * 1e6 / T, ms * 1e-3 / T * clk * 1e3, clk / 1000, ms * 1e-3 / T * clk * 1e3 / per_sm);
  }
}
int main() {
  float *out, *em; cudaMalloc(&out, 148 * 8 * 224 * 4); cudaMalloc(&em, 256); cudaMemset(em, 0, 256);
  run<3, 1, 1, 3>("full (3 ex2, lg2, bar, 3 lds)", out, em);
  run<3, 1, 0, 3>("no barrier", out, em);
  run<3, 1, 2, 3>("syncwarp only", out, em);
  run<0, 0, 1, 3>("no MUFU", out, em);
  run<1, 1, 1, 3>("1 ex2 + lg2", out, em);
  run<3, 1, 1, 1>("1 lds", out, em);
  run<0, 0, 0, 1>("bare", out, em);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

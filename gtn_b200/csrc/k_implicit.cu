/*
 * k_implicit.cu -- forwardScore(intersect(g, emissions)) and its backward WITHOUT
 * materialising the lattice ("implicit lattice": SURVEY.md section 8(d), B_io formulation).
 *
 * compose(g, linearGraph(T, C)) is frame synchronous: composed node (u, t) exists iff graph
 * node u can be reached from a start node in exactly t arcs and can reach an accept node in
 * exactly T - t arcs (compose.cpp:64-104 findReachable + the forward BFS :400-470); its
 * in-arcs are the graph's in-arcs (i -> u, label c, weight w) whose source (i, t-1) exists,
 * with weight fl(w + emissions[t-1][c]) (compose.cpp:435).  So instead of building the
 * lattice and sweeping it, one CTA per utterance keeps the graph operand in shared memory /
 * registers and sweeps the frames directly:
 *
 *   forward  S_t[u] = logadd_{(i->u)} ( S_{t-1}[i] + fl(w + e[t-1][c]) ),  S_0 = 0 on start
 *            nodes, -inf elsewhere (shortest.cpp:102-136).  Nodes the reference prunes are
 *            either unreachable (their score is -inf and contributes exp(-inf) = 0 to every
 *            sum, exactly like an absent arc) or cannot reach an accept node (they never feed
 *            a surviving node), so the surviving nodes get the reference's scores.
 *   backward shortestDistanceGrad (shortest.cpp:33-82) + compose's gradFunc
 *            (compose.cpp:496-518) on the same implicit structure: a node with score -inf
 *            does not exist (no gradient), a node that cannot reach an accept node receives
 *            exactly 0.  Per level the only value on the serial chain is the node gradient:
 *            the arc factors exp(S_{t-1}[i] + w' - S_t[u]) do not depend on it and are
 *            computed one level ahead, in the shadow of the barrier.
 *
 * HBM traffic per utterance: emissions read once per pass (4TC each), dense per-frame node
 * scores written by forward and read by backward (4(T+1)pitch each), emission gradients
 * accumulated with one red.global per (node, frame).  No row_ptr / arc / gradInfo arrays.
 *
 * Valid only when every weight is finite (then "score == -inf" <=> "node does not exist");
 * both kernels raise a per-utterance status bit on a non-finite emission or arc weight and
 * the caller re-runs the batch through the materialised path, which reproduces the
 * reference's NaN / inf propagation arc by arc.  Requires: no epsilon labels, all labels in
 * [0, C), all in-arcs of a graph node carry the same label (CTC, forced alignment, ASG).
 */
#include <cuda_runtime.h>
#include <math_constants.h>

#include <algorithm>

#include "gtnb_internal.h"

namespace gtnb {

namespace {

constexpr int kImpThreads = 256;
constexpr int kPf = 8; // frames of register prefetch (emissions, saved scores)
// Fast path (graph operand with <= kImpThreads nodes, in/out degree <= 3): one node per thread,
// branch-free.  Score rows and arc-gradient buffers have compile-time strides so that the two
// halves of each double buffer differ by an immediate offset; absent arcs point at a dummy
// slot (score -inf / gradient 0) instead of being predicated off.
constexpr int kFastRow = kImpThreads + 8; // floats per score row
constexpr int kRowDummy = kImpThreads + 4; // holds -inf in both rows
constexpr int kFastCur = 3 * kImpThreads + 8; // floats per arc-gradient buffer
constexpr int kCurDummy = 3 * kImpThreads + 4; // holds 0 in both buffers

struct ImpLayout {
  int pitch; // floats per score row in shared memory (max over the batch, multiple of 4)
  int tabN, tabA; // table capacities
  int off_S, off_inptr, off_src, off_w, off_lab, off_flags, off_red;
  int off_outptr, off_outent, off_cursor, off_cur, off_ng; // backward only
  int total;
};

ImpLayout make_imp_layout(int maxN, int maxA, bool backward) {
  ImpLayout o;
  o.pitch = (maxN + 3) & ~3;
  o.tabN = maxN + 1;
  o.tabA = std::max(maxA, 1);
  int off = 0;
  auto take = [&](int bytes) {
    int at = off;
    off += (bytes + 15) & ~15;
    return at;
  };
  o.off_S = take(4 * 2 * std::max(o.pitch, kFastRow));
  o.off_inptr = take(4 * (o.tabN + 1));
  o.off_src = take(4 * o.tabA);
  o.off_w = take(4 * o.tabA);
  o.off_lab = take(4 * o.tabN);
  o.off_flags = take(o.tabN);
  o.off_red = take(4 * 16);
  o.off_outptr = o.off_outent = o.off_cursor = o.off_cur = o.off_ng = 0;
  if (backward) {
    o.off_outptr = take(4 * (o.tabN + 1));
    o.off_outent = take(4 * o.tabA);
    o.off_cursor = take(4 * o.tabN);
    o.off_cur = take(4 * 2 * std::max(o.tabA, kFastCur));
    o.off_ng = take(4 * o.pitch);
  }
  o.total = off;
  return o;
}

__device__ __forceinline__ float ninf() {
  return -CUDART_INF_F;
}
__device__ __forceinline__ bool finite_f(float x) {
  return fabsf(x) < CUDART_INF_F;
}
/* exp / log1p on the SFU, as in k_staged.cu */
__device__ __forceinline__ float iexp(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
  return y;
}
__device__ __forceinline__ float ilog1p(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(1.0f + x));
  return (x < 1e-3f) ? x * (1.0f - 0.5f * x) : y * 0.6931471805599453f;
}

__device__ __forceinline__ float block_max(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < kImpThreads / 32; w++) r = fmaxf(r, red[w]);
  return r;
}
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < kImpThreads / 32; w++) r += red[w];
  return r;
}

/* the graph operand's in-arc tables -> shared memory; returns "some weight is not finite" */
__device__ __forceinline__ bool load_tables(
    const GraphMeta& m, const uint8_t* __restrict__ sg_flags, const int32_t* __restrict__ sg_in_ptr,
    const int32_t* __restrict__ sg_in_src, const int32_t* __restrict__ sg_in_label,
    const float* __restrict__ sg_in_w, int* t_inptr, int* t_src, float* t_w, int* t_lab,
    unsigned char* t_flags) {
  const int tid = threadIdx.x, N1 = m.sg_N, A1 = m.sg_A;
  const int32_t* ip = sg_in_ptr + m.sg_node_base;
  const int32_t* is = sg_in_src + m.sg_arc_base;
  const int32_t* il = sg_in_label + m.sg_arc_base;
  const float* iw = sg_in_w + m.sg_arc_base;
  const uint8_t* fl = sg_flags + m.sg_node_base;
  bool bad = false;
  for (int i = tid; i <= N1; i += kImpThreads) t_inptr[i] = ip[i];
  for (int e = tid; e < A1; e += kImpThreads) {
    t_src[e] = is[e];
    const float w = iw[e];
    t_w[e] = w;
    bad |= !finite_f(w);
  }
  __syncthreads();
  for (int i = tid; i < N1; i += kImpThreads) {
    t_lab[i] = (t_inptr[i] < t_inptr[i + 1]) ? il[t_inptr[i]] : -1;
    t_flags[i] = fl[i];
  }
  __syncthreads();
  return bad;
}

/* Predicated read-only load straight into its destination register.  With a plain
 * `if (p) x = __ldg(q)` in the unrolled frame loop ptxas gathers the eight prefetches of a
 * block at its end, into temporaries, and copies them to the loop-carried registers at the
 * back edge -- a full-latency stall every kPf frames (ncu: 21 % of all samples on that MOV). */
__device__ __forceinline__ void ldg_if(bool p, float& dst, const float* src) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "setp.ne.s32 q, %2, 0;\n"
      "@q ld.global.nc.f32 %0, [%1];\n"
      "}\n"
      : "+f"(dst)
      : "l"(src), "r"((int)p));
}
/* (Measured alternatives, both slower at B=256: a 2-deep register ring behind prefetch.global.L1
 * issued 8 frames ahead -- 0.33 vs 0.27 ms for the backward; prefetching the three source scores
 * of every arc into registers instead of staging the row in shared memory -- 0.36 ms.) */
/* Unpredicated variant for the steady part of the frame loop (address always valid): a strong
 * (relaxed, CTA scope) load, which ptxas must keep on its side of the bar.sync it was written
 * on, writing the loop-carried ring register directly. */
__device__ __forceinline__ void ldg_keep(float& dst, const float* src) {
  asm volatile("ld.relaxed.cta.global.f32 %0, [%1];" : "=f"(dst) : "l"(src) : "memory");
}
/* predicated red.global.add.f32 (no branch, no reconvergence point in the frame loop) */
__device__ __forceinline__ void red_if(bool p, float* dst, float v) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "setp.ne.s32 q, %2, 0;\n"
      "@q red.global.add.f32 [%0], %1;\n"
      "}\n" ::"l"(dst),
      "f"(v), "r"((int)p)
      : "memory");
}

/* three in-arcs, branch-free: absent arcs carry -inf and add exp(-inf) = 0 in arc order
 * (shortest.cpp:102-114); an all--inf (or +inf) maximum is returned as is.  The sum is >= 1
 * (the maximum contributes exp(0)), where lg2.approx has an absolute error <= 2^-22: far
 * below one ulp of any score this recursion produces after a few frames. */
__device__ __forceinline__ float lse3(float v0, float v1, float v2) {
  const float mx = fmaxf(fmaxf(v0, v1), v2);
  const float sum = (iexp(v0 - mx) + iexp(v1 - mx)) + iexp(v2 - mx);
  float lg;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(sum));
  const float r = fmaf(lg, 0.6931471805599453f, mx);
  return (fabsf(mx) == CUDART_INF_F) ? mx : r;
}

/* one node, any in-degree: the reference's two passes (shortest.cpp:121-136, :102-114) */
__device__ __noinline__ float imp_fwd_node_any(
    int e0, int e1, float e, const int* t_src, const float* t_w, const float* prev) {
  float mx = ninf();
#pragma unroll 1
  for (int a = e0; a < e1; a++) mx = fmaxf(mx, prev[t_src[a]] + (t_w[a] + e));
  if (mx == CUDART_INF_F || mx == -CUDART_INF_F) return mx;
  float sum = -1.0f;
#pragma unroll 1
  for (int a = e0; a < e1; a++) sum += iexp((prev[t_src[a]] + (t_w[a] + e)) - mx);
  return mx + ilog1p(sum);
}

/* ------------------------------------------------------------------ */
/* forward                                                             */
/* ------------------------------------------------------------------ */

__global__ void __launch_bounds__(kImpThreads) implicit_forward_kernel(
    const GraphMeta* __restrict__ meta,
    const uint8_t* __restrict__ sg_flags,
    const int32_t* __restrict__ sg_in_ptr,
    const int32_t* __restrict__ sg_in_src,
    const int32_t* __restrict__ sg_in_label,
    const float* __restrict__ sg_in_w,
    const float* __restrict__ emissions,
    float* __restrict__ scores,
    float* __restrict__ out_scores,
    int32_t* __restrict__ status,
    int C,
    const ImpLayout lay) {
  extern __shared__ __align__(16) unsigned char smem[];
  const GraphMeta m = meta[blockIdx.x];
  const int tid = threadIdx.x;
  const int N1 = m.sg_N, T = m.T;
  const int pitch = (N1 + 3) & ~3;
  float* S = reinterpret_cast<float*>(smem + lay.off_S);
  int* t_inptr = reinterpret_cast<int*>(smem + lay.off_inptr);
  int* t_src = reinterpret_cast<int*>(smem + lay.off_src);
  float* t_w = reinterpret_cast<float*>(smem + lay.off_w);
  int* t_lab = reinterpret_cast<int*>(smem + lay.off_lab);
  unsigned char* t_flags = smem + lay.off_flags;
  float* red = reinterpret_cast<float*>(smem + lay.off_red);

  bool bad = load_tables(m, sg_flags, sg_in_ptr, sg_in_src, sg_in_label, sg_in_w, t_inptr, t_src, t_w,
                         t_lab, t_flags);
  float* sc = scores + m.node_base;
  const float* em = emissions + m.emis_off;
  // fast path: one node per thread, every node with at most three in-arcs (block-uniform)
  int my_deg = 0;
  if (tid < N1) my_deg = t_inptr[tid + 1] - t_inptr[tid];
  const bool fast = (N1 <= kImpThreads) && !__syncthreads_or(my_deg > 3);
  float* S0 = S;
  float* S1 = S + (fast ? kFastRow : lay.pitch);
  // frame 0: start nodes carry the implicit 0 (shortest.cpp:129-135), nothing else exists
  if (fast) {
    const float v = (tid < N1 && (t_flags[tid] & 1)) ? 0.0f : ninf();
    S0[tid] = v;
    if (tid < N1) sc[tid] = v;
    if (tid == 0) S0[kRowDummy] = S1[kRowDummy] = ninf();
  } else {
    for (int i = tid; i < N1; i += kImpThreads) {
      const float v = (t_flags[i] & 1) ? 0.0f : ninf();
      S0[i] = v;
      sc[i] = v;
    }
  }
  __syncthreads();

  if (fast) {
    const bool act = tid < N1;
    const int deg = my_deg;
    const bool has = deg > 0;
    int i0 = kRowDummy, i1 = kRowDummy, i2 = kRowDummy;
    float w0 = 0.0f, w1 = 0.0f, w2 = 0.0f;
    if (has) {
      const int e0 = t_inptr[tid];
      i0 = t_src[e0];
      w0 = t_w[e0];
      if (deg > 1) {
        i1 = t_src[e0 + 1];
        w1 = t_w[e0 + 1];
      }
      if (deg > 2) {
        i2 = t_src[e0 + 2];
        w2 = t_w[e0 + 2];
      }
    }
    const float* eml = em + (has ? t_lab[tid] : 0);
    float pf[kPf]; // emissions of the next kPf frames (the node's label), loaded kPf frames ahead
#pragma unroll
    for (int j = 0; j < kPf; j++) pf[j] = (has && j < T) ? __ldg(eml + (long long)j * C) : 0.0f;
    const float* epf = eml + (long long)kPf * C;
    float* grow = sc + pitch + tid; // this node's score in frame f + 1
    float chk = 0.0f; // becomes NaN when a used emission is not finite
    int f0 = 0;
#define GTNB_FWD_STEP(j, GUARDED)                                        \
  {                                                                      \
    const float e = has ? pf[j] : 0.0f; /* idle threads load node 0's */ \
    if (GUARDED)                                                         \
      ldg_if(has && f0 + (j) + kPf < T, pf[j], epf);                     \
    else                                                                 \
      ldg_keep(pf[j], epf);                                              \
    epf += C;                                                            \
    const float* P = ((j)&1) ? S1 : S0;                                  \
    float* Q = ((j)&1) ? S0 : S1;                                        \
    const float sv = lse3(P[i0] + (w0 + e), P[i1] + (w1 + e), P[i2] + (w2 + e)); \
    chk = fmaf(e, 0.0f, chk);                                            \
    Q[tid] = sv;                                                         \
    if (act) *grow = sv;                                                 \
    grow += pitch;                                                       \
    __syncthreads();                                                     \
  }
    for (; f0 + 2 * kPf <= T; f0 += kPf) {
#pragma unroll
      for (int j = 0; j < kPf; j++) GTNB_FWD_STEP(j, false)
    }
    for (; f0 < T; f0 += kPf) {
#pragma unroll
      for (int j = 0; j < kPf; j++) {
        if (f0 + j >= T) break;
        GTNB_FWD_STEP(j, true)
      }
    }
#undef GTNB_FWD_STEP
    bad |= !(chk == 0.0f);
  } else {
    // ---- any graph: several nodes per thread, tables in shared memory
    float* prev = S0;
    float* cur = S1;
    for (int t = 1; t <= T; t++) {
      const float* er = em + (long long)(t - 1) * C;
      float* srow = sc + (long long)t * pitch;
      for (int i = tid; i < N1; i += kImpThreads) {
        const int e0 = t_inptr[i], e1 = t_inptr[i + 1];
        float sv = ninf();
        if (e1 > e0) {
          const float e = __ldg(er + t_lab[i]);
          bad |= !finite_f(e);
          sv = imp_fwd_node_any(e0, e1, e, t_src, t_w, prev);
        }
        cur[i] = sv;
        srow[i] = sv;
      }
      float* tmp = prev;
      prev = cur;
      cur = tmp;
      __syncthreads();
    }
  }

  // accept nodes of the last frame (shortest.cpp:147-159)
  const float* Sf = (T & 1) ? S1 : S0;
  float mx = ninf();
  for (int i = tid; i < N1; i += kImpThreads)
    if (t_flags[i] & 2) mx = fmaxf(mx, Sf[i]);
  mx = block_max(mx, red);
  float out = mx;
  if (mx != CUDART_INF_F && mx != -CUDART_INF_F) {
    float sum = 0.0f;
    for (int i = tid; i < N1; i += kImpThreads)
      if (t_flags[i] & 2) sum += expf(Sf[i] - mx);
    sum = block_sum(sum, red);
    out = mx + log1pf(sum - 1.0f);
  }
  if (tid == 0) out_scores[blockIdx.x] = out;
  if (bad) atomicOr(&status[blockIdx.x], 1);
}

/* ------------------------------------------------------------------ */
/* backward                                                            */
/* ------------------------------------------------------------------ */

__global__ void __launch_bounds__(kImpThreads) implicit_backward_kernel(
    const GraphMeta* __restrict__ meta,
    const uint8_t* __restrict__ sg_flags,
    const int32_t* __restrict__ sg_in_ptr,
    const int32_t* __restrict__ sg_in_src,
    const int32_t* __restrict__ sg_in_label,
    const float* __restrict__ sg_in_w,
    const float* __restrict__ emissions,
    const float* __restrict__ scores,
    const float* __restrict__ out_scores,
    const float* __restrict__ deltas,
    float* __restrict__ grad_emis,
    long long grad_stride,
    int C,
    const ImpLayout lay) {
  extern __shared__ __align__(16) unsigned char smem[];
  const GraphMeta m = meta[blockIdx.x];
  const int tid = threadIdx.x;
  const int N1 = m.sg_N, A1 = m.sg_A, T = m.T;
  const int pitch = (N1 + 3) & ~3;
  float* S = reinterpret_cast<float*>(smem + lay.off_S); // [2][pitch]: rows of saved scores
  int* t_inptr = reinterpret_cast<int*>(smem + lay.off_inptr);
  int* t_src = reinterpret_cast<int*>(smem + lay.off_src);
  float* t_w = reinterpret_cast<float*>(smem + lay.off_w);
  int* t_lab = reinterpret_cast<int*>(smem + lay.off_lab);
  unsigned char* t_flags = smem + lay.off_flags;
  float* red = reinterpret_cast<float*>(smem + lay.off_red);
  int* t_outptr = reinterpret_cast<int*>(smem + lay.off_outptr);
  int* t_outent = reinterpret_cast<int*>(smem + lay.off_outent);
  int* t_cursor = reinterpret_cast<int*>(smem + lay.off_cursor);
  float* CUR = reinterpret_cast<float*>(smem + lay.off_cur); // [2][tabA]: arc gradients of a level
  float* NG = reinterpret_cast<float*>(smem + lay.off_ng); // node gradients (general path)

  load_tables(m, sg_flags, sg_in_ptr, sg_in_src, sg_in_label, sg_in_w, t_inptr, t_src, t_w, t_lab, t_flags);

  // out-arc lists of the graph operand (in-entries grouped by source, ascending): the node
  // gradients are gathered through them, fixed order, no floating-point atomics
  for (int i = tid; i <= N1; i += kImpThreads) t_outptr[i] = 0;
  __syncthreads();
  for (int e = tid; e < A1; e += kImpThreads) atomicAdd(&t_outptr[t_src[e] + 1], 1);
  __syncthreads();
  if (tid == 0)
    for (int u = 0; u < N1; u++) t_outptr[u + 1] += t_outptr[u];
  __syncthreads();
  for (int u = tid; u < N1; u += kImpThreads) t_cursor[u] = t_outptr[u];
  __syncthreads();
  for (int e = tid; e < A1; e += kImpThreads) t_outent[atomicAdd(&t_cursor[t_src[e]], 1)] = e;
  __syncthreads();
  for (int u = tid; u < N1; u += kImpThreads) {
    const int q0 = t_outptr[u], q1 = t_outptr[u + 1];
    for (int q = q0 + 1; q < q1; q++) {
      const int v = t_outent[q];
      int p = q - 1;
      while (p >= q0 && t_outent[p] > v) {
        t_outent[p + 1] = t_outent[p];
        p--;
      }
      t_outent[p + 1] = v;
    }
  }
  __syncthreads();

  const float out = out_scores[blockIdx.x];
  if (!finite_f(out) || T < 1) return; // no accepting path (empty lattice): no gradient
  const float delta = deltas ? deltas[blockIdx.x] : 1.0f;
  const float* sc = scores + m.node_base;
  const float* em = emissions + m.emis_off;
  float* gem = grad_emis + (long long)blockIdx.x * grad_stride;
  // seeds: nodeGrad[accept] = exp(score - maxA) / exp(out - maxA) (shortest.cpp:49-60)
  float mxa = ninf();
  for (int i = tid; i < N1; i += kImpThreads)
    if (t_flags[i] & 2) mxa = fmaxf(mxa, sc[(long long)T * pitch + i]);
  mxa = block_max(mxa, red);
  const float denom = expf(out - mxa);

  int my_deg = 0, my_od = 0;
  if (tid < N1) {
    my_deg = t_inptr[tid + 1] - t_inptr[tid];
    my_od = t_outptr[tid + 1] - t_outptr[tid];
  }
  const bool fast = (N1 <= kImpThreads) && !__syncthreads_or(my_deg > 3 || my_od > 3);
  float* Sa = S;
  float* Sb = S + (fast ? kFastRow : lay.pitch);
  float* CURa = CUR;
  float* CURb = CUR + (fast ? kFastCur : lay.tabA);

  if (fast) {
    const bool act = tid < N1;
    const int deg = my_deg, od = my_od;
    const bool has = deg > 0;
    int i0 = kRowDummy, i1 = kRowDummy, i2 = kRowDummy; // sources of the in-arcs
    int a0 = kCurDummy, a1 = kCurDummy, a2 = kCurDummy; // in-entries (arc-gradient slots)
    int q0 = kCurDummy, q1 = kCurDummy, q2 = kCurDummy; // out-arcs (in-entries of the successors)
    float w0 = 0.0f, w1 = 0.0f, w2 = 0.0f;
    if (has) {
      const int e0 = t_inptr[tid];
      i0 = t_src[e0];
      w0 = t_w[e0];
      a0 = e0;
      if (deg > 1) {
        i1 = t_src[e0 + 1];
        w1 = t_w[e0 + 1];
        a1 = e0 + 1;
      }
      if (deg > 2) {
        i2 = t_src[e0 + 2];
        w2 = t_w[e0 + 2];
        a2 = e0 + 2;
      }
    }
    if (od > 0) q0 = t_outent[t_outptr[tid]];
    if (od > 1) q1 = t_outent[t_outptr[tid] + 1];
    if (od > 2) q2 = t_outent[t_outptr[tid] + 2];
    const float* eml = em + (has ? t_lab[tid] : 0);
    const float* scu = sc + (act ? tid : 0);

    // iteration i handles level t = T - i and prepares level t - 1, for which it needs frame
    // f = t - 2 = T - 2 - i of the saved scores and of the emissions: ring slot i % kPf
    float ps[kPf], pe[kPf];
#pragma unroll
    for (int j = 0; j < kPf; j++) {
      const int f = T - 2 - j;
      ps[j] = (act && f >= 0) ? __ldg(scu + (long long)f * pitch) : ninf();
      pe[j] = (has && f >= 0) ? __ldg(eml + (long long)f * C) : 0.0f;
    }
    const float* spf = scu + (long long)(T - 2 - kPf) * pitch; // only dereferenced while >= row 0
    const float* epf = eml + (long long)(T - 2 - kPf) * C;
    // prologue: the arc factors of level T need row T-1 in shared memory
    float s_own = act ? __ldg(scu + (long long)T * pitch) : ninf(); // S_T[u]
    float s_nxt = act ? __ldg(scu + (long long)(T - 1) * pitch) : ninf(); // S_{T-1}[u]
    Sb[tid] = s_nxt;
    if (tid == 0) {
      Sa[kRowDummy] = Sb[kRowDummy] = ninf();
      CURa[kCurDummy] = CURb[kCurDummy] = 0.0f;
    }
    float g = (act && (t_flags[tid] & 2)) ? expf(s_own - mxa) / denom : 0.0f;
    __syncthreads();
    float E0, E1, E2;
    {
      const float e = has ? __ldg(eml + (long long)(T - 1) * C) : 0.0f;
      const float se = (s_own == -CUDART_INF_F) ? CUDART_INF_F : s_own;
      E0 = iexp(Sb[i0] + (w0 + e) - se);
      E1 = iexp(Sb[i1] + (w1 + e) - se);
      E2 = iexp(Sb[i2] + (w2 + e) - se);
    }
    float* gl = gem + (long long)(T - 1) * C + (has ? t_lab[tid] : 0); // emission gradient of level t
    int i0_ = 0;
    // buffers alternate with the iteration parity: iteration i writes row buffer (i & 1 ? Sb : Sa)
    // and arc-gradient buffer (i & 1 ? CURb : CURa) before its barrier and reads them after it
#define GTNB_BWD_STEP(j, GUARDED)                                                     \
  {                                                                                   \
    float* Cw = ((j)&1) ? CURb : CURa;                                                \
    float* Sw = ((j)&1) ? Sb : Sa;                                                    \
    const float c0 = g * E0, c1 = g * E1, c2 = g * E2;                                \
    Cw[a0] = c0;                                                                      \
    Cw[a1] = c1;                                                                      \
    Cw[a2] = c2;                                                                      \
    const float acc = (c0 + c1) + c2;                                                 \
    red_if(acc != 0.0f, gl, acc * delta);                                             \
    gl -= C;                                                                          \
    const float s_f = ps[j]; /* S_{t-2}[u] */                                         \
    const float e_f = has ? pe[j] : 0.0f; /* e[t-2][label] */                         \
    Sw[tid] = s_f;                                                                    \
    if (GUARDED) {                                                                    \
      const bool inb = T - 2 - (i0_ + (j)) - kPf >= 0;                                \
      ldg_if(act && inb, ps[j], spf);                                                 \
      ldg_if(has && inb, pe[j], epf);                                                 \
    } else {                                                                          \
      ldg_keep(ps[j], spf);                                                           \
      ldg_keep(pe[j], epf);                                                           \
    }                                                                                 \
    spf -= pitch;                                                                     \
    epf -= C;                                                                         \
    __syncthreads();                                                                  \
    g = (Cw[q0] + Cw[q1]) + Cw[q2]; /* node gradient of level t-1: the serial chain */ \
    s_own = s_nxt;                                                                    \
    s_nxt = s_f;                                                                      \
    { /* arc factors of level t-1; a node that does not exist (score -inf) gets exp(-inf) */ \
      const float se = (s_own == -CUDART_INF_F) ? CUDART_INF_F : s_own;               \
      E0 = iexp(Sw[i0] + (w0 + e_f) - se);                                            \
      E1 = iexp(Sw[i1] + (w1 + e_f) - se);                                            \
      E2 = iexp(Sw[i2] + (w2 + e_f) - se);                                            \
    }                                                                                 \
  }
    for (; i0_ + 2 * kPf + 1 <= T; i0_ += kPf) {
#pragma unroll
      for (int j = 0; j < kPf; j++) GTNB_BWD_STEP(j, false)
    }
    for (; i0_ < T; i0_ += kPf) {
#pragma unroll
      for (int j = 0; j < kPf; j++) {
        if (i0_ + j >= T) break;
        GTNB_BWD_STEP(j, true)
      }
    }
#undef GTNB_BWD_STEP
  } else {
    // ---- general path: several nodes per thread, two barriers per level
    for (int i = tid; i < N1; i += kImpThreads) {
      const float sT = sc[(long long)T * pitch + i];
      Sa[i] = sT;
      NG[i] = (t_flags[i] & 2) ? expf(sT - mxa) / denom : 0.0f;
    }
    float* Sown = Sa; // row t
    float* Sprev = Sb; // row t-1
    for (int t = T; t >= 1; t--) {
      const float* er = em + (long long)(t - 1) * C;
      float* gr = gem + (long long)(t - 1) * C;
      for (int i = tid; i < N1; i += kImpThreads) Sprev[i] = sc[(long long)(t - 1) * pitch + i];
      __syncthreads();
      for (int i = tid; i < N1; i += kImpThreads) {
        const int e0 = t_inptr[i], e1 = t_inptr[i + 1];
        if (e1 == e0) continue;
        const float gi = NG[i], sn = Sown[i];
        const bool ex = sn > -CUDART_INF_F;
        const float e = __ldg(er + t_lab[i]);
        float acc = 0.0f;
#pragma unroll 1
        for (int a = e0; a < e1; a++) {
          const float c = ex ? gi * iexp(Sprev[t_src[a]] + (t_w[a] + e) - sn) : 0.0f;
          CURa[a] = c;
          acc += c;
        }
        if (acc != 0.0f) atomicAdd(gr + t_lab[i], acc * delta);
      }
      __syncthreads();
      for (int u = tid; u < N1; u += kImpThreads) {
        float gn = 0.0f;
#pragma unroll 1
        for (int q = t_outptr[u]; q < t_outptr[u + 1]; q++) gn += CURa[t_outent[q]];
        NG[u] = gn;
      }
      float* ts = Sown;
      Sown = Sprev;
      Sprev = ts;
      // the next iteration's first barrier orders NG / Sprev writes before their reads
    }
  }
}

} // namespace

bool implicit_dims_supported(const SgDims* dims, int n_graphs) {
  int maxN = 0, maxA = 0;
  for (int g = 0; g < n_graphs; g++) {
    if (!dims[g].uniform || !dims[g].all_valid) return false;
    maxN = std::max(maxN, dims[g].N);
    maxA = std::max(maxA, dims[g].A);
  }
  return make_imp_layout(maxN, maxA, true).total <= 200 * 1024;
}

/* true when every graph of the batch can take the implicit path */
bool implicit_supported(const gtnb_lattice* lat) {
  if (!lat->composed) return false;
  for (int b = 0; b < lat->B; b++)
    if (!lat->meta_h[b].sg_uniform || !lat->meta_h[b].sg_all_valid) return false;
  const ImpLayout lay = make_imp_layout(lat->max_lvl_nodes, lat->max_lvl_arcs, true);
  return lay.total <= 200 * 1024;
}

/* utterances [b0, b0 + nb) of the batch (nb < 0: all): one CTA each, on ctx->stream */
int launch_implicit_forward(gtnb_ctx* ctx, gtnb_lattice* lat, int32_t* status_dev, int b0, int nb) {
  if (nb < 0) nb = lat->B - b0;
  if (nb <= 0) return GTNB_OK;
  const ImpLayout lay = make_imp_layout(lat->max_lvl_nodes, lat->max_lvl_arcs, false);
  if (lay.total > 48 * 1024) {
    int rc = ensure_max_smem(ctx, (const void*)implicit_forward_kernel);
    if (rc) return rc;
  }
  GTNB_LAUNCH(ctx, "implicit_forward",
              implicit_forward_kernel<<<nb, kImpThreads, lay.total, ctx->stream>>>(
                  lat->meta + b0, lat->sg_flags, lat->sg_in_ptr, lat->sg_in_src, lat->sg_in_label, lat->sg_in_w,
                  lat->emissions, lat->scores, lat->out_scores + b0, status_dev + b0, lat->C, lay));
  return GTNB_OK;
}

int launch_implicit_backward(
    gtnb_ctx* ctx, gtnb_lattice* lat, const float* deltas_dev, float* grad_emis, int64_t grad_stride,
    int b0, int nb) {
  if (nb < 0) nb = lat->B - b0;
  if (nb <= 0) return GTNB_OK;
  const ImpLayout lay = make_imp_layout(lat->max_lvl_nodes, lat->max_lvl_arcs, true);
  if (lay.total > 48 * 1024) {
    int rc = ensure_max_smem(ctx, (const void*)implicit_backward_kernel);
    if (rc) return rc;
  }
  GTNB_LAUNCH(ctx, "implicit_backward",
              implicit_backward_kernel<<<nb, kImpThreads, lay.total, ctx->stream>>>(
                  lat->meta + b0, lat->sg_flags, lat->sg_in_ptr, lat->sg_in_src, lat->sg_in_label, lat->sg_in_w,
                  lat->emissions, lat->scores, lat->out_scores + b0, deltas_dev ? deltas_dev + b0 : nullptr,
                  grad_emis + (long long)b0 * grad_stride, (long long)grad_stride, lat->C, lay));
  return GTNB_OK;
}

} // namespace gtnb

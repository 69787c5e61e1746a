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
#ifdef GTNB_HOST_EMU // this file compiled by g++ against tests/emu/simt_emu.h (CPU test suite)
#include <algorithm>

#include "gtnb_meta.h"
#include "simt_emu.h"
#else
#include <cuda_runtime.h>
#include <math_constants.h>

#include <algorithm>

#include "gtnb_internal.h"
#endif

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
  int off_outptr, off_outent, off_cursor, off_cur, off_ng, off_accg; // backward only
  int off_sw, off_act, off_erow, off_rows3; // wide kernels
  int total;
};

ImpLayout make_imp_layout(int maxN, int maxA, bool backward, bool wide = false, int C = 0) {
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
  o.off_outptr = o.off_outent = o.off_cursor = o.off_cur = o.off_ng = o.off_accg = 0;
  o.off_sw = o.off_act = o.off_erow = o.off_rows3 = 0;
  if (backward) {
    o.off_outptr = take(4 * (o.tabN + 1));
    o.off_outent = take(4 * o.tabA);
    o.off_cursor = take(4 * o.tabN);
    o.off_cur = take(4 * 2 * std::max(o.tabA, kFastCur));
    o.off_ng = take(4 * o.pitch);
    o.off_accg = take(4 * o.tabA);
  }
  if (wide) {
    o.off_sw = take(8 * o.tabA);
    o.off_act = take(4 * 2 * o.tabN);
    o.off_erow = take(4 * 2 * std::max(C, 4));
    o.off_rows3 = take(4 * 3 * o.pitch);
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
/* the two SFU primitives (the host emulation of tests/emu substitutes exp2f / log2f) */
#ifdef GTNB_HOST_EMU
__device__ __forceinline__ float sfu_ex2(float x) {
  return exp2f(x);
}
__device__ __forceinline__ float sfu_lg2(float x) {
  return log2f(x);
}
#else
__device__ __forceinline__ float sfu_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sfu_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
#endif
/* exp / log1p on the SFU, as in k_staged.cu */
__device__ __forceinline__ float iexp(float x) {
  return sfu_ex2(x * 1.4426950408889634f);
}
__device__ __forceinline__ float ilog1p(float x) {
  const float y = sfu_lg2(1.0f + x);
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
#ifdef GTNB_HOST_EMU
  if (p) dst = *src;
  return;
#else
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "setp.ne.s32 q, %2, 0;\n"
      "@q ld.global.nc.f32 %0, [%1];\n"
      "}\n"
      : "+f"(dst)
      : "l"(src), "r"((int)p));
#endif
}
/* (Measured alternatives, both slower at B=256: a 2-deep register ring behind prefetch.global.L1
 * issued 8 frames ahead -- 0.33 vs 0.27 ms for the backward; prefetching the three source scores
 * of every arc into registers instead of staging the row in shared memory -- 0.36 ms.) */
/* Unpredicated variant for the steady part of the frame loop (address always valid): a strong
 * (relaxed, CTA scope) load, which ptxas must keep on its side of the bar.sync it was written
 * on, writing the loop-carried ring register directly. */
__device__ __forceinline__ void ldg_keep(float& dst, const float* src) {
#ifdef GTNB_HOST_EMU
  dst = *src;
#else
  asm volatile("ld.relaxed.cta.global.f32 %0, [%1];" : "=f"(dst) : "l"(src) : "memory");
#endif
}
/* predicated red.global.add.f32 (no branch, no reconvergence point in the frame loop) */
__device__ __forceinline__ void red_if(bool p, float* dst, float v) {
#ifdef GTNB_HOST_EMU
  if (p) atomicAdd(dst, v);
  return;
#else
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "setp.ne.s32 q, %2, 0;\n"
      "@q red.global.add.f32 [%0], %1;\n"
      "}\n" ::"l"(dst),
      "f"(v), "r"((int)p)
      : "memory");
#endif
}

/* Barrier over the first `nthreads` threads of the CTA only (a multiple of 32): warps that hold no
 * graph node leave the frame loop altogether instead of issuing its instruction stream for
 * nothing (CTC at U=100: 201 nodes = 7 of the 8 warps; the 8th cost 1/8 of the issue slots).
 * Named barrier 1; __syncthreads() (barrier 0) keeps synchronising the whole CTA elsewhere. */
__device__ __forceinline__ void bar_nodes(int nthreads) {
#ifdef GTNB_HOST_EMU
  emu::named_barrier(1, nthreads);
#else
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
#endif
}

/* three in-arcs, branch-free: absent arcs carry -inf and add exp(-inf) = 0 in arc order
 * (shortest.cpp:102-114); an all--inf (or +inf) maximum is returned as is.  The sum is >= 1
 * (the maximum contributes exp(0)), where lg2.approx has an absolute error <= 2^-22: far
 * below one ulp of any score this recursion produces after a few frames. */
__device__ __forceinline__ float lse3(float v0, float v1, float v2) {
  const float mx = fmaxf(fmaxf(v0, v1), v2);
  const float sum = (iexp(v0 - mx) + iexp(v1 - mx)) + iexp(v2 - mx);
  const float r = fmaf(sfu_lg2(sum), 0.6931471805599453f, mx);
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
  GTNB_DYNAMIC_SMEM(unsigned char, smem);
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
    // Emissions of the node's label reach the thread through two register blocks: pf[] holds
    // frames f0 .. f0+kPf-1 (complete), nx[] frames f0+kPf .. f0+2kPf-1, in flight since the
    // previous block boundary.  At a block boundary nx[] is copied to pf[] and the kPf loads of
    // the block after next are issued into nx[], all together.  ptxas tracks every one of these
    // loads on ONE scoreboard and waits for it once per block, so a load issued one frame before
    // that wait (the former one-load-per-frame ring) stalled every block for a full memory
    // latency: 31 % of this kernel's samples sat on that single wait (profiles/r1h_ncu_summary.txt).
    // (Where the refill sits matters to ptxas: at the END of the block here -- all loads on one
    // scoreboard, waited 7 frames later; at the top it spreads them over the scoreboards the
    // LDS / MUFU results use.  The backward sweep needs it at the top.  Checked in SASS.)
    float pf[kPf], nx[kPf];
#pragma unroll
    for (int j = 0; j < kPf; j++) pf[j] = (has && j < T) ? __ldg(eml + (long long)j * C) : 0.0f;
#pragma unroll
    for (int j = 0; j < kPf; j++) nx[j] = (has && kPf + j < T) ? __ldg(eml + (long long)(kPf + j) * C) : 0.0f;
    const float* epf = eml + (long long)(2 * kPf) * C; // first frame of the next refill
    float* grow = sc + pitch + tid; // this node's score in frame f + 1
    float chk = 0.0f; // becomes NaN when a used emission is not finite
    int f0 = 0;
#define GTNB_FWD_STEP(j)                                                 \
  {                                                                      \
    const float e = has ? pf[j] : 0.0f; /* idle threads load node 0's */ \
    const float* P = ((j)&1) ? S1 : S0;                                  \
    float* Q = ((j)&1) ? S0 : S1;                                        \
    const float sv = lse3(P[i0] + (w0 + e), P[i1] + (w1 + e), P[i2] + (w2 + e)); \
    chk = fmaf(e, 0.0f, chk);                                            \
    Q[tid] = sv;                                                         \
    bar_nodes(nact);                                                     \
    if (act) *grow = sv; /* after the barrier: nothing waits for it */   \
    grow += pitch;                                                       \
  }
#define GTNB_FWD_REFILL(GUARDED)                                         \
  {                                                                      \
    _Pragma("unroll") for (int j = 0; j < kPf; j++) pf[j] = nx[j];       \
    _Pragma("unroll") for (int j = 0; j < kPf; j++) {                    \
      if (GUARDED)                                                       \
        ldg_if(has && f0 + 2 * kPf + j < T, nx[j], epf + (long long)j * C); \
      else                                                               \
        ldg_keep(nx[j], epf + (long long)j * C);                         \
    }                                                                    \
    epf += (long long)kPf * C;                                           \
  }
    const int nact = (N1 + 31) & ~31; // threads of the warps that hold nodes
    if (tid < nact) {
      for (; f0 + 3 * kPf <= T; f0 += kPf) {
#pragma unroll
        for (int j = 0; j < kPf; j++) GTNB_FWD_STEP(j)
        GTNB_FWD_REFILL(false)
      }
      for (; f0 < T; f0 += kPf) {
#pragma unroll
        for (int j = 0; j < kPf; j++) {
          if (f0 + j >= T) break;
          GTNB_FWD_STEP(j)
        }
        GTNB_FWD_REFILL(true)
      }
    }
#undef GTNB_FWD_REFILL
#undef GTNB_FWD_STEP
    bad |= !(chk == 0.0f);
    __syncthreads(); // the idle warps meet the others again (S rows are read below)
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

/* out-arc lists of the graph operand (in-entries grouped by source, ascending): the node
 * gradients are gathered through them, fixed order, no floating-point atomics */
__device__ __forceinline__ void build_out_lists(
    int N1, int A1, const int* t_src, int* t_outptr, int* t_outent, int* t_cursor) {
  const int tid = threadIdx.x;
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
}

/* ------------------------------------------------------------------ */
/* backward                                                            */
/* ------------------------------------------------------------------ */

template <bool GRAPH_GRAD>
__global__ void __launch_bounds__(kImpThreads) implicit_backward_kernel(
    const GraphMeta* __restrict__ meta,
    const uint8_t* __restrict__ sg_flags,
    const int32_t* __restrict__ sg_in_ptr,
    const int32_t* __restrict__ sg_in_src,
    const int32_t* __restrict__ sg_in_label,
    const float* __restrict__ sg_in_w,
    const int32_t* __restrict__ sg_in_arc,
    float* __restrict__ grad_graph,
    const float* __restrict__ emissions,
    const float* __restrict__ scores,
    const float* __restrict__ out_scores,
    const float* __restrict__ deltas,
    float* __restrict__ grad_emis,
    long long grad_stride,
    int C,
    const ImpLayout lay) {
  GTNB_DYNAMIC_SMEM(unsigned char, smem);
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
  float* ACCG = reinterpret_cast<float*>(smem + lay.off_accg); // graph-arc gradients (general path)
  const int32_t* in_arc = sg_in_arc + m.sg_arc_base;
  float* ggr = GRAPH_GRAD ? grad_graph + m.grad_graph_off : nullptr;

  load_tables(m, sg_flags, sg_in_ptr, sg_in_src, sg_in_label, sg_in_w, t_inptr, t_src, t_w, t_lab, t_flags);

  build_out_lists(N1, A1, t_src, t_outptr, t_outent, t_cursor);

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
    // f = t - 2 = T - 2 - i of the saved scores and of the emissions.  Two register blocks as in
    // the forward sweep: a block of kPf iterations copies nxs/nxe (loaded during the previous
    // block) to ps/pe and issues all loads of the next block.
    float ps[kPf], pe[kPf], nxs[kPf], nxe[kPf];
#pragma unroll
    for (int j = 0; j < kPf; j++) {
      const int f = T - 2 - j;
      nxs[j] = (act && f >= 0) ? __ldg(scu + (long long)f * pitch) : ninf();
      nxe[j] = (has && f >= 0) ? __ldg(eml + (long long)f * C) : 0.0f;
    }
    // first frame of the next refill; only dereferenced while >= row 0
    const float* spf = scu + (long long)(T - 2 - kPf) * pitch;
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
    float ga0 = 0.0f, ga1 = 0.0f, ga2 = 0.0f; // gradients of the node's in-arcs of the graph operand
    int i0_ = 0;
    // buffers alternate with the iteration parity: iteration i writes row buffer (i & 1 ? Sb : Sa)
    // and arc-gradient buffer (i & 1 ? CURb : CURa) before its barrier and reads them after it
#define GTNB_BWD_STEP(j)                                                              \
  {                                                                                   \
    float* Cw = ((j)&1) ? CURb : CURa;                                                \
    float* Sw = ((j)&1) ? Sb : Sa;                                                    \
    const float c0 = g * E0, c1 = g * E1, c2 = g * E2;                                \
    Cw[a0] = c0;                                                                      \
    Cw[a1] = c1;                                                                      \
    Cw[a2] = c2;                                                                      \
    const float acc = (c0 + c1) + c2;                                                 \
    if (GRAPH_GRAD) {                                                                 \
      ga0 += c0;                                                                      \
      ga1 += c1;                                                                      \
      ga2 += c2;                                                                      \
    }                                                                                 \
    const float s_f = ps[j]; /* S_{t-2}[u] */                                         \
    const float e_f = has ? pe[j] : 0.0f; /* e[t-2][label] */                         \
    Sw[tid] = s_f;                                                                    \
    bar_nodes(nact);                                                                  \
    g = (Cw[q0] + Cw[q1]) + Cw[q2]; /* node gradient of level t-1: the serial chain */ \
    /* the emission gradient of level t leaves after the barrier, in the shadow of the gather \
       (ptxas turns the predicated red into a branch: before the barrier it delayed every   \
       warp's arrival) */                                                             \
    red_if(acc != 0.0f, gl, acc * delta);                                             \
    gl -= C;                                                                          \
    s_own = s_nxt;                                                                    \
    s_nxt = s_f;                                                                      \
    { /* arc factors of level t-1; a node that does not exist (score -inf) gets exp(-inf) */ \
      const float se = (s_own == -CUDART_INF_F) ? CUDART_INF_F : s_own;               \
      E0 = iexp(Sw[i0] + (w0 + e_f) - se);                                            \
      E1 = iexp(Sw[i1] + (w1 + e_f) - se);                                            \
      E2 = iexp(Sw[i2] + (w2 + e_f) - se);                                            \
    }                                                                                 \
  }
#define GTNB_BWD_REFILL(GUARDED)                                                      \
  {                                                                                   \
    _Pragma("unroll") for (int j = 0; j < kPf; j++) {                                 \
      ps[j] = nxs[j];                                                                 \
      pe[j] = nxe[j];                                                                 \
    }                                                                                 \
    _Pragma("unroll") for (int j = 0; j < kPf; j++) {                                 \
      if (GUARDED) {                                                                  \
        const bool inb = T - 2 - (i0_ + kPf + j) >= 0;                                \
        ldg_if(act && inb, nxs[j], spf - (long long)j * pitch);                       \
        ldg_if(has && inb, nxe[j], epf - (long long)j * C);                           \
      } else {                                                                        \
        ldg_keep(nxs[j], spf - (long long)j * pitch);                                 \
        ldg_keep(nxe[j], epf - (long long)j * C);                                     \
      }                                                                               \
    }                                                                                 \
    spf -= (long long)kPf * pitch;                                                    \
    epf -= (long long)kPf * C;                                                        \
  }
    const int nact = (N1 + 31) & ~31; // threads of the warps that hold nodes (see bar_nodes)
    if (tid < nact) {
      for (; i0_ + 2 * kPf + 1 <= T; i0_ += kPf) {
        GTNB_BWD_REFILL(false)
#pragma unroll
        for (int j = 0; j < kPf; j++) GTNB_BWD_STEP(j)
      }
      for (; i0_ < T; i0_ += kPf) {
        GTNB_BWD_REFILL(true)
#pragma unroll
        for (int j = 0; j < kPf; j++) {
          if (i0_ + j >= T) break;
          GTNB_BWD_STEP(j)
        }
      }
    }
#undef GTNB_BWD_REFILL
#undef GTNB_BWD_STEP
    if (GRAPH_GRAD && has) {
      // compose gradFunc, graph side (compose.cpp:500-506): the arc's gradient over all frames
      const int e0 = t_inptr[tid];
      atomicAdd(&ggr[in_arc[e0]], ga0 * delta);
      if (deg > 1) atomicAdd(&ggr[in_arc[e0 + 1]], ga1 * delta);
      if (deg > 2) atomicAdd(&ggr[in_arc[e0 + 2]], ga2 * delta);
    }
  } else {
    // ---- general path: several nodes per thread, two barriers per level
    if (GRAPH_GRAD)
      for (int a = tid; a < A1; a += kImpThreads) ACCG[a] = 0.0f;
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
          if (GRAPH_GRAD) ACCG[a] += c; // entry a belongs to this thread only
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
    if (GRAPH_GRAD) {
      __syncthreads();
      for (int a = tid; a < A1; a += kImpThreads)
        if (ACCG[a] != 0.0f) atomicAdd(&ggr[in_arc[a]], ACCG[a] * delta);
    }
  }
}


/* ------------------------------------------------------------------ */
/* wide graphs (in-/out-degree > 3, e.g. the dense ASG transitions):   */
/* G lanes per node, each holding up to kWideCap arcs in registers     */
/* ------------------------------------------------------------------ */

constexpr int kWideCap = 20;

/* packed {source, weight bits} per in-entry; compact lists of the nodes that have in-arcs
 * (t_act[0..n), n at t_act[N1]) and, for the backward, out-arcs (t_acto likewise) */
__device__ __forceinline__ void build_wide_tables(
    int N1, int A1, const int* t_inptr, const int* t_src, const float* t_w, const int* t_outptr, int2* t_sw,
    int* t_act, int* t_acto) {
  const int tid = threadIdx.x;
  for (int a = tid; a < A1; a += kImpThreads) t_sw[a] = make_int2(t_src[a], __float_as_int(t_w[a]));
  if (tid == 0) {
    int n = 0;
    for (int u = 0; u < N1; u++)
      if (t_inptr[u + 1] > t_inptr[u]) t_act[n++] = u;
    t_act[N1] = n;
  }
  if (tid == 32 && t_outptr) {
    int n = 0;
    for (int u = 0; u < N1; u++)
      if (t_outptr[u + 1] > t_outptr[u]) t_acto[n++] = u;
    t_acto[N1] = n;
  }
  __syncthreads();
}

template <int G>
__device__ __forceinline__ unsigned group_mask() {
  return (G >= 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) & ~(G - 1)));
}

template <int G>
__global__ void __launch_bounds__(kImpThreads) implicit_forward_wide_kernel(
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
  GTNB_DYNAMIC_SMEM(unsigned char, smem);
  const GraphMeta m = meta[blockIdx.x];
  const int tid = threadIdx.x;
  const int N1 = m.sg_N, A1 = m.sg_A, T = m.T;
  const int pitch = (N1 + 3) & ~3;
  float* S = reinterpret_cast<float*>(smem + lay.off_S);
  int* t_inptr = reinterpret_cast<int*>(smem + lay.off_inptr);
  int* t_src = reinterpret_cast<int*>(smem + lay.off_src);
  float* t_w = reinterpret_cast<float*>(smem + lay.off_w);
  int* t_lab = reinterpret_cast<int*>(smem + lay.off_lab);
  unsigned char* t_flags = smem + lay.off_flags;
  float* red = reinterpret_cast<float*>(smem + lay.off_red);
  int2* t_sw = reinterpret_cast<int2*>(smem + lay.off_sw);
  int* t_act = reinterpret_cast<int*>(smem + lay.off_act);
  float* E_s = reinterpret_cast<float*>(smem + lay.off_erow); // [2][C]: emission rows, one frame ahead

  bool bad = load_tables(m, sg_flags, sg_in_ptr, sg_in_src, sg_in_label, sg_in_w, t_inptr, t_src, t_w,
                         t_lab, t_flags);
  build_wide_tables(N1, A1, t_inptr, t_src, t_w, nullptr, t_sw, t_act, nullptr);
  float* sc = scores + m.node_base;
  const float* em = emissions + m.emis_off;
  float* prev = S;
  float* cur = S + lay.pitch;
  for (int i = tid; i < N1; i += kImpThreads) {
    const float v = (t_flags[i] & 1) ? 0.0f : ninf();
    prev[i] = v;
    sc[i] = v;
  }
  if (T > 0 && tid < C) {
    const float e0v = __ldg(em + tid);
    E_s[tid] = e0v;
    bad |= !finite_f(e0v);
  }
  __syncthreads();

  constexpr int kSlots = kImpThreads / G;
  const int sub = tid % G, slot0 = tid / G;
  const unsigned gmask = group_mask<G>();
  const int n_act = t_act[N1];
  const int rounds = (n_act + kSlots - 1) / kSlots;
  for (int t = 1; t <= T; t++) {
    const float* Ecur = E_s + ((t - 1) & 1) * C;
    float* Enext = E_s + (t & 1) * C;
    const bool ld = (t < T) && tid < C; // next frame's emissions: in flight during this level
    float enext = 0.0f;
    if (ld) enext = __ldg(em + (long long)t * C + tid);
    float* srow = sc + (long long)t * pitch;
    for (int r = 0; r < rounds; r++) {
      const int slot = r * kSlots + slot0;
      const bool on = slot < n_act;
      const int u = on ? t_act[slot] : 0;
      const int e0 = on ? t_inptr[u] : 0, e1 = on ? t_inptr[u + 1] : 0;
      const float e = on ? Ecur[t_lab[u]] : 0.0f;
      float v[kWideCap];
      float mx = ninf();
#pragma unroll
      for (int k = 0; k < kWideCap; k++) {
        const int a = e0 + sub + k * G;
        v[k] = ninf();
        if (a < e1) {
          const int2 sw = t_sw[a];
          v[k] = prev[sw.x] + (__int_as_float(sw.y) + e);
        }
        mx = fmaxf(mx, v[k]);
      }
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(gmask, mx, o));
      float sum = 0.0f;
#pragma unroll
      for (int k = 0; k < kWideCap; k++) sum += iexp(v[k] - mx);
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(gmask, sum, o);
      const float lg = sfu_lg2(sum);
      const float sv = (fabsf(mx) == CUDART_INF_F) ? mx : fmaf(lg, 0.6931471805599453f, mx);
      if (on && sub == 0) {
        cur[u] = sv;
        srow[u] = sv;
      }
    }
    // a node without in-arcs does not exist past frame 0
    for (int i = tid; i < N1; i += kImpThreads)
      if (t_inptr[i + 1] == t_inptr[i]) {
        cur[i] = ninf();
        srow[i] = ninf();
      }
    if (ld) {
      Enext[tid] = enext;
      bad |= !finite_f(enext);
    }
    float* tmp = prev;
    prev = cur;
    cur = tmp;
    __syncthreads();
  }

  // accept nodes of the last frame (shortest.cpp:147-159); `prev` is the last row written
  const float* Sf = prev;
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

template <int G, bool GRAPH_GRAD>
__global__ void __launch_bounds__(kImpThreads) implicit_backward_wide_kernel(
    const GraphMeta* __restrict__ meta,
    const uint8_t* __restrict__ sg_flags,
    const int32_t* __restrict__ sg_in_ptr,
    const int32_t* __restrict__ sg_in_src,
    const int32_t* __restrict__ sg_in_label,
    const float* __restrict__ sg_in_w,
    const int32_t* __restrict__ sg_in_arc,
    float* __restrict__ grad_graph,
    const float* __restrict__ emissions,
    const float* __restrict__ scores,
    const float* __restrict__ out_scores,
    const float* __restrict__ deltas,
    float* __restrict__ grad_emis,
    long long grad_stride,
    int C,
    const ImpLayout lay) {
  GTNB_DYNAMIC_SMEM(unsigned char, smem);
  const GraphMeta m = meta[blockIdx.x];
  const int tid = threadIdx.x;
  const int N1 = m.sg_N, A1 = m.sg_A, T = m.T;
  const int pitch = (N1 + 3) & ~3;
  int* t_inptr = reinterpret_cast<int*>(smem + lay.off_inptr);
  int* t_src = reinterpret_cast<int*>(smem + lay.off_src);
  float* t_w = reinterpret_cast<float*>(smem + lay.off_w);
  int* t_lab = reinterpret_cast<int*>(smem + lay.off_lab);
  unsigned char* t_flags = smem + lay.off_flags;
  float* red = reinterpret_cast<float*>(smem + lay.off_red);
  int* t_outptr = reinterpret_cast<int*>(smem + lay.off_outptr);
  int* t_outent = reinterpret_cast<int*>(smem + lay.off_outent);
  int* t_cursor = reinterpret_cast<int*>(smem + lay.off_cursor);
  float* CUR = reinterpret_cast<float*>(smem + lay.off_cur);
  float* NG = reinterpret_cast<float*>(smem + lay.off_ng);
  float* ACCG = reinterpret_cast<float*>(smem + lay.off_accg);
  int2* t_sw = reinterpret_cast<int2*>(smem + lay.off_sw);
  int* t_act = reinterpret_cast<int*>(smem + lay.off_act);
  int* t_acto = t_act + lay.tabN;
  float* E_s = reinterpret_cast<float*>(smem + lay.off_erow);
  float* R3 = reinterpret_cast<float*>(smem + lay.off_rows3); // three score rows: t, t-1, staging
  const int32_t* in_arc = sg_in_arc + m.sg_arc_base;
  float* ggr = GRAPH_GRAD ? grad_graph + m.grad_graph_off : nullptr;

  load_tables(m, sg_flags, sg_in_ptr, sg_in_src, sg_in_label, sg_in_w, t_inptr, t_src, t_w, t_lab, t_flags);
  build_out_lists(N1, A1, t_src, t_outptr, t_outent, t_cursor);
  build_wide_tables(N1, A1, t_inptr, t_src, t_w, t_outptr, t_sw, t_act, t_acto);

  const float out = out_scores[blockIdx.x];
  if (!finite_f(out) || T < 1) return; // no accepting path (empty lattice): no gradient
  const float delta = deltas ? deltas[blockIdx.x] : 1.0f;
  const float* sc = scores + m.node_base;
  const float* em = emissions + m.emis_off;
  float* gem = grad_emis + (long long)blockIdx.x * grad_stride;

  float mxa = ninf();
  for (int i = tid; i < N1; i += kImpThreads)
    if (t_flags[i] & 2) mxa = fmaxf(mxa, sc[(long long)T * pitch + i]);
  mxa = block_max(mxa, red);
  const float denom = expf(out - mxa);

  float* own = R3; // row t
  float* prv = R3 + lay.pitch; // row t-1
  float* nxt = R3 + 2 * lay.pitch; // row t-2, staged during level t
  for (int i = tid; i < N1; i += kImpThreads) {
    const float sT = sc[(long long)T * pitch + i];
    own[i] = sT;
    prv[i] = sc[(long long)(T - 1) * pitch + i];
    NG[i] = (t_flags[i] & 2) ? expf(sT - mxa) / denom : 0.0f; // shortest.cpp:49-60
  }
  if (GRAPH_GRAD)
    for (int a = tid; a < A1; a += kImpThreads) ACCG[a] = 0.0f;
  if (tid < C) E_s[((T - 1) & 1) * C + tid] = __ldg(em + (long long)(T - 1) * C + tid);
  __syncthreads();

  constexpr int kSlots = kImpThreads / G;
  const int sub = tid % G, slot0 = tid / G;
  const unsigned gmask = group_mask<G>();
  const int n_act = t_act[N1], n_acto = t_acto[N1];
  const int rounds = (n_act + kSlots - 1) / kSlots, rounds_o = (n_acto + kSlots - 1) / kSlots;
  for (int t = T; t >= 1; t--) {
    const float* Ecur = E_s + ((t - 1) & 1) * C; // frame t-1
    float* Enext = E_s + (t & 1) * C; // frame t-2 (same parity as t)
    const bool more = t >= 2;
    float snext = ninf(), enext = 0.0f; // in flight during this level
    if (more && tid < N1) snext = __ldg(sc + (long long)(t - 2) * pitch + tid);
    if (more && tid < C) enext = __ldg(em + (long long)(t - 2) * C + tid);
    float* gr = gem + (long long)(t - 1) * C;
    // phase 1: arc gradients of level t (shortest.cpp:62-80), emission gradient per node
    for (int r = 0; r < rounds; r++) {
      const int slot = r * kSlots + slot0;
      const bool on = slot < n_act;
      const int u = on ? t_act[slot] : 0;
      const int e0 = on ? t_inptr[u] : 0, e1 = on ? t_inptr[u + 1] : 0;
      const int lab = on ? t_lab[u] : 0;
      const float e = on ? Ecur[lab] : 0.0f;
      const float g = NG[u], sn = own[u];
      const float se = (sn == -CUDART_INF_F) ? CUDART_INF_F : sn; // missing node: factors exp(-inf)
      float acc = 0.0f;
#pragma unroll
      for (int k = 0; k < kWideCap; k++) {
        const int a = e0 + sub + k * G;
        if (a < e1) {
          const int2 sw = t_sw[a];
          const float c = g * iexp(prv[sw.x] + (__int_as_float(sw.y) + e) - se);
          CUR[a] = c;
          acc += c;
          if (GRAPH_GRAD) ACCG[a] += c; // entry a belongs to this lane only
        }
      }
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(gmask, acc, o);
      if (on && sub == 0 && acc != 0.0f) atomicAdd(gr + lab, acc * delta);
    }
    __syncthreads();
    // phase 2: node gradients of frame t-1, gathered over the out-arcs
    for (int r = 0; r < rounds_o; r++) {
      const int slot = r * kSlots + slot0;
      const bool on = slot < n_acto;
      const int u = on ? t_acto[slot] : 0;
      const int o0 = on ? t_outptr[u] : 0, o1 = on ? t_outptr[u + 1] : 0;
      float sum = 0.0f;
#pragma unroll
      for (int k = 0; k < kWideCap; k++) {
        const int q = o0 + sub + k * G;
        if (q < o1) sum += CUR[t_outent[q]];
      }
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(gmask, sum, o);
      if (on && sub == 0) NG[u] = sum;
    }
    for (int i = tid; i < N1; i += kImpThreads)
      if (t_outptr[i + 1] == t_outptr[i]) NG[i] = 0.0f;
    if (tid < N1) nxt[tid] = snext;
    if (more && tid < C) Enext[tid] = enext;
    float* tmp = own;
    own = prv;
    prv = nxt;
    nxt = tmp;
    __syncthreads();
  }
  if (GRAPH_GRAD)
    for (int a = tid; a < A1; a += kImpThreads)
      if (ACCG[a] != 0.0f) atomicAdd(&ggr[in_arc[a]], ACCG[a] * delta);
}

} // namespace

namespace {

/* 0: one node per thread / several nodes per thread (implicit_*_kernel decides per CTA);
 * otherwise the lanes per node of the wide kernels */
int wide_lanes(int max_in, int max_out, int maxN, int C, bool backward) {
  const int deg = backward ? std::max(max_in, max_out) : max_in;
  if (deg <= 3 || maxN > kImpThreads || C > kImpThreads) return 0;
  int G = 1;
  while (G < 32 && (deg + G - 1) / G > kWideCap) G *= 2;
  if ((deg + G - 1) / G > kWideCap) return 0;
  return G;
}

} // namespace

#ifndef GTNB_HOST_EMU

bool implicit_dims_supported(const SgDims* dims, int n_graphs) {
  int maxN = 0, maxA = 0;
  for (int g = 0; g < n_graphs; g++) {
    if (!dims[g].uniform || !dims[g].all_valid) return false;
    maxN = std::max(maxN, dims[g].N);
    maxA = std::max(maxA, dims[g].A);
  }
  return make_imp_layout(maxN, maxA, true, true, kImpThreads).total <= 200 * 1024;
}

/* true when every graph of the batch can take the implicit path */
bool implicit_supported(const gtnb_lattice* lat) {
  if (!lat->composed) return false;
  for (int b = 0; b < lat->B; b++)
    if (!lat->meta_h[b].sg_uniform || !lat->meta_h[b].sg_all_valid) return false;
  const ImpLayout lay = make_imp_layout(lat->max_lvl_nodes, lat->max_lvl_arcs, true, true, kImpThreads);
  return lay.total <= 200 * 1024;
}

#define GTNB_FWD_ARGS                                                                             \
  lat->meta + b0, lat->sg_flags, lat->sg_in_ptr, lat->sg_in_src, lat->sg_in_label, lat->sg_in_w, \
      lat->emissions, lat->scores, lat->out_scores + b0, status_dev + b0, lat->C, lay
#define GTNB_BWD_ARGS                                                                             \
  lat->meta + b0, lat->sg_flags, lat->sg_in_ptr, lat->sg_in_src, lat->sg_in_label, lat->sg_in_w, \
      lat->sg_in_arc, grad_graph, lat->emissions, lat->scores, lat->out_scores + b0,             \
      deltas_dev ? deltas_dev + b0 : nullptr, grad_emis + (long long)b0 * grad_stride,           \
      (long long)grad_stride, lat->C, lay

/* utterances [b0, b0 + nb) of the batch (nb < 0: all): one CTA each, on ctx->stream */
int launch_implicit_forward(gtnb_ctx* ctx, gtnb_lattice* lat, int32_t* status_dev, int b0, int nb) {
  if (nb < 0) nb = lat->B - b0;
  if (nb <= 0) return GTNB_OK;
  const int G = wide_lanes(lat->max_in_deg, lat->max_out_deg, lat->max_lvl_nodes, lat->C, false);
  const ImpLayout lay = make_imp_layout(lat->max_lvl_nodes, lat->max_lvl_arcs, false, G > 0, lat->C);
#define GTNB_LAUNCH_FWD(K)                                                                \
  do {                                                                                    \
    if (lay.total > 48 * 1024) {                                                          \
      int rc = ensure_max_smem(ctx, (const void*)K);                                      \
      if (rc) return rc;                                                                  \
    }                                                                                     \
    GTNB_LAUNCH(ctx, "implicit_forward", K<<<nb, kImpThreads, lay.total, ctx->stream>>>(GTNB_FWD_ARGS)); \
  } while (0)
  switch (G) {
    case 0: GTNB_LAUNCH_FWD(implicit_forward_kernel); break;
    case 1: GTNB_LAUNCH_FWD(implicit_forward_wide_kernel<1>); break;
    case 2: GTNB_LAUNCH_FWD(implicit_forward_wide_kernel<2>); break;
    case 4: GTNB_LAUNCH_FWD(implicit_forward_wide_kernel<4>); break;
    case 8: GTNB_LAUNCH_FWD(implicit_forward_wide_kernel<8>); break;
    case 16: GTNB_LAUNCH_FWD(implicit_forward_wide_kernel<16>); break;
    default: GTNB_LAUNCH_FWD(implicit_forward_wide_kernel<32>); break;
  }
#undef GTNB_LAUNCH_FWD
  return GTNB_OK;
}

/* grad_graph (may be NULL): gradients w.r.t. the graph operands' arc weights, slab of graph b at
 * meta[b].grad_graph_off (0 for a graph shared by the batch), accumulated with atomics */
int launch_implicit_backward(
    gtnb_ctx* ctx, gtnb_lattice* lat, const float* deltas_dev, float* grad_emis, int64_t grad_stride,
    int b0, int nb, float* grad_graph) {
  if (nb < 0) nb = lat->B - b0;
  if (nb <= 0) return GTNB_OK;
  const int G = wide_lanes(lat->max_in_deg, lat->max_out_deg, lat->max_lvl_nodes, lat->C, true);
  const ImpLayout lay = make_imp_layout(lat->max_lvl_nodes, lat->max_lvl_arcs, true, G > 0, lat->C);
#define GTNB_LAUNCH_BWD(K)                                                                \
  do {                                                                                    \
    if (lay.total > 48 * 1024) {                                                          \
      int rc = ensure_max_smem(ctx, (const void*)K);                                      \
      if (rc) return rc;                                                                  \
    }                                                                                     \
    GTNB_LAUNCH(ctx, "implicit_backward", K<<<nb, kImpThreads, lay.total, ctx->stream>>>(GTNB_BWD_ARGS)); \
  } while (0)
#define GTNB_LAUNCH_BWD_G(GG)                                               \
  do {                                                                      \
    if (grad_graph)                                                         \
      GTNB_LAUNCH_BWD((implicit_backward_wide_kernel<GG, true>));           \
    else                                                                    \
      GTNB_LAUNCH_BWD((implicit_backward_wide_kernel<GG, false>));          \
  } while (0)
  switch (G) {
    case 0:
      if (grad_graph)
        GTNB_LAUNCH_BWD(implicit_backward_kernel<true>);
      else
        GTNB_LAUNCH_BWD(implicit_backward_kernel<false>);
      break;
    case 1: GTNB_LAUNCH_BWD_G(1); break;
    case 2: GTNB_LAUNCH_BWD_G(2); break;
    case 4: GTNB_LAUNCH_BWD_G(4); break;
    case 8: GTNB_LAUNCH_BWD_G(8); break;
    case 16: GTNB_LAUNCH_BWD_G(16); break;
    default: GTNB_LAUNCH_BWD_G(32); break;
  }
#undef GTNB_LAUNCH_BWD_G
#undef GTNB_LAUNCH_BWD
  return GTNB_OK;
}

#endif // GTNB_HOST_EMU

} // namespace gtnb

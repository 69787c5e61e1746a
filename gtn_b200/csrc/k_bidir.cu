/*
 * k_bidir.cu -- the criterion's hot kernel since round 2: forwardScore(intersect(g, emissions)), its
 * backward, forwardScore(emissions) and its backward -- everything gtnb_ctc_loss needs after the
 * target graphs exist -- in ONE launch, with the serial dependence over T cut in half.
 *
 * The implicit-lattice sweeps of k_implicit.cu run the frames twice in a row (alpha front to back,
 * then shortestDistanceGrad back to front: 2T dependent steps per utterance, and at B = 256 the step
 * time IS that chain's latency).  Here each utterance gets a thread-block cluster of two CTAs that run
 * at the same time and meet in the middle:
 *
 *   CTA A (cluster rank 0)  alpha_s[u] = logsumexp over in-arcs (i -> u, w) of alpha_{s-1}[i] + w + e[s-1][label(u)]
 *                           for levels s = 1 .. T   (shortest.cpp:102-136 on the implicit lattice)
 *   CTA B (cluster rank 1)  beta_s[u]  = logsumexp over out-arcs (u -> v, w) of w + e[s][label(v)] + beta_{s+1}[v]
 *                           for levels s = T .. 1   (beta_T = 0 on accept nodes)
 *
 * With M = T/2 rounded down to a multiple of 8, A saves alpha_1..alpha_M and B saves beta_T..beta_{M+1}
 * (phase 1, T/2 dependent steps each, concurrently); one barrier.cluster later A continues through
 * levels M+1..T reading B's saved beta, B through levels M..1 reading A's saved alpha (phase 2), and the
 * node posterior gamma_s[u] = exp(alpha_s[u] + beta_s[u] - Z) is formed ON THE FLY, off the serial
 * chain.  That posterior is exactly what shortestDistanceGrad + compose's gradFunc deliver to the
 * emissions (shortest.cpp:33-82, compose.cpp:496-518): the arc gradients into lattice node (u, s) sum
 * to its node gradient, and all of them land on emission (s-1, label(u)).  Z = logsumexp_u(alpha + beta)
 * at the meeting level equals forwardScore of the lattice.  The dependent chain is T steps instead of
 * 2T, and HBM traffic does not grow: each half of the saved scores is written once and read once.
 *
 * Inside a CTA (warp-specialised):
 *   node warps    one graph node per thread; per level 3 LDS of the neighbours' scores (absent arcs read
 *                 a slot that holds a large negative finite number, so there is no -inf / NaN special
 *                 casing on the chain), max, 3 ex2, lg2, one STS, one named barrier.  Scores are kept in
 *                 log2 units (emissions and arc weights are multiplied by log2(e) on the way in), which
 *                 removes the multiply from every exp and log.
 *   producer warp cp.async.bulk (TMA 1-D bulk copies, mbarrier complete_tx) of 8 emission frames at a
 *                 time into a 3-stage ring.
 *   helper warps  (2) per block of 8 frames: row logsumexp of the emissions (forwardScore(emissions)
 *                 and its softmax gradient, k_linear.cu's job before), the posteriors summed by label
 *                 through per-label node lists (no floating-point atomics anywhere), and the finished
 *                 gradient rows written with coalesced 16-byte stores.  They also issue the bulk loads of
 *                 the partner's saved scores; the node threads turn those rows into posteriors in place.
 *
 * Numerics: gamma's exponent is evaluated as ((x - Zh) + y) - Zl with x the larger of alpha/beta: both
 * subtractions are exact or nearly so (Sterbenz), so the posterior carries only the rounding error the
 * recursions themselves accumulated -- like the reference's own chain of arc-factor products.  Measured
 * against a float64 evaluation it is closer than the reference's fp32 result (DESIGN.md "Tolerances").
 *
 * Valid when every weight is finite (|x| < 1e29): otherwise a status bit is raised and the caller repeats
 * the batch through the materialised path, which reproduces the reference's inf / NaN propagation arc by
 * arc.  Requires: <= 256 graph nodes, in- and out-degree <= 3, one label per node (CTC, forced alignment),
 * C a multiple of 4 and <= 128, 16-byte aligned emissions.  Anything else takes k_implicit.cu.
 */
#ifdef GTNB_HOST_EMU // this file compiled by g++ against tests/emu/simt_emu.h (CPU test suite)
#include <algorithm>

#include "gtn_b200.h"
#include "gtnb_meta.h"
#include "simt_emu.h"
#else
#include <cuda_runtime.h>
#include <math_constants.h>

#include <algorithm>

#include "gtnb_internal.h"
#endif

namespace gtnb {

namespace bidir {

constexpr int kMaxNodes = 256; // graph nodes per utterance (one per thread)
constexpr int kRowF = 264; // floats per chain row in shared memory
constexpr int kDummy = 256; // slot of every chain row that holds kNeg (absent arcs point here)
constexpr int kBlk = 8; // frames per block (one TMA copy, one helper pass, one renormalisation)
constexpr int kSE = 4; // emission stages
constexpr int kSO = 3; // stages of the partner's saved scores
constexpr int kSG = 2; // posterior blocks between the node warps and the helper warps
constexpr int kHelpers = 2; // helper warps
constexpr int kHeavyLen = 8; // labels carried by more nodes than this are summed cooperatively
constexpr int kMaxHeavy = 4;
constexpr float kNeg = -1.0e30f; // "this node does not exist" (finite: absorbs every later add)
constexpr float kHuge = 1.0e29f; // inputs at or above this magnitude take the exact (materialised) path
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kIntMin = -2147483647 - 1;

struct Layout {
  int off_ring, off_e, off_o, off_g, off_od, off_out, off_seg, off_perm, off_nlab, off_hlist, off_xch, off_red, off_redi,
      off_bar;
  int e_stage_bytes, o_stage_bytes;
  int total;
};

inline Layout make_layout(int C, int max_pitch) {
  Layout L;
  int off = 0;
  auto take = [&](int bytes) {
    int at = off;
    off += (bytes + 127) & ~127;
    return at;
  };
  L.off_ring = take(2 * kRowF * 4);
  L.e_stage_bytes = kBlk * C * 4;
  L.off_e = take(kSE * L.e_stage_bytes);
  L.o_stage_bytes = kBlk * max_pitch * 4;
  // the out-arc scratch of CTA B's set-up (cnt[256] + dst[768] + w[768]) aliases the score stages
  L.off_o = take(std::max(kSO * L.o_stage_bytes, (kMaxNodes + 6 * kMaxNodes) * 4));
  L.off_g = take(kSG * L.o_stage_bytes);
  L.off_od = take(kSO * 16);
  L.off_out = take(kBlk * C * 4);
  L.off_seg = take(C * 4);
  L.off_perm = take(kMaxNodes * 4);
  L.off_nlab = take(kMaxNodes * 4);
  L.off_hlist = take((kMaxHeavy + 1) * 4);
  L.off_xch = take(kHelpers * kBlk * (2 + kMaxHeavy) * 4);
  L.off_red = take(32 * 4);
  L.off_redi = take(8 * 4);
  L.off_bar = take((2 * kSE + 2 * kSO + 2 * kSG + 1) * 8);
  L.total = off;
  return L;
}

struct Params {
  const GraphMeta* meta;
  const uint8_t* sg_flags;
  const int32_t* sg_in_ptr;
  const int32_t* sg_in_src;
  const int32_t* sg_in_label;
  const float* sg_in_w;
  const float* emissions;
  float* saved; // [node_base + f * pitch + u]: alpha_{f+1} for f < M (written by A), beta_{f+1} for f >= M (by B)
  float* boff; // [utterance][2][nblk_cap][4]: the score offset in effect for the rows of each block, per CTA
  float* out_scores; // [B] forwardScore of the lattice (natural log)
  float* zparts; // [2B] forwardScore(emissions): the part each CTA of the pair summed
  int32_t* status; // [B] bit 0: a weight was not finite
  float* grad; // [B][grad_stride] or NULL (loss only)
  long long grad_stride;
  int C;
  int nwn; // node warps in the launch (ceil(max nodes / 32))
  int nblk_cap;
  Layout lay;
};

/* ------------------------------------------------------------------ */
/* PTX helpers (restated for the host emulation of tests/emu)          */
/* ------------------------------------------------------------------ */
#ifdef GTNB_HOST_EMU
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return emu::shared_window(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  emu::mbar_init(bar, count);
}
__device__ __forceinline__ void mbar_init_fence() {}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  emu::mbar_expect_tx(bar, bytes);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  emu::mbar_arrive(bar);
}
__device__ __forceinline__ void mbar_arrive_n(uint32_t bar, int n) {
  emu::mbar_arrive_n(bar, n);
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  emu::mbar_wait(bar, parity);
}
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  emu::mbar_wait(bar, parity);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  emu::bulk_g2s(dst, src, bytes, bar);
}
__device__ __forceinline__ void bar_named(int id, int nthreads) {
  emu::named_barrier(id, nthreads);
}
__device__ __forceinline__ float lds(uint32_t addr) {
  return *emu::shared_ptr<float>(addr);
}
__device__ __forceinline__ void sts(uint32_t addr, float v) {
  *emu::shared_ptr<float>(addr) = v;
}
__device__ __forceinline__ uint32_t lds_u(uint32_t addr) {
  return *emu::shared_ptr<uint32_t>(addr);
}
__device__ __forceinline__ void sts_u(uint32_t addr, uint32_t v) {
  *emu::shared_ptr<uint32_t>(addr) = v;
}
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
  return *emu::shared_ptr<float4>(addr);
}
__device__ __forceinline__ void stg_v4(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
__device__ __forceinline__ float ex2(float x) {
  return exp2f(x);
}
__device__ __forceinline__ float lg2(float x) {
  return log2f(x);
}
__device__ __forceinline__ void cluster_sync_all() {
  emu::cluster_sync();
}
__device__ __forceinline__ void cluster_arrive_only() {
  emu::cluster_arrive();
}
__device__ __forceinline__ void cluster_wait_only() {
  emu::cluster_wait();
}
__device__ __forceinline__ unsigned cluster_rank() {
  return emu::cluster_ctarank();
}
__device__ __forceinline__ void fence_proxy_async() {}
__device__ __forceinline__ void stg(float* p, float v) {
  *p = v;
}
#else
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_init_fence() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
/* n arrivals at once (a node warp arriving on behalf of the helper warps in the phases they sit out) */
__device__ __forceinline__ void mbar_arrive_n(uint32_t bar, int n) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
/* for the warps that wait a whole block ahead of the math: sleep between polls */
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (;;) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(100);
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void bar_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float lds(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_u(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void stg_v4(float* p, float4 v) {
  asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
/* barrier.cluster: arrive has release, wait has acquire semantics at cluster scope -- the saved scores
 * a CTA stored before it are visible to its partner after it */
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
/* the split form for the producer warp, which has nothing to publish and nothing to acquire: it arrives
 * before its copy loop (so that the others' barrier never waits for it) and waits after it */
__device__ __forceinline__ void cluster_arrive_only() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_only() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
/* orders generic-proxy accesses (the partner's stores to global memory, this CTA's reads / writes of a
 * shared-memory stage) before the async-proxy copies issued next */
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async;" ::: "memory");
}
__device__ __forceinline__ void stg(float* p, float v) {
  asm volatile("st.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
#endif

__device__ __forceinline__ bool finite_ok(float x) {
  return fabsf(x) < kHuge;
}

/* logsumexp of three scores in log2 units; absent arcs carry kNeg: ex2(kNeg - m) = 0, and three
 * absent arcs give kNeg + lg2(3) = kNeg (absorbed) */
__device__ __forceinline__ float lse3(float a, float b, float c) {
  const float m = fmaxf(fmaxf(a, b), c);
  const float s = (ex2(a - m) + ex2(b - m)) + ex2(c - m);
  return m + lg2(s);
}

struct NodeState {
  uint32_t so0, so1, so2; // byte offsets of the three neighbours in a chain row
  float w0, w1, w2; // arc weights, log2 units
  uint32_t u4; // this node's byte offset in a row
  uint32_t pc, qc; // chain rows: read / write (swapped every level)
  uint32_t ea; // this node's emission in the current stage row
  uint32_t oa; // this node's slot in the current stage row of the partner's scores
  uint32_t pdelta; // where the posterior goes, relative to oa: (posterior block - score stage) + (label-sorted position - node id) * 4
  int estep, ostep; // +-(C * 4), +-(pitch * 4)
  float* gs; // where this node's score of the current level is saved (phase 1)
  long long gstep; // +-pitch
  float adj; // renormalisation: subtracted from every score at the first level of a block (an integer)
  float D; // sum of the adjustments so far: true score = kept score + D (exact: integers below 2^24)
  float Zh, zsub; // phase 2: posterior = ex2(((own - Zh) + other) - zsub)
  float Zl, D0, Do0; // the pieces of Z: Z = D0 + Do0 + Zh + Zl (log2 units)
  uint32_t redi; // per-warp maxima of the block's last level (ints), 8 slots
  bool act;
  int nact; // threads of the node warps of this CTA
};

/*
 * Renormalisation.  fp32 scores of a T = 1000 utterance reach several thousand, where one ulp is 5e-4:
 * the posteriors exp(alpha + beta - Z) of an un-normalised recursion carry that as a RELATIVE error
 * (the reference's own fp32 gradient is 2e-3 off the exact one there).  Here every block of 8 levels
 * ends with the integer part of the largest kept score being published (one F2I, one REDUX.MAX, one STS
 * by lane 0 -- no extra barrier: the level's own barrier orders it), and the next block's first level
 * subtracts it from every score.  Integers subtract exactly and their sum D stays exact in fp32, so the
 * kept scores stay within a few hundred of zero (ulp ~ 3e-5) and nothing is lost: true = kept + D.
 */
__device__ __forceinline__ void publish_block_max(const NodeState& st, float stored) {
  const int iv = st.act ? __float2int_rn(fmaxf(stored, -2.0e9f)) : kIntMin;
  const int wm = __reduce_max_sync(0xffffffffu, iv);
  if ((threadIdx.x & 31) == 0) sts_u(st.redi + 4u * (threadIdx.x >> 5), (uint32_t)wm);
}
/* at the start of the next block: the adjustment every node thread applies (same value in all of them) */
__device__ __forceinline__ void fetch_block_adjust(NodeState& st, int nwarps) {
  int mx = kIntMin;
  for (int w = 0; w < nwarps; w++) mx = max(mx, (int)lds_u(st.redi + 4u * w));
  // nothing alive (or nothing published yet): no adjustment
  const float off = (mx <= -2000000000) ? 0.0f : (float)mx;
  st.adj = off;
  st.D += off;
}

/* one level.  DIR 0: CTA A (alpha), 1: CTA B (beta).  PH 1: save the score; 2: posterior in place.
 * LAST: the block's last level (publishes the maximum for the renormalisation) */
template <int DIR, int PH>
__device__ __forceinline__ void node_step(NodeState& st, bool last) {
  const float x = lds(st.ea);
  st.ea += st.estep;
  const float v = lse3(lds(st.pc + st.so0) + st.w0, lds(st.pc + st.so1) + st.w1, lds(st.pc + st.so2) + st.w2);
  // A: alpha_s = v + e[s-1][label(u)].  B: beta_s = v; what the predecessors read is beta_s + e[s-1][label(u)]
  const float val = DIR == 0 ? v + fmaf(x, kLog2e, -st.adj) : v - st.adj;
  const float stored = DIR == 0 ? val : fmaf(x, kLog2e, val);
  st.adj = 0.0f;
  sts(st.qc + st.u4, stored);
  if (PH == 2) {
    // the posterior, written at the node's label-sorted position of the helpers' block
    sts(st.oa + st.pdelta, ex2(((val - st.Zh) + lds(st.oa)) - st.zsub));
    st.oa += st.ostep;
  }
  if (last) publish_block_max(st, stored);
  bar_named(1, st.nact);
  if (PH == 1) { // after the barrier: nothing waits for it
    if (st.act) stg(st.gs, val);
    st.gs += st.gstep;
  }
  const uint32_t t = st.pc;
  st.pc = st.qc;
  st.qc = t;
}

template <int DIR, int PH>
__device__ __forceinline__ void node_rows(NodeState& st, int n) {
  if (n == kBlk) {
#pragma unroll
    for (int r = 0; r < kBlk; r++) node_step<DIR, PH>(st, r == kBlk - 1);
  } else {
#pragma unroll 1
    for (int r = 0; r < n; r++) node_step<DIR, PH>(st, r == n - 1);
  }
}

/* reduction over the node threads of the CTA (named barrier 1): max, then sum */
__device__ __forceinline__ float node_reduce_max(float v, uint32_t red, int warp, int nwarps, int nact) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) sts(red + 4u * warp, v);
  bar_named(1, nact);
  float r = lds(red);
  for (int w = 1; w < nwarps; w++) r = fmaxf(r, lds(red + 4u * w));
  bar_named(1, nact);
  return r;
}
__device__ __forceinline__ float node_reduce_sum(float v, uint32_t red, int warp, int nwarps, int nact) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sts(red + 4u * warp, v);
  bar_named(1, nact);
  float r = lds(red);
  for (int w = 1; w < nwarps; w++) r += lds(red + 4u * w);
  bar_named(1, nact);
  return r;
}

/*
 * The first level of phase 2: the partner's row of the same level is there, so
 *   Z = logsumexp_u(alpha_s[u] + beta_s[u])   (log2 units; + both CTAs' offsets)
 * is reduced over the node threads -- as the pair (Zh, Zl) = (max of the rounded sums, lg2 of the sum of
 * 2^(exact sum - Zh)): the rounding error of every alpha + beta is carried into Zl by a TwoSum -- and the
 * level's posterior follows.  An utterance without any accepting path (Z = kNeg-like) gets Zh = +1e30,
 * which makes every posterior ex2(-huge) = 0.  `doth`: the partner's offset for this block.
 */
template <int DIR>
__device__ __forceinline__ void node_first_phase2(
    NodeState& st, uint32_t red, int warp, int nwarps, float doth, bool last, float* z_out) {
  const float x = lds(st.ea);
  st.ea += st.estep;
  const float v = lse3(lds(st.pc + st.so0) + st.w0, lds(st.pc + st.so1) + st.w1, lds(st.pc + st.so2) + st.w2);
  const float val = DIR == 0 ? v + fmaf(x, kLog2e, -st.adj) : v - st.adj;
  const float stored = DIR == 0 ? val : fmaf(x, kLog2e, val);
  st.adj = 0.0f;
  sts(st.qc + st.u4, stored);
  const float o = st.act ? lds(st.oa) : kNeg;
  // TwoSum (Knuth): hi + lo == val + o exactly
  const float hi = val + o;
  const float bb = hi - val;
  const float lo = (val - (hi - bb)) + (o - bb);
  const float hv = st.act ? hi : 2.0f * kNeg;
  const float m = node_reduce_max(hv, red, warp, nwarps, st.nact);
  const float term = (st.act && m > 1.5f * kNeg) ? ex2((hi - m) + lo) : 0.0f;
  const float S = node_reduce_sum(term, red, warp, nwarps, st.nact);
  const bool feasible = m > kNeg * 0.5f && S > 0.0f;
  st.Zh = feasible ? m : 1.0e30f;
  st.Zl = feasible ? lg2(S) : 0.0f;
  st.D0 = st.D;
  st.Do0 = doth;
  st.zsub = st.Zl;
  if (z_out && threadIdx.x == 0)
    *z_out = feasible ? (float)((((double)st.D0 + (double)doth) + ((double)m + (double)st.Zl)) * 0.6931471805599453)
                      : -CUDART_INF_F;
  sts(st.oa + st.pdelta, ex2(((val - st.Zh) + o) - st.zsub));
  st.oa += st.ostep;
  if (last) publish_block_max(st, stored);
  bar_named(1, st.nact);
  const uint32_t t = st.pc;
  st.pc = st.qc;
  st.qc = t;
}

/* ------------------------------------------------------------------ */
/* helper warps                                                        */
/* ------------------------------------------------------------------ */

struct HelperState {
  int hw, lane, r, part; // helper warp, lane, row of the block, label chunk
  int C, pitch;
  uint32_t seg, hlist, xch, out; // shared-window addresses of the tables
  float zacc; // lanes 0..7 of helper warp 0: sum of the rows' logsumexp
  bool bad;
};

/* seg[c] = first position (10 bits) | number of nodes (10 bits) | heavy slot + 1 (above): the posteriors
 * of the nodes that carry label c are CONTIGUOUS in a stage row, because the node threads write theirs at
 * the node's label-sorted position */
__device__ __forceinline__ uint32_t seg_pack(int start, int len, int slot1) {
  return (uint32_t)start | ((uint32_t)len << 10) | ((uint32_t)slot1 << 20);
}

/* One block of `nfr` frames (rows 0 .. nfr-1 of the stages).  WANT_G: the posteriors of the block are in
 * the score stage `ostage`; write the gradient rows to `gout` (global, nfr * C floats).  Otherwise only
 * the rows' logsumexp is accumulated.  The emission stage is handed back (e_empty) as soon as its rows
 * are in registers. */
template <int CH, bool WANT_G>
__device__ __forceinline__ void helper_block(
    HelperState& hs, int nfr, uint32_t estage, uint32_t e_empty_bar, uint32_t ostage, uint32_t g_empty_bar, int nheavy,
    float* gout) {
  const int C = hs.C, r = hs.r;
  const bool row_on = r < nfr;
  const int c0 = CH * hs.part;
  float x[CH];
  float mx = -3.0e38f;
#pragma unroll
  for (int i = 0; i < CH; i++) {
    const int c = c0 + ((i + r) % CH);
    const bool ok = row_on && c < C;
    x[i] = ok ? lds(estage + (uint32_t)(r * C + c) * 4u) : -3.0e38f;
    if (ok) hs.bad |= !finite_ok(x[i]);
    mx = fmaxf(mx, x[i]);
  }
  __syncwarp();
  if (hs.lane == 0) mbar_arrive(e_empty_bar);
  // the 4 parts of a row sit in lanes r, r + 8, r + 16, r + 24
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
  float sm = 0.0f;
#pragma unroll
  for (int i = 0; i < CH; i++) {
    x[i] = ex2((x[i] - mx) * kLog2e); // invalid entries: ex2(-huge) = 0
    sm += x[i];
  }
  sm += __shfl_xor_sync(0xffffffffu, sm, 8);
  sm += __shfl_xor_sync(0xffffffffu, sm, 16);
  const uint32_t orow = ostage + (uint32_t)(r * hs.pitch) * 4u;
  // labels carried by many nodes (CTC: blank): this warp's share of the segment, summed over its 4 parts
  float hp[kMaxHeavy];
#pragma unroll
  for (int h = 0; h < kMaxHeavy; h++) {
    hp[h] = 0.0f;
    if (WANT_G && h < nheavy) {
      const uint32_t w = lds_u(hs.seg + 4u * lds_u(hs.hlist + 4u * h));
      const int start = w & 1023, len = (w >> 10) & 1023;
      float a0 = 0.0f, a1 = 0.0f;
      if (row_on) {
        int k = hs.part;
        for (; k + 4 * kHelpers < len; k += 8 * kHelpers) {
          a0 += lds(orow + (uint32_t)(start + k) * 4u);
          a1 += lds(orow + (uint32_t)(start + k + 4 * kHelpers) * 4u);
        }
        if (k < len) a0 += lds(orow + (uint32_t)(start + k) * 4u);
      }
      float acc = a0 + a1;
      acc += __shfl_xor_sync(0xffffffffu, acc, 8);
      acc += __shfl_xor_sync(0xffffffffu, acc, 16);
      hp[h] = acc;
    }
  }
  // exchange with the other helper warp: (max, sum) of its half of the row, heavy partial sums
  const uint32_t mine = hs.xch + (uint32_t)((hs.hw * kBlk + r) * (2 + kMaxHeavy)) * 4u;
  const uint32_t other = hs.xch + (uint32_t)(((1 - hs.hw) * kBlk + r) * (2 + kMaxHeavy)) * 4u;
  if (hs.lane < kBlk) {
    sts(mine, mx);
    sts(mine + 4, sm);
    if (WANT_G) {
#pragma unroll
      for (int h = 0; h < kMaxHeavy; h++) sts(mine + 8 + 4 * h, hp[h]);
    }
  }
  bar_named(2, 32 * kHelpers);
  const float omx = lds(other), osm = lds(other + 4);
  const float m = fmaxf(mx, omx);
  const float sc_mine = ex2((mx - m) * kLog2e);
  const float S = sm * sc_mine + osm * ex2((omx - m) * kLog2e);
  if (row_on && hs.hw == 0 && hs.lane < kBlk) hs.zacc += m + lg2(S) * kLn2; // logsumexp of the row
  if (WANT_G) {
    const float scale = sc_mine / S;
#pragma unroll
    for (int h = 0; h < kMaxHeavy; h++)
      if (h < nheavy) hp[h] += lds(other + 8 + 4 * h);
#pragma unroll
    for (int i = 0; i < CH; i++) {
      const int c = c0 + ((i + r) % CH);
      if (row_on && c < C) {
        const uint32_t w = lds_u(hs.seg + 4u * c);
        const int len = (w >> 10) & 1023, slot1 = w >> 20;
        const uint32_t a = orow + (w & 1023) * 4u;
        float occ;
        if (slot1) {
          occ = hp[0];
#pragma unroll
          for (int h = 1; h < kMaxHeavy; h++)
            if (h == slot1 - 1) occ = hp[h];
        } else {
          const float g0 = len > 0 ? lds(a) : 0.0f;
          const float g1 = len > 1 ? lds(a + 4) : 0.0f;
          const float g2 = len > 2 ? lds(a + 8) : 0.0f;
          const float g3 = len > 3 ? lds(a + 12) : 0.0f;
          occ = (g0 + g1) + (g2 + g3);
          for (int k = 4; k < len; k++) occ += lds(a + 4u * k);
        }
        // d(forwardScore(e) - forwardScore(lattice)) / d e[f][c] = softmax - posterior mass on label c
        sts(hs.out + (uint32_t)(r * C + c) * 4u, x[i] * scale - occ);
      }
    }
    __syncwarp();
    if (hs.lane == 0) mbar_arrive(g_empty_bar); // this warp is done with the posterior block
    bar_named(2, 32 * kHelpers);
    // finished rows -> global, 16 bytes per lane
    const int tid64 = hs.hw * 32 + hs.lane;
    const int n4 = (nfr * C) >> 2;
    for (int k = tid64; k < n4; k += 32 * kHelpers) stg_v4(gout + 4 * k, lds_v4(hs.out + 16u * k));
  } else {
    bar_named(2, 32 * kHelpers); // xch is rewritten by the next block
  }
}

/* ------------------------------------------------------------------ */
/* the kernel                                                          */
/* ------------------------------------------------------------------ */

template <int CH>
__global__ void __launch_bounds__(32 * (8 + 1 + kHelpers), 4) bidir_ctc_kernel(const Params P) {
  GTNB_DYNAMIC_SMEM_128(unsigned char, smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x >> 1;
  const int dir = (int)cluster_rank(); // 0: CTA A (alpha), 1: CTA B (beta)
  const GraphMeta m = P.meta[b];
  const int N1 = m.sg_N, T = m.T, C = P.C;
  const int pitch = (N1 + 3) & ~3;
  const bool want_g = P.grad != nullptr;
  const Layout& L = P.lay;
  float* ring = reinterpret_cast<float*>(smem + L.off_ring);
  const uint32_t bars = smem_u32(smem + L.off_bar);
  const uint32_t e_base = smem_u32(smem + L.off_e), o_base = smem_u32(smem + L.off_o);
  const uint32_t od_base = smem_u32(smem + L.off_od);
  const uint32_t seg_a = smem_u32(smem + L.off_seg), perm_a = smem_u32(smem + L.off_perm);
  const uint32_t nlab_a = smem_u32(smem + L.off_nlab), hlist_a = smem_u32(smem + L.off_hlist);
  const uint32_t red_a = smem_u32(smem + L.off_red), redi_a = smem_u32(smem + L.off_redi);
  auto e_full = [&](int s) { return bars + 8u * (uint32_t)s; };
  auto e_empty = [&](int s) { return bars + 8u * (uint32_t)(kSE + s); };
  auto o_full = [&](int s) { return bars + 8u * (uint32_t)(2 * kSE + s); };
  auto o_empty = [&](int s) { return bars + 8u * (uint32_t)(2 * kSE + kSO + s); };
  auto g_full = [&](int s) { return bars + 8u * (uint32_t)(2 * kSE + 2 * kSO + s); };
  auto g_empty = [&](int s) { return bars + 8u * (uint32_t)(2 * kSE + 2 * kSO + kSG + s); };
  const uint32_t tbl_ready = bars + 8u * (uint32_t)(2 * kSE + 2 * kSO + 2 * kSG);
  const uint32_t g_base = smem_u32(smem + L.off_g);

  const int nw_act = (N1 + 31) >> 5; // node warps of THIS utterance
  const int nact = 32 * nw_act;
  const int nblk = (T + kBlk - 1) / kBlk;
  const int M = ((T / 2) / kBlk) * kBlk; // levels 1..M: alpha saved, posterior by B; M+1..T: beta saved, posterior by A
  const int jb = M / kBlk; // blocks [0, jb) belong to B's phase 2, [jb, nblk) to A's
  // blocks this CTA walks, in its own order: A 0, 1, ..; B nblk-1, nblk-2, ..
  const int n_ph1 = dir == 0 ? jb : nblk - jb;
  int n_ph2 = dir == 0 ? nblk - jb : jb;
  // loss only: A needs one more level (for Z), B nothing more
  if (!want_g) n_ph2 = dir == 0 ? (T > 0 ? 1 : 0) : 0;
  const int n_visit = n_ph1 + n_ph2;
  auto block_of = [&](int v) { return dir == 0 ? v : nblk - 1 - v; };
  // the helper warps read the emission blocks of ONE phase (2 when there is a gradient, else 1); for the
  // other blocks node warp 0 arrives on the stage's "empty" barrier on their behalf
  auto helpers_read = [&](int v) { return want_g ? v >= n_ph1 : v < n_ph1; };
  auto release_e = [&](int v) { // lane 0 of a node warp, after the block's last level
    mbar_arrive_n(e_empty(v % kSE), (warp == 0 && !helpers_read(v)) ? 1 + kHelpers : 1);
  };
  auto rows_of = [&](int j) { return min(kBlk, T - kBlk * j); };
  float* boff_own = P.boff + ((long long)(2 * b + dir) * P.nblk_cap) * 4;
  const float* boff_oth = P.boff + ((long long)(2 * b + (1 - dir)) * P.nblk_cap) * 4;

  if (tid == 0) {
    for (int s = 0; s < kSE; s++) {
      mbar_init(e_full(s), 1);
      mbar_init(e_empty(s), nw_act + kHelpers);
    }
    for (int s = 0; s < kSO; s++) {
      mbar_init(o_full(s), 1);
      mbar_init(o_empty(s), nw_act);
    }
    for (int s = 0; s < kSG; s++) {
      mbar_init(g_full(s), nw_act);
      mbar_init(g_empty(s), kHelpers);
    }
    mbar_init(tbl_ready, kHelpers);
    mbar_init_fence();
  }
  // chain rows: everything kNeg (the dummy slot stays that way for good)
  for (int i = tid; i < 2 * kRowF; i += blockDim.x) ring[i] = kNeg;
  if (tid < 8) sts_u(redi_a + 4u * tid, (uint32_t)kIntMin);

  const uint8_t* fl = P.sg_flags + m.sg_node_base;
  const int32_t* ip = P.sg_in_ptr + m.sg_node_base;
  const int32_t* is = P.sg_in_src + m.sg_arc_base;
  const int32_t* il = P.sg_in_label + m.sg_arc_base;
  const float* iw = P.sg_in_w + m.sg_arc_base;
  const float* em = P.emissions + m.emis_off;
  float* saved = P.saved + m.node_base;

  // B: the out-arcs of every node, gathered from the in-arc tables (scratch aliases the score stages)
  int* o_cnt = reinterpret_cast<int*>(smem + L.off_o);
  int* o_dst = o_cnt + kMaxNodes;
  float* o_w = reinterpret_cast<float*>(o_dst + 3 * kMaxNodes);
  if (dir == 1)
    for (int i = tid; i < kMaxNodes; i += blockDim.x) o_cnt[i] = 0;
  __syncthreads();
  bool bad = false;
  int my_lab = 0, my_deg = 0;
  if (tid < N1) {
    const int e0 = ip[tid];
    my_deg = ip[tid + 1] - e0;
    if (my_deg > 0) my_lab = il[e0];
    // node labels for the helper warps' tables
    sts_u(nlab_a + 4u * tid, (uint32_t)(my_deg > 0 ? my_lab : -1));
    if (dir == 1)
      for (int k = 0; k < my_deg; k++) {
        const int s = is[e0 + k];
        const int slot = atomicAdd(&o_cnt[s], 1);
        if (slot < 3) {
          o_dst[3 * s + slot] = tid;
          o_w[3 * s + slot] = iw[e0 + k];
        }
      }
  }
  __syncthreads();

  if (warp < P.nwn) {
    /* ============================ node warps ============================ */
    NodeState st;
    st.act = tid < N1;
    st.nact = nact;
    st.u4 = 4u * (uint32_t)tid;
    st.so0 = st.so1 = st.so2 = 4u * kDummy;
    st.w0 = st.w1 = st.w2 = 0.0f;
    if (st.act) {
      int nb_[3] = {kDummy, kDummy, kDummy};
      float wb_[3] = {0.0f, 0.0f, 0.0f};
      if (dir == 0) {
        const int e0 = ip[tid];
        for (int k = 0; k < my_deg && k < 3; k++) {
          nb_[k] = is[e0 + k];
          wb_[k] = iw[e0 + k];
        }
      } else {
        const int od = min(o_cnt[tid], 3);
        for (int k = 0; k < od; k++) {
          nb_[k] = o_dst[3 * tid + k];
          wb_[k] = o_w[3 * tid + k];
        }
        // the atomics filled the slots in arbitrary order: sort by (node, weight) so that the sum of
        // the three exponentials is evaluated in a fixed order
#define GTNB_CSWAP(a, b)                                                          \
  if (nb_[a] > nb_[b] || (nb_[a] == nb_[b] && wb_[a] > wb_[b])) {                \
    const int tn = nb_[a];                                                        \
    nb_[a] = nb_[b];                                                              \
    nb_[b] = tn;                                                                  \
    const float tw = wb_[a];                                                      \
    wb_[a] = wb_[b];                                                              \
    wb_[b] = tw;                                                                  \
  }
        GTNB_CSWAP(0, 1) GTNB_CSWAP(1, 2) GTNB_CSWAP(0, 1)
#undef GTNB_CSWAP
      }
      bad |= !finite_ok(wb_[0]) || !finite_ok(wb_[1]) || !finite_ok(wb_[2]);
      st.so0 = 4u * (uint32_t)nb_[0];
      st.so1 = 4u * (uint32_t)nb_[1];
      st.so2 = 4u * (uint32_t)nb_[2];
      st.w0 = wb_[0] * kLog2e;
      st.w1 = wb_[1] * kLog2e;
      st.w2 = wb_[2] * kLog2e;
    }
    // the scratch is about to be overwritten by the score stages: everybody is done reading it
    __syncthreads(); // (S1) all roles
    if (warp >= nw_act) { // node warp without a node of this utterance
      cluster_sync_all();
      return;
    }
    st.pc = smem_u32(ring);
    st.qc = smem_u32(ring + kRowF);
    st.estep = dir == 0 ? 4 * C : -4 * C;
    // lanes of the last node warp that hold no node must not write a posterior into the stage rows:
    // they aim at a spare slot of the chain row instead (index kDummy + 1, read by nobody)
    st.ostep = !st.act ? 0 : dir == 0 ? 4 * pitch : -4 * pitch;
    const uint32_t spare = smem_u32(ring) + 4u * (kDummy + 1);
    st.gstep = dir == 0 ? pitch : -pitch;
    st.Zh = st.zsub = st.Zl = st.D0 = st.Do0 = 0.0f;
    st.adj = st.D = 0.0f;
    st.pdelta = 0;
    st.redi = redi_a;
    const uint32_t lab4 = 4u * (uint32_t)my_lab;

    int v = 0; // blocks visited
    if (dir == 0) {
      // level 0: start nodes carry the implicit 0 (shortest.cpp:129-135)
      if (st.act && (fl[tid] & 1)) sts(st.pc + st.u4, 0.0f);
      bar_named(1, nact);
      st.gs = saved + tid;
      for (; v < n_ph1; v++) {
        mbar_wait(e_full(v % kSE), (v / kSE) & 1);
        st.ea = e_base + (uint32_t)((v % kSE) * L.e_stage_bytes) + lab4;
        if (v > 0) fetch_block_adjust(st, nw_act);
        if (tid == 0) stg(boff_own + 4 * v, st.D);
        node_rows<0, 1>(st, kBlk); // A's phase-1 blocks are always full
        __syncwarp();
        if (lane == 0) release_e(v);
      }
    } else if (T > 0) {
      // level T: accept nodes carry 0; what the predecessors read is beta_T + e[T-1][label]
      const int j0 = nblk - 1, n0 = rows_of(j0);
      mbar_wait(e_full(0), 0);
      st.ea = e_base + (uint32_t)((n0 - 1) * C * 4) + lab4;
      if (tid == 0) stg(boff_own + 4 * j0, 0.0f);
      {
        const float e2 = lds(st.ea) * kLog2e;
        st.ea += st.estep;
        const float val = (st.act && (fl[tid] & 2)) ? 0.0f : kNeg;
        sts(st.pc + st.u4, val + e2); // written to the row the first step READS
        st.gs = saved + (long long)(T - 1) * pitch + tid;
        if (n0 == 1) publish_block_max(st, val + e2);
        bar_named(1, nact);
        if (st.act) stg(st.gs, val);
        st.gs += st.gstep;
      }
      node_rows<1, 1>(st, n0 - 1);
      __syncwarp();
      if (lane == 0) release_e(0);
      for (v = 1; v < n_ph1; v++) {
        mbar_wait(e_full(v % kSE), (v / kSE) & 1);
        st.ea = e_base + (uint32_t)((v % kSE) * L.e_stage_bytes + (kBlk - 1) * C * 4) + lab4;
        fetch_block_adjust(st, nw_act);
        if (tid == 0) stg(boff_own + 4 * block_of(v), st.D);
        node_rows<1, 1>(st, kBlk);
        __syncwarp();
        if (lane == 0) release_e(v);
      }
    }
    cluster_sync_all();
    // ---- phase 2
    uint32_t perm_rel = 0; // (label-sorted position - node id) * 4
    if (n_ph2 > 0 && want_g) {
      mbar_wait(tbl_ready, 0); // the helper warps' label-sorted positions
      if (st.act) perm_rel = lds_u(perm_a + st.u4) - st.u4;
    }
    int v2 = 0;
    for (; v2 < n_ph2; v2++, v++) {
      const int j = block_of(v), nfr = rows_of(j);
      const int so = v2 % kSO, sg = v2 % kSG;
      mbar_wait(e_full(v % kSE), (v / kSE) & 1);
      mbar_wait(o_full(so), (v2 / kSO) & 1);
      if (v2 >= kSG) mbar_wait(g_empty(sg), ((v2 / kSG) - 1) & 1);
      const int r0 = dir == 0 ? 0 : nfr - 1;
      st.ea = e_base + (uint32_t)((v % kSE) * L.e_stage_bytes + r0 * C * 4) + lab4;
      st.oa = st.act ? o_base + (uint32_t)(so * L.o_stage_bytes + r0 * pitch * 4) + st.u4 : spare;
      st.pdelta = st.act ? (g_base + (uint32_t)(sg * L.o_stage_bytes)) - (o_base + (uint32_t)(so * L.o_stage_bytes)) + perm_rel
                         : 0u;
      if (v > 0) fetch_block_adjust(st, nw_act);
      const float doth = lds(od_base + 16u * so);
      if (v2 == 0) {
        if (dir == 0) {
          node_first_phase2<0>(st, red_a, warp, nw_act, doth, nfr == 1, P.out_scores + b);
          if (want_g) node_rows<0, 2>(st, nfr - 1);
        } else {
          node_first_phase2<1>(st, red_a, warp, nw_act, doth, nfr == 1, nullptr);
          node_rows<1, 2>(st, nfr - 1);
        }
      } else {
        // offsets moved since the meeting level: all integers, the differences are exact
        st.zsub = st.Zl - ((st.D - st.D0) + (doth - st.Do0));
        if (dir == 0)
          node_rows<0, 2>(st, nfr);
        else
          node_rows<1, 2>(st, nfr);
      }
      __syncwarp();
      if (lane == 0) {
        release_e(v);
        mbar_arrive(o_empty(so));
        mbar_arrive(g_full(sg));
      }
    }
    if (bad) atomicOr(&P.status[b], 1);
    if (T == 0 && dir == 0 && tid == 0) {
      // no frames: the lattice is the graph's start-and-accept nodes, each with score 0
      int n = 0;
      for (int i = 0; i < N1; i++) n += (fl[i] & 3) == 3;
      P.out_scores[b] = n ? logf((float)n) : -CUDART_INF_F;
    }
    return;
  }

  __syncthreads(); // (S1)

  if (warp == P.nwn) {
    /* ============================ producer warp: TMA bulk copies ============================ */
    // Emission blocks all the way, and in phase 2 the partner's saved scores (with the block's offset).
    // The emission copies run ahead of the math across the phase boundary, so the warp must not be what
    // the others' cluster barrier waits for: it ARRIVES right away and WAITS only before the first copy
    // that reads what the partner wrote.
    cluster_arrive_only();
    bool waited = false;
    for (int v = 0; v < n_visit; v++) {
      const int s = v % kSE, j = block_of(v);
      if (v >= kSE) mbar_wait_relaxed(e_empty(s), ((v / kSE) - 1) & 1);
      if (lane == 0) {
        const uint32_t bytes = (uint32_t)(rows_of(j) * C * 4);
        mbar_expect_tx(e_full(s), bytes);
        bulk_g2s(e_base + (uint32_t)(s * L.e_stage_bytes), em + (long long)j * kBlk * C, bytes, e_full(s));
      }
      if (v >= n_ph1) {
        const int v2 = v - n_ph1, so = v2 % kSO;
        if (!waited) {
          __syncwarp();
          cluster_wait_only();
          // the partner's saved scores were written with ordinary stores; the bulk copies read them
          // through the async proxy
          fence_proxy_async();
          waited = true;
        }
        if (v2 >= kSO) mbar_wait_relaxed(o_empty(so), ((v2 / kSO) - 1) & 1);
        if (lane == 0) {
          const uint32_t bytes = (uint32_t)(rows_of(j) * pitch * 4);
          mbar_expect_tx(o_full(so), bytes + 16u);
          bulk_g2s(o_base + (uint32_t)(so * L.o_stage_bytes), saved + (long long)j * kBlk * pitch, bytes, o_full(so));
          bulk_g2s(od_base + 16u * so, boff_oth + 4 * j, 16u, o_full(so));
        }
      }
    }
    __syncwarp();
    if (!waited) cluster_wait_only();
    return;
  }

  /* ============================ helper warps ============================ */
  HelperState hs;
  hs.hw = warp - P.nwn - 1;
  hs.lane = lane;
  hs.r = lane & 7;
  hs.part = (lane >> 3) + 4 * hs.hw;
  hs.C = C;
  hs.pitch = pitch;
  hs.seg = seg_a;
  hs.hlist = hlist_a;
  hs.xch = smem_u32(smem + L.off_xch);
  hs.out = smem_u32(smem + L.off_out);
  hs.zacc = 0.0f;
  hs.bad = false;
  const int tid64 = hs.hw * 32 + lane;
  int nheavy = 0;
  if (want_g) {
    // Label-sorted positions: the nodes that carry label c get the consecutive positions
    // [start_c, start_c + n_c), ascending node ids inside (a fixed summation order, no atomics); nodes
    // without in-arcs (they exist at level 0 only) go behind all of them.
    for (int c = tid64; c < C; c += 32 * kHelpers) {
      int n = 0;
      for (int u = 0; u < N1; u++) n += (int)lds_u(nlab_a + 4u * u) == c;
      sts_u(seg_a + 4u * c, (uint32_t)n);
    }
    bar_named(2, 32 * kHelpers);
    if (tid64 == 0) {
      int nh = 0, at = 0;
      for (int c = 0; c < C; c++) {
        const int n = (int)lds_u(seg_a + 4u * c);
        int slot1 = 0;
        if (n > kHeavyLen && nh < kMaxHeavy) {
          sts_u(hlist_a + 4u * nh, (uint32_t)c);
          slot1 = ++nh;
        }
        sts_u(seg_a + 4u * c, seg_pack(at, n, slot1));
        at += n;
      }
      sts_u(hlist_a + 4u * kMaxHeavy, (uint32_t)nh);
      for (int u = 0; u < N1; u++)
        if ((int)lds_u(nlab_a + 4u * u) < 0) sts_u(perm_a + 4u * u, 4u * (uint32_t)(at++));
    }
    bar_named(2, 32 * kHelpers);
    nheavy = (int)lds_u(hlist_a + 4u * kMaxHeavy);
    for (int c = tid64; c < C; c += 32 * kHelpers) {
      uint32_t at = lds_u(seg_a + 4u * c) & 1023u;
      for (int u = 0; u < N1; u++)
        if ((int)lds_u(nlab_a + 4u * u) == c) sts_u(perm_a + 4u * u, 4u * (at++));
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(tbl_ready);
  }
  int v = 0;
  if (!want_g) {
    for (; v < n_ph1; v++) {
      const int s = v % kSE;
      mbar_wait_relaxed(e_full(s), (v / kSE) & 1);
      helper_block<CH, false>(hs, rows_of(block_of(v)), e_base + (uint32_t)(s * L.e_stage_bytes), e_empty(s), 0, 0, 0,
                              nullptr);
    }
  }
  cluster_sync_all();
  if (want_g) {
    v = n_ph1;
    for (int v2 = 0; v2 < n_ph2; v2++, v++) {
      const int s = v % kSE, sg = v2 % kSG, j = block_of(v);
      mbar_wait_relaxed(e_full(s), (v / kSE) & 1);
      mbar_wait_relaxed(g_full(sg), (v2 / kSG) & 1);
      helper_block<CH, true>(hs, rows_of(j), e_base + (uint32_t)(s * L.e_stage_bytes), e_empty(s),
                             g_base + (uint32_t)(sg * L.o_stage_bytes), g_empty(sg), nheavy,
                             P.grad + (long long)b * P.grad_stride + (long long)j * kBlk * C);
    }
  }
  // forwardScore(emissions): the rows this CTA summed
  float z = (hs.hw == 0 && lane < kBlk) ? hs.zacc : 0.0f;
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
  if (tid64 == 0) P.zparts[2 * b + dir] = z;
  if (hs.bad) atomicOr(&P.status[b], 1);
}

} // namespace bidir

#ifndef GTNB_HOST_EMU

/* true when every utterance of the batch qualifies for the bidirectional kernel */
bool bidir_supported(const gtnb_lattice* lat, const float* emissions, int64_t stride, const float* grad, int64_t grad_stride) {
  if (!lat->composed || lat->C % 4 != 0 || lat->C > 128 || lat->C < 4) return false;
  if (lat->max_lvl_nodes > bidir::kMaxNodes || lat->max_in_deg > 3 || lat->max_out_deg > 3) return false;
  if (((uintptr_t)emissions & 15) || (stride & 3)) return false;
  if (grad && (((uintptr_t)grad & 15) || (grad_stride & 3))) return false;
  for (int b = 0; b < lat->B; b++)
    if (!lat->meta_h[b].sg_uniform || !lat->meta_h[b].sg_all_valid || lat->meta_h[b].sg_N < 1) return false;
  return true;
}

/* blocks of 8 frames per utterance: boff_dev holds B * 2 * bidir_blocks(max T) * 4 floats */
int bidir_blocks(int max_T) {
  return std::max(1, (max_T + bidir::kBlk - 1) / bidir::kBlk);
}

/* utterances [b0, b0 + nb): one cluster of two CTAs each, on ctx->stream.  zparts_dev: [2 * B] */
int launch_bidir_ctc(
    gtnb_ctx* ctx, gtnb_lattice* lat, int32_t* status_dev, float* zparts_dev, float* boff_dev, float* grad_emis,
    int64_t grad_stride, int b0, int nb) {
  if (nb < 0) nb = lat->B - b0;
  if (nb <= 0) return GTNB_OK;
  bidir::Params P;
  P.meta = lat->meta + b0;
  P.sg_flags = lat->sg_flags;
  P.sg_in_ptr = lat->sg_in_ptr;
  P.sg_in_src = lat->sg_in_src;
  P.sg_in_label = lat->sg_in_label;
  P.sg_in_w = lat->sg_in_w;
  P.emissions = lat->emissions;
  P.saved = lat->scores;
  P.nblk_cap = bidir_blocks(lat->max_T);
  P.boff = boff_dev + (long long)b0 * 2 * P.nblk_cap * 4;
  P.out_scores = lat->out_scores + b0;
  P.zparts = zparts_dev + 2ll * b0;
  P.status = status_dev + b0;
  P.grad = grad_emis ? grad_emis + (long long)b0 * grad_stride : nullptr;
  P.grad_stride = grad_stride;
  P.C = lat->C;
  P.nwn = std::max(1, (lat->max_lvl_nodes + 31) / 32);
  const int max_pitch = (lat->max_lvl_nodes + 3) & ~3;
  P.lay = bidir::make_layout(lat->C, max_pitch);
  const int CH = (lat->C + 7) / 8;
  void (*kern)(const bidir::Params) = CH <= 4 ? bidir::bidir_ctc_kernel<4>
                                      : CH <= 8 ? bidir::bidir_ctc_kernel<8>
                                                : bidir::bidir_ctc_kernel<16>;
  if (P.lay.total > 48 * 1024) {
    int rc = ensure_max_smem(ctx, (const void*)kern);
    if (rc) return rc;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * nb, 1, 1);
  cfg.blockDim = dim3(32 * (P.nwn + 1 + bidir::kHelpers), 1, 1);
  cfg.dynamicSmemBytes = P.lay.total;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  GTNB_LAUNCH(ctx, "bidir_ctc", GTNB_CUDA(ctx, cudaLaunchKernelEx(&cfg, kern, P)));
  return GTNB_OK;
}

#endif // GTNB_HOST_EMU

} // namespace gtnb

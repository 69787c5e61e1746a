/*
 * k_dense.cu -- factored dense-trellis Viterbi: viterbiPath / viterbiScore of
 * compose(emissions, transitions) (test/criterion_test.cpp:328-344)
 * WITHOUT materialising the lattice.  At T=2000, C=128 the composed graph has
 * 32.75 M arcs per utterance (16.8 G for B=512): it cannot exist in HBM, so the
 * lattice stays factored, w(t, j->i) = e[t][i] + trans[j->i], and only the [T][C]
 * back-pointers (1 byte each) and the rolling [C] score vector are stored.
 *
 * Exactness (the path must be bit-identical to the reference's):
 *   - the reference's arc weight is first.weight + second.weight = fl(e + trans)
 *     (compose.cpp:435) and the relaxation is fl(score[src] + w) (shortest.cpp:213):
 *     both adds are kept as separate __fadd_rn, in that association;
 *   - ties: shortestPath keeps the FIRST relaxation that reaches the maximum
 *     (strict '>', shortest.cpp:214).  For this lattice the reference pops the
 *     frame-t nodes in ascending state order, so the smallest predecessor j wins;
 *     the best accept node is the first maximum in ascending state order (:226-237).
 *
 * One CTA per utterance; the C x C transition tile lives in shared memory for the
 * whole utterance (padded so that the G lanes sharing an output state hit distinct
 * banks); G = 256 / C lanes split the predecessor range of one state and combine with
 * warp shuffles; the next frame's emission row is prefetched while the current frame
 * is evaluated; the traceback runs on one warp with 8 back-pointer rows in flight.
 */
#ifdef GTNB_HOST_EMU // this file compiled by g++ against tests/emu/simt_emu.h (CPU test suite)
#include "simt_emu.h"
#else
#include <cuda_runtime.h>
#include <math_constants.h>

#include "gtnb_internal.h"
#endif

namespace gtnb {

namespace {

constexpr int kDenseThreads = 256;

template <int G>
__global__ void __launch_bounds__(kDenseThreads) viterbi_dense_kernel(
    int T_max,
    int C,
    const int32_t* __restrict__ T_arr, // nullable
    const float* __restrict__ emis,
    long long stride,
    const float* __restrict__ trans, // [C] start arcs, then [C][C] as trans[C + i*C + j] = w(j -> i)
    uint8_t* __restrict__ bp, // [B][T_max][C]
    int32_t* __restrict__ paths, // [B][T_max]
    float* __restrict__ scores) {
  GTNB_DYNAMIC_SMEM(unsigned char, smem_raw);
  const int P = C + 32 / G; // padded row stride of the transition tile
  float* s_tr = reinterpret_cast<float*>(smem_raw); // s_tr[j * P + i] = w(j -> i)
  float* s_sc = s_tr + (size_t)C * P; // [2][C]
  const int b = blockIdx.x;
  const int T = T_arr ? T_arr[b] : T_max;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  constexpr int SPW = 32 / G; // states per warp
  const int sub = lane / SPW; // which slice of the predecessors
  const float* e = emis + (long long)b * stride;
  uint8_t* mybp = bp + (long long)b * T_max * C;

  for (int k = tid; k < C * C; k += kDenseThreads) {
    const int i = k / C, j = k % C;
    s_tr[j * P + i] = trans[C + k];
  }
  // frame 0: single predecessor, the start node (score 0): fl(0 + fl(e + w_start))
  for (int i = tid; i < C; i += kDenseThreads) {
    if (T > 0) {
      s_sc[i] = __fadd_rn(0.0f, __fadd_rn(e[i], trans[i]));
      mybp[i] = 0;
    }
  }
  __syncthreads();

  const int n_pass = (C + kDenseThreads / G - 1) / (kDenseThreads / G);
  for (int t = 1; t < T; t++) {
    const float* prev = s_sc + ((t - 1) & 1) * C;
    float* cur = s_sc + (t & 1) * C;
    for (int p = 0; p < n_pass; p++) {
      const int i = p * (kDenseThreads / G) + warp * SPW + (lane % SPW);
      const bool live = i < C;
      const float et = live ? __ldg(&e[(long long)t * C + i]) : 0.0f;
      float best = -CUDART_INF_F;
      int bj = 0x7fffffff;
      if (live) {
        const float* col = s_tr + i;
#pragma unroll 4
        for (int j = sub; j < C; j += G) {
          const float v = __fadd_rn(prev[j], __fadd_rn(et, col[j * P]));
          if (v > best) { // strict: the smallest j of a slice wins ties
            best = v;
            bj = j;
          }
        }
      }
      if (G > 1) {
#pragma unroll
        for (int o = SPW; o < 32; o <<= 1) {
          const float ob = __shfl_xor_sync(0xffffffffu, best, o);
          const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
          if (ob > best || (ob == best && oj < bj)) {
            best = ob;
            bj = oj;
          }
        }
      }
      if (live && sub == 0) {
        cur[i] = best;
        mybp[(long long)t * C + i] = (uint8_t)(bj == 0x7fffffff ? 0 : bj);
      }
    }
    __syncthreads();
  }

  // best accept node: first maximum in ascending state order; then the traceback
  if (warp == 0) {
    if (T == 0) {
      if (lane == 0) scores[b] = 0.0f; // the start node of the chain is also its accept node
      return;
    }
    const float* fin = s_sc + ((T - 1) & 1) * C;
    float best = -CUDART_INF_F;
    int bi = 0x7fffffff;
    for (int i = lane; i < C; i += 32) {
      if (fin[i] > best) {
        best = fin[i];
        bi = i;
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) {
        best = ob;
        bi = oi;
      }
    }
    int32_t* out = paths + (long long)b * T_max;
    if (lane == 0) scores[b] = best;
    if (bi == 0x7fffffff) { // every path is -inf: no accepting path (shortest.cpp:233)
      for (int t = lane; t < T; t += 32) out[t] = -1;
      return;
    }
    // make this CTA's back-pointer stores visible to its own loads
    __threadfence_block();
    int state = bi;
    const int words = (C + 3) / 4; // <= 64
    for (int t0 = T - 1; t0 >= 0; t0 -= 8) {
      // rows t0, t0-1, ..., t0-7 in flight; lane holds bytes [4*lane, 4*lane+4) (+128)
      uint32_t lo[8], hi[8];
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int t = t0 - k;
        lo[k] = hi[k] = 0;
        if (t >= 1) {
          const uint32_t* row = reinterpret_cast<const uint32_t*>(mybp + (long long)t * C);
          if (lane < words) lo[k] = __ldcg(row + lane);
          if (lane + 32 < words) hi[k] = __ldcg(row + lane + 32);
        }
      }
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int t = t0 - k;
        if (t < 0) break;
        if (lane == 0) out[t] = state;
        if (t >= 1) {
          const int w = state >> 2;
          const uint32_t word = (w < 32) ? __shfl_sync(0xffffffffu, lo[k], w) : __shfl_sync(0xffffffffu, hi[k], w - 32);
          state = (int)((word >> (8 * (state & 3))) & 0xffu);
        }
      }
    }
  }
}

} // namespace

#ifndef GTNB_HOST_EMU

int launch_viterbi_dense(
    gtnb_ctx* ctx, int B, int T_max, int C, const int32_t* T_dev, const float* emis, int64_t stride,
    const float* trans_dev, uint8_t* bp, int32_t* paths, float* scores) {
  if (B == 0) return GTNB_OK;
  if (C < 1 || C > 256 || (C % 4) != 0)
    return fail(ctx, GTNB_ERR_UNSUPPORTED, "gtnb_viterbi_dense: needs C <= 256 and C % 4 == 0");
  int G = 1;
  while (G < 32 && C * (G * 2) <= kDenseThreads) G *= 2;
  const size_t smem = sizeof(float) * ((size_t)C * (C + 32 / G) + 2 * (size_t)C);
  if (smem > (size_t)kMaxDynamicSmem)
    return fail(ctx, GTNB_ERR_UNSUPPORTED, "gtnb_viterbi_dense: transitions tile does not fit shared memory");
#define LAUNCH_DENSE(G_)                                                                           \
  do {                                                                                             \
    int rc__ = ensure_max_smem(ctx, (const void*)viterbi_dense_kernel<G_>);                        \
    if (rc__) return rc__;                                                                         \
    GTNB_LAUNCH(ctx, "viterbi_dense",                                                              \
                viterbi_dense_kernel<G_><<<B, kDenseThreads, smem, ctx->stream>>>(                 \
                    T_max, C, T_dev, emis, (long long)stride, trans_dev, bp, paths, scores));      \
  } while (0)
  switch (G) {
    case 1: LAUNCH_DENSE(1); break;
    case 2: LAUNCH_DENSE(2); break;
    case 4: LAUNCH_DENSE(4); break;
    case 8: LAUNCH_DENSE(8); break;
    case 16: LAUNCH_DENSE(16); break;
    default: LAUNCH_DENSE(32); break;
  }
#undef LAUNCH_DENSE
  return GTNB_OK;
}

#endif // GTNB_HOST_EMU

} // namespace gtnb

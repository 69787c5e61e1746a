/*
 * k_linear.cu -- forwardScore / viterbiScore of the emissions chain itself,
 * i.e. detail::shortestDistance (shortest.cpp:86-188) on gtn::linearGraph(T, C)
 * (creations.cpp:20-33), and its gradient (shortest.cpp:33-82).
 *
 * On a chain the recursion factorises: score = sum_t lse_c e[t][c] (log) or
 * sum_t max_c e[t][c] (tropical), and d score / d e[t][c] = softmax_c / first-
 * argmax indicator.  Rows are independent, so this is one warp per frame row,
 * fully coalesced, followed by a fixed-order per-utterance reduction.
 */
#include <cuda_runtime.h>
#include <math_constants.h>

#include "gtnb_internal.h"

namespace gtnb {

namespace {

constexpr int kRowWarps = 8;

template <bool TROPICAL>
__global__ void __launch_bounds__(32 * kRowWarps) linear_rows_kernel(
    const int32_t* __restrict__ T,
    int maxT,
    int C,
    const float* __restrict__ emis,
    long long stride,
    float* __restrict__ row_score, // [B][maxT]
    float* __restrict__ grad,
    long long grad_stride,
    const float* __restrict__ deltas,
    float delta_all,
    int overwrite) {
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * kRowWarps + warp;
  if (t >= T[b]) return;
  const float* e = emis + (long long)b * stride + (long long)t * C;
  float mx = -CUDART_INF_F;
  int arg = 0x7fffffff;
  for (int c = lane; c < C; c += 32) {
    const float v = e[c];
    if (v > mx) {
      mx = v;
      arg = c;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (om > mx || (om == mx && oa < arg)) { // first maximum wins (shortest.cpp:124)
      mx = om;
      arg = oa;
    }
  }
  float score = mx;
  float sum = 0.0f;
  const bool finite = (mx != CUDART_INF_F && mx != -CUDART_INF_F);
  if (!TROPICAL && finite) {
    for (int c = lane; c < C; c += 32) sum += expf(e[c] - mx);
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    score = mx + log1pf(sum - 1.0f);
  }
  if (lane == 0) row_score[(long long)b * maxT + t] = score;
  if (grad) {
    const float delta = deltas ? deltas[b] : delta_all;
    float* g = grad + (long long)b * grad_stride + (long long)t * C;
    if (TROPICAL) {
      for (int c = lane; c < C; c += 32) {
        const float v = (c == arg) ? delta : 0.0f;
        g[c] = overwrite ? v : g[c] + v;
      }
    } else {
      // g = exp(e - mx) / exp(lse - mx)   (shortest.cpp:71-72 on a chain)
      const float denom = expf(score - mx);
      for (int c = lane; c < C; c += 32) {
        const float v = expf(e[c] - mx) / denom * delta;
        g[c] = overwrite ? v : g[c] + v;
      }
    }
  }
}

__global__ void __launch_bounds__(256) linear_reduce_kernel(
    const int32_t* __restrict__ T, int maxT, const float* __restrict__ row_score,
    float* __restrict__ scores) {
  __shared__ double part[256];
  const int b = blockIdx.x;
  const int Tb = T[b];
  double s = 0.0;
  for (int t = threadIdx.x; t < Tb; t += 256) s += (double)row_score[(long long)b * maxT + t];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) scores[b] = (float)part[0];
}

} // namespace

int launch_linear_forward(
    gtnb_ctx* ctx, int B, const int32_t* T_dev, int maxT, int C, const float* emis, int64_t stride,
    int tropical, float* scores, float* grad, int64_t grad_stride, const float* deltas,
    float delta_all, int overwrite, float* scratch) {
  if (B == 0) return GTNB_OK;
  // per-frame logsumexp, B * maxT floats: the caller's scratch (no allocation on this stream) or ours
  float* row_score = scratch;
  if (!scratch) {
    int rc = dev_alloc(ctx, &row_score, (long long)B * std::max(maxT, 1));
    if (rc) return rc;
  }
  if (maxT > 0) {
    dim3 grid((maxT + kRowWarps - 1) / kRowWarps, B);
    if (tropical)
      GTNB_LAUNCH(ctx, "linear_rows", linear_rows_kernel<true><<<grid, 32 * kRowWarps, 0, ctx->stream>>>(
          T_dev, maxT, C, emis, stride, row_score, grad, grad_stride, deltas, delta_all, overwrite));
    else
      GTNB_LAUNCH(ctx, "linear_rows", linear_rows_kernel<false><<<grid, 32 * kRowWarps, 0, ctx->stream>>>(
          T_dev, maxT, C, emis, stride, row_score, grad, grad_stride, deltas, delta_all, overwrite));
  }
  GTNB_LAUNCH(ctx, "linear_reduce", linear_reduce_kernel<<<B, 256, 0, ctx->stream>>>(T_dev, maxT, row_score, scores));
  if (!scratch) dev_free(ctx, row_score);
  return GTNB_OK;
}

} // namespace gtnb

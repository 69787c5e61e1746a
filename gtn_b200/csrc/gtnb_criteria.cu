/*
 * gtnb_criteria.cu -- whole-minibatch criteria: what the reference does with
 * parallelMap(fwd) + parallelMap(bwd) (benchmarks/ctc.cpp:150-165) as ONE call
 * that only enqueues device work: no per-utterance host objects, one pinned
 * staging upload, one small read-back.
 */
#include <algorithm>
#include <cstring>
#include <vector>

#include "gtnb_internal.h"

using namespace gtnb;

extern "C" int gtnb_ctc_loss(
    gtnb_ctx* ctx, int B, int T, int C, const float* emissions, int emissions_on_device,
    const int32_t* input_lens, const int32_t* targets, const int32_t* target_lens, int blank,
    float* losses_host, float* grads, int grads_on_device) {
  if (!ctx || B < 0 || T < 0 || C <= 0 || !emissions || !target_lens || !losses_host)
    return fail(ctx, GTNB_ERR_INVALID_ARGUMENT, "gtnb_ctc_loss: bad arguments");
  if (B == 0) return GTNB_OK;
  GTNB_CUDA(ctx, cudaSetDevice(ctx->device));
  const long long per = (long long)T * C;
  int rc = GTNB_OK;
  float* e_dev = nullptr;
  float* g_dev = nullptr;
  float* z_dev = nullptr;
  float* deltas_dev = nullptr;
  int32_t* small_dev = nullptr; // [targets | offsets | lens | T]
  gtnb_lattice* lat = nullptr;
  std::vector<long long> sgn, sga;

  // host: only sizes.  Graph b has 2U+1 nodes; arcs = self loops + step arcs + skip arcs
  // (benchmarks/ctc.cpp:40-58), skip arcs exist where consecutive labels differ.
  std::vector<SgDims> dims(B);
  std::vector<int32_t> Tb(B), off(B);
  long long tot_t = 0;
  for (int b = 0; b < B; b++) {
    const int U = target_lens[b];
    if (U < 0) return fail(ctx, GTNB_ERR_INVALID_ARGUMENT, "gtnb_ctc_loss: negative target length");
    if (U > 0 && !targets) return fail(ctx, GTNB_ERR_INVALID_ARGUMENT, "gtnb_ctc_loss: targets is NULL");
    off[b] = (int32_t)tot_t;
    const int32_t* tg = targets + tot_t;
    int skips = 0;
    for (int u = 1; u < U; u++) skips += tg[u] != tg[u - 1];
    const int L = 2 * U + 1;
    dims[b] = SgDims{L, L + (L - 1) + skips, L >= 2 ? 2 : 1};
    tot_t += U;
    Tb[b] = input_lens ? input_lens[b] : T;
    if (Tb[b] < 0 || Tb[b] > T)
      return fail(ctx, GTNB_ERR_INVALID_ARGUMENT, "gtnb_ctc_loss: input_lens out of range");
  }
  int maxT = 0;
  for (int b = 0; b < B; b++) maxT = std::max(maxT, Tb[b]);

#define TRY(x)                 \
  do {                         \
    if ((rc = (x))) goto done; \
  } while (0)
#define TRYCUDA(call)                                      \
  do {                                                     \
    cudaError_t e__ = (call);                              \
    if (e__ != cudaSuccess) {                              \
      rc = cuda_fail(ctx, e__, #call, __FILE__, __LINE__); \
      goto done;                                           \
    }                                                      \
  } while (0)

  if (emissions_on_device) {
    e_dev = const_cast<float*>(emissions);
  } else {
    TRY(dev_alloc(ctx, &e_dev, per * B));
    TRYCUDA(cudaMemcpyAsync(e_dev, emissions, sizeof(float) * per * B, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (grads) {
    if (grads_on_device)
      g_dev = grads;
    else
      TRY(dev_alloc(ctx, &g_dev, per * B));
    TRYCUDA(cudaMemsetAsync(g_dev, 0, sizeof(float) * per * B, ctx->stream));
  }
  TRY(dev_alloc(ctx, &z_dev, B));
  TRY(dev_alloc(ctx, &deltas_dev, B));
  TRY(dev_alloc(ctx, &small_dev, tot_t + 3ll * B));
  TRY(composed_alloc(ctx, B, dims.data(), B, 0, Tb.data(), C, e_dev, per, sgn, sga, &lat));

  // one pinned staging pass for everything the kernels need from the host
  TRY(stage_begin(ctx));
  TRY(stage_upload(ctx, lat->meta, lat->meta_h.data(), sizeof(GraphMeta) * B));
  if (tot_t) TRY(stage_upload(ctx, small_dev, targets, sizeof(int32_t) * tot_t));
  TRY(stage_upload(ctx, small_dev + tot_t, off.data(), sizeof(int32_t) * B));
  TRY(stage_upload(ctx, small_dev + tot_t + B, target_lens, sizeof(int32_t) * B));
  TRY(stage_upload(ctx, small_dev + tot_t + 2ll * B, Tb.data(), sizeof(int32_t) * B));
  {
    std::vector<float> minus1(B, -1.0f); // subtract's gradFunc, functions.cpp:53-58
    TRY(stage_upload(ctx, deltas_dev, minus1.data(), sizeof(float) * B));
  }
  TRY(stage_end(ctx));

  // forwardScore(emissions) and its +1 gradient
  TRY(launch_linear_forward(ctx, B, small_dev + tot_t + 2ll * B, maxT, C, e_dev, per, 0, z_dev, g_dev,
                            per, nullptr, 1.0f));
  // ctcGraph -> intersect(ctc, emissions) -> forwardScore -> backward(-1) -> compose gradFunc
  TRY(launch_ctc_build(ctx, lat, small_dev, small_dev + tot_t, small_dev + tot_t + B, blank));
  TRY(launch_compose(ctx, lat));
  TRY(launch_forward(ctx, lat, MODE_LOG));
  lat->forward_done = true;
  lat->forward_mode = MODE_LOG;
  if (grads) {
    if (!lat->arc_grad) TRY(dev_alloc(ctx, &lat->arc_grad, lat->tot_A));
    TRY(launch_backward(ctx, lat, 0, deltas_dev));
    TRY(launch_compose_grad(ctx, lat, nullptr, g_dev, per));
    if (!grads_on_device)
      TRYCUDA(cudaMemcpyAsync(grads, g_dev, sizeof(float) * per * B, cudaMemcpyDeviceToHost, ctx->stream));
  }
  TRY(readback_reserve(ctx, 2 * sizeof(float) * B));
  {
    float* z = reinterpret_cast<float*>(ctx->readback);
    float* s = z + B;
    TRYCUDA(cudaMemcpyAsync(z, z_dev, sizeof(float) * B, cudaMemcpyDeviceToHost, ctx->stream));
    TRYCUDA(cudaMemcpyAsync(s, lat->out_scores, sizeof(float) * B, cudaMemcpyDeviceToHost, ctx->stream));
    TRYCUDA(cudaStreamSynchronize(ctx->stream));
    for (int b = 0; b < B; b++) losses_host[b] = z[b] - s[b]; // subtract, functions.cpp:52
  }

done:
  if (lat) gtnb_lattice_destroy(ctx, lat);
  if (!emissions_on_device) dev_free(ctx, e_dev);
  if (grads && !grads_on_device) dev_free(ctx, g_dev);
  dev_free(ctx, z_dev);
  dev_free(ctx, deltas_dev);
  dev_free(ctx, small_dev);
  return rc;
#undef TRY
#undef TRYCUDA
}

/*
 * viterbiPath(compose(linearGraph(T, C; e_b), transitions)) for a whole minibatch, with the
 * dense transitions graph of test/criterion_test.cpp:244-254 / :316-326, through the
 * factored kernel (k_dense.cu): the T*C*C lattice is never materialised.
 */
extern "C" int gtnb_viterbi_dense(
    gtnb_ctx* ctx, int B, int T, int C, const float* emissions, int emissions_on_device,
    const int32_t* input_lens, const float* trans_w_host, int32_t* paths_host, float* scores_host) {
  if (!ctx || B < 0 || T < 0 || C <= 0 || !emissions || !trans_w_host)
    return fail(ctx, GTNB_ERR_INVALID_ARGUMENT, "gtnb_viterbi_dense: bad arguments");
  if (B == 0) return GTNB_OK;
  GTNB_CUDA(ctx, cudaSetDevice(ctx->device));
  const long long per = (long long)T * C;
  int rc = GTNB_OK;
  float *e_dev = nullptr, *tr_dev = nullptr, *sc_dev = nullptr;
  int32_t *T_dev = nullptr, *paths_dev = nullptr;
  uint8_t* bp = nullptr;
#define TRY(x)                 \
  do {                         \
    if ((rc = (x))) goto done; \
  } while (0)
#define TRYCUDA(call)                                      \
  do {                                                     \
    cudaError_t e__ = (call);                              \
    if (e__ != cudaSuccess) {                              \
      rc = cuda_fail(ctx, e__, #call, __FILE__, __LINE__); \
      goto done;                                           \
    }                                                      \
  } while (0)
  if (input_lens)
    for (int b = 0; b < B; b++)
      if (input_lens[b] < 0 || input_lens[b] > T)
        return fail(ctx, GTNB_ERR_INVALID_ARGUMENT, "gtnb_viterbi_dense: input_lens out of range");
  if (emissions_on_device) {
    e_dev = const_cast<float*>(emissions);
  } else {
    TRY(dev_alloc(ctx, &e_dev, per * B));
    TRYCUDA(cudaMemcpyAsync(e_dev, emissions, sizeof(float) * per * B, cudaMemcpyHostToDevice, ctx->stream));
  }
  TRY(dev_alloc(ctx, &tr_dev, (long long)C + (long long)C * C));
  TRY(dev_alloc(ctx, &sc_dev, B));
  TRY(dev_alloc(ctx, &paths_dev, (long long)B * std::max(T, 1)));
  TRY(dev_alloc(ctx, &bp, (long long)B * std::max(T, 1) * C));
  if (input_lens) TRY(dev_alloc(ctx, &T_dev, B));
  TRY(stage_begin(ctx));
  TRY(stage_upload(ctx, tr_dev, trans_w_host, sizeof(float) * ((size_t)C + (size_t)C * C)));
  if (input_lens) TRY(stage_upload(ctx, T_dev, input_lens, sizeof(int32_t) * B));
  TRY(stage_end(ctx));
  TRY(launch_viterbi_dense(ctx, B, T, C, T_dev, e_dev, per, tr_dev, bp, paths_dev, sc_dev));
  if (paths_host && T > 0)
    TRYCUDA(cudaMemcpyAsync(paths_host, paths_dev, sizeof(int32_t) * (size_t)B * T, cudaMemcpyDeviceToHost, ctx->stream));
  if (scores_host)
    TRYCUDA(cudaMemcpyAsync(scores_host, sc_dev, sizeof(float) * B, cudaMemcpyDeviceToHost, ctx->stream));
  TRYCUDA(cudaStreamSynchronize(ctx->stream));
done:
  if (!emissions_on_device) dev_free(ctx, e_dev);
  dev_free(ctx, tr_dev);
  dev_free(ctx, sc_dev);
  dev_free(ctx, paths_dev);
  dev_free(ctx, bp);
  dev_free(ctx, T_dev);
  return rc;
#undef TRY
#undef TRYCUDA
}

/*
 * gtnb_criteria.cu -- whole-minibatch criteria: what the reference does with
 * parallelMap(fwd) + parallelMap(bwd) (benchmarks/ctc.cpp:150-165) as ONE call
 * that only enqueues device work.
 */
#include <algorithm>
#include <cstring>
#include <vector>

#include "gtnb_internal.h"

using namespace gtnb;

namespace {

/* CTC target graph, benchmarks/ctc.cpp:40-58 (including its arcSort()). */
struct CtcGraphHost {
  std::vector<uint8_t> flags;
  std::vector<int32_t> src, dst, label, in_ptr, in_arcs, out_ptr, out_arcs, accept;

  void build(const int32_t* target, int U, int blank) {
    const int L = 2 * U + 1;
    flags.assign(L, 0);
    src.clear();
    dst.clear();
    label.clear();
    std::vector<std::vector<int32_t>> in(L), out(L);
    auto add = [&](int s, int d, int lab) {
      int a = (int)src.size();
      src.push_back(s);
      dst.push_back(d);
      label.push_back(lab);
      out[s].push_back(a);
      in[d].push_back(a);
    };
    accept.clear();
    for (int l = 0; l < L; l++) {
      const int idx = (l - 1) / 2;
      flags[l] = (l == 0 ? 1 : 0) | ((l == L - 1 || l == L - 2) ? 2 : 0);
      if (flags[l] & 2) accept.push_back(l);
      const int lab = (l % 2) ? target[idx] : blank;
      add(l, l, lab);
      if (l > 0) add(l - 1, l, lab);
      if ((l % 2) && l > 1 && lab != target[idx - 1]) add(l - 2, l, lab);
    }
    // Graph::arcSort (graph.cpp:162-177): per-node sort of the in/out lists by label
    auto by_label = [&](int a, int b) { return label[a] < label[b]; };
    in_ptr.assign(L + 1, 0);
    out_ptr.assign(L + 1, 0);
    in_arcs.clear();
    out_arcs.clear();
    for (int n = 0; n < L; n++) {
      std::stable_sort(in[n].begin(), in[n].end(), by_label);
      std::stable_sort(out[n].begin(), out[n].end(), by_label);
      in_arcs.insert(in_arcs.end(), in[n].begin(), in[n].end());
      out_arcs.insert(out_arcs.end(), out[n].begin(), out[n].end());
      in_ptr[n + 1] = (int)in_arcs.size();
      out_ptr[n + 1] = (int)out_arcs.size();
    }
  }

  gtnb_graph_view view() const {
    gtnb_graph_view v;
    std::memset(&v, 0, sizeof(v));
    v.num_nodes = (int)flags.size();
    v.num_arcs = (int)src.size();
    v.node_flags = flags.data();
    v.arc_src = src.data();
    v.arc_dst = dst.data();
    v.arc_ilabel = label.data();
    v.arc_olabel = label.data();
    v.weights = nullptr;
    v.in_ptr = in_ptr.data();
    v.in_arcs = in_arcs.data();
    v.out_ptr = out_ptr.data();
    v.out_arcs = out_arcs.data();
    v.accept = accept.data();
    v.num_accept = (int)accept.size();
    return v;
  }
};

} // namespace

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
  float *z_dev = nullptr;
  gtnb_lattice* lat = nullptr;
  std::vector<float> z(B), s(B);

  // host: B tiny target graphs
  std::vector<CtcGraphHost> ctc(B);
  std::vector<gtnb_graph_view> views(B);
  std::vector<int32_t> Tb(B);
  {
    long long off = 0;
    for (int b = 0; b < B; b++) {
      ctc[b].build(targets + off, target_lens[b], blank);
      off += target_lens[b];
      views[b] = ctc[b].view();
      Tb[b] = input_lens ? input_lens[b] : T;
      if (Tb[b] < 0 || Tb[b] > T)
        return fail(ctx, GTNB_ERR_INVALID_ARGUMENT, "gtnb_ctc_loss: input_lens out of range");
    }
  }

#define TRY(x)                 \
  do {                         \
    if ((rc = (x))) goto done; \
  } while (0)
#define TRYCUDA(call)                                                      \
  do {                                                                     \
    cudaError_t e__ = (call);                                              \
    if (e__ != cudaSuccess) {                                              \
      rc = cuda_fail(ctx, e__, #call, __FILE__, __LINE__);                 \
      goto done;                                                           \
    }                                                                      \
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

  // forwardScore(emissions) and its +1 gradient
  TRY(gtnb_linear_forward(ctx, B, Tb.data(), C, e_dev, per, 0, z_dev, g_dev, per, nullptr, 1.0f));
  // intersect(ctc, emissions) -> forwardScore -> backward(-1) -> compose gradFunc
  TRY(gtnb_compose_linear(ctx, B, views.data(), B, 0, Tb.data(), C, e_dev, per, &lat));
  TRY(gtnb_forward(ctx, lat, 0, nullptr, nullptr));
  if (grads) {
    std::vector<float> deltas(B, -1.0f); // subtract's gradFunc, functions.cpp:53-58
    TRY(gtnb_backward(ctx, lat, 0, deltas.data()));
    TRY(gtnb_compose_grad(ctx, lat, nullptr, g_dev, per));
    if (!grads_on_device)
      TRYCUDA(cudaMemcpyAsync(grads, g_dev, sizeof(float) * per * B, cudaMemcpyDeviceToHost, ctx->stream));
  }
  TRYCUDA(cudaMemcpyAsync(z.data(), z_dev, sizeof(float) * B, cudaMemcpyDeviceToHost, ctx->stream));
  TRYCUDA(cudaMemcpyAsync(s.data(), lat->out_scores, sizeof(float) * B, cudaMemcpyDeviceToHost, ctx->stream));
  TRYCUDA(cudaStreamSynchronize(ctx->stream));
  for (int b = 0; b < B; b++) losses_host[b] = z[b] - s[b]; // subtract, functions.cpp:52

done:
  if (lat) gtnb_lattice_destroy(ctx, lat);
  if (!emissions_on_device) dev_free(ctx, e_dev);
  if (grads && !grads_on_device) dev_free(ctx, g_dev);
  dev_free(ctx, z_dev);
  return rc;
#undef TRY
#undef TRYCUDA
}

/*
 * examples/ctc_cpp_api.cpp -- the reference's batched CTC benchmark
 * (benchmarks/ctc.cpp:136-168, timeBatchedCtc) written against the drop-in C++ headers:
 * user code builds one target graph and one emissions graph per utterance and maps
 * fwd / bwd over the minibatch with parallelMap; every worker thread drives its own
 * CUDA stream ("one utterance per stream").  The result is cross-checked against the
 * batched C-ABI entry point gtnb_ctc_loss.
 *
 *   ctc_cpp_api [B=32] [T=200] [C=32] [U=20] [iters=3]
 */
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "gtn/gtn.h"
#include "gtn_b200.h"

using namespace gtn;

static Graph ctcGraph(const std::vector<int>& target) {
  const int blank = 0;
  const size_t L = 2 * target.size() + 1;
  Graph ctc;
  for (size_t l = 0; l < L; l++) {
    const size_t idx = (l - 1) / 2;
    ctc.addNode(l == 0, l == L - 1 || l == L - 2);
    const int label = l % 2 ? target[idx] : blank;
    ctc.addArc(l, l, label);
    if (l > 0) ctc.addArc(l - 1, l, label);
    if (l % 2 && l > 1 && label != target[idx - 1]) ctc.addArc(l - 2, l, label);
  }
  ctc.arcSort();
  return ctc;
}

int main(int argc, char** argv) {
  const int B = argc > 1 ? std::atoi(argv[1]) : 32;
  const int T = argc > 2 ? std::atoi(argv[2]) : 200;
  const int C = argc > 3 ? std::atoi(argv[3]) : 32;
  const int U = argc > 4 ? std::atoi(argv[4]) : 20;
  const int iters = argc > 5 ? std::atoi(argv[5]) : 3;

  std::vector<std::vector<int>> targets(B);
  std::vector<std::vector<float>> emissions(B);
  std::vector<int32_t> flatTargets, lens(B, U);
  std::vector<float> flatEmissions;
  for (int b = 0; b < B; b++) {
    std::mt19937 rng(1234 + b);
    std::uniform_real_distribution<float> uni(-5.0f, 5.0f);
    std::uniform_int_distribution<int> lab(1, C - 1);
    for (int u = 0; u < U; u++) targets[b].push_back(lab(rng));
    emissions[b].resize((size_t)T * C);
    for (auto& v : emissions[b]) v = uni(rng);
    flatTargets.insert(flatTargets.end(), targets[b].begin(), targets[b].end());
    flatEmissions.insert(flatEmissions.end(), emissions[b].begin(), emissions[b].end());
  }

  std::vector<Graph> emGraphs(B);
  std::vector<int> idx(B);
  for (int b = 0; b < B; b++) idx[b] = b;
  auto fwd = [&](int b) {
    auto ctc = ctcGraph(targets[b]);
    auto e = linearGraph(T, C);
    e.setWeights(emissions[b].data());
    emGraphs[b] = e;
    return subtract(forwardScore(e), forwardScore(intersect(ctc, e)));
  };
  auto bwd = [](const Graph& g) { backward(g); };

  std::vector<Graph> losses;
  double best = 1e30;
  for (int it = 0; it < iters; it++) {
    auto t0 = std::chrono::steady_clock::now();
    losses = parallelMap(fwd, idx);
    parallelMap(bwd, losses);
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    best = std::min(best, sec);
  }

  // the batched C-ABI criterion on the same inputs
  gtnb_ctx* ctx = nullptr;
  if (gtnb_ctx_create(0, nullptr, &ctx) != GTNB_OK) {
    std::fprintf(stderr, "%s\n", gtnb_last_error(nullptr));
    return 2;
  }
  std::vector<float> refLoss(B), refGrad((size_t)B * T * C);
  int rc = gtnb_ctc_loss(ctx, B, T, C, flatEmissions.data(), 0, nullptr, flatTargets.data(), lens.data(), 0,
                         refLoss.data(), refGrad.data(), 0);
  if (rc != GTNB_OK) {
    std::fprintf(stderr, "gtnb_ctc_loss: %s\n", gtnb_last_error(ctx));
    return 2;
  }
  double maxLoss = 0, maxGrad = 0;
  for (int b = 0; b < B; b++) {
    maxLoss = std::max(maxLoss, (double)std::fabs(losses[b].item() - refLoss[b]) / std::fabs(refLoss[b]));
    const float* g = emGraphs[b].grad().weights();
    for (size_t i = 0; i < (size_t)T * C; i++)
      maxGrad = std::max(maxGrad, (double)std::fabs(g[i] - refGrad[(size_t)b * T * C + i]));
  }
  gtnb_ctx_destroy(ctx);
  std::printf("{\"B\": %d, \"T\": %d, \"C\": %d, \"U\": %d, \"cpp_api_utt_per_s\": %.1f, "
              "\"max_rel_loss_diff\": %.3g, \"max_abs_grad_diff\": %.3g}\n",
              B, T, C, U, B / best, maxLoss, maxGrad);
  return (maxLoss < 1e-4 && maxGrad < 1e-3) ? 0 : 1;
}

/*
 * oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, not product code.
 *
 * A small C-ABI shim (written for this repo) around the UNMODIFIED reference
 * library.  The reference sources are compiled where they lie under
 * /root/reference by oracle/Makefile (`make ref`) and linked with this file
 * into oracle/_ref/libgtn_ref.so.  Nothing from the reference is copied into
 * the repo; this file only *calls* its public API:
 *   gtn::Graph                     (gtn/graph.h:56-465)
 *   gtn::compose / gtn::intersect  (gtn/functions.h:112,125)
 *   gtn::forwardScore / viterbiScore / viterbiPath (gtn/functions.h:134-152)
 *   gtn::backward                  (gtn/autograd.h:27,37)
 *   gtn::linearGraph / scalarGraph (gtn/creations.h:25,32)
 *   gtn::parallelMap               (gtn/parallel/parallel_map.h:153)
 *   gtn::concat / closure / union_ / remove / clone (gtn/functions.h:42-104)
 *   gtn::equal / isomorphic / save / load / saveTxt / loadTxt / draw (gtn/utils.h:23-153)
 *   gtn::sample / randEquivalent   (gtn/rand.h:22-40)
 *
 * Users: tests/ (to pin oracle/gtn_oracle.c and as the parity checker),
 * tests/golden/make_golden.py (fixture generator) and bench.py's
 * cpu_baseline / --impl reference legs.  The product path (gtn_b200/) never
 * loads this library.
 */
#include <chrono>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "gtn/gtn.h"

using namespace gtn;

namespace {

std::mutex gMutex;
std::map<int, Graph> gTable;
int gNext = 1;
thread_local std::string gErr;

int put(Graph g) {
  std::lock_guard<std::mutex> l(gMutex);
  int h = gNext++;
  gTable.emplace(h, std::move(g));
  return h;
}

Graph get(int h) {
  std::lock_guard<std::mutex> l(gMutex);
  auto it = gTable.find(h);
  if (it == gTable.end()) {
    throw std::out_of_range("ref_driver: bad graph handle");
  }
  return it->second;
}

// Error classes mirror the pybind11 exception translation the reference's
// Python tests pin (bindings/python/test/test_bindings.py:302,319).
template <typename F>
int guard(F&& f) {
  try {
    return f();
  } catch (const std::invalid_argument& e) {
    gErr = e.what();
    return -2; // ValueError
  } catch (const std::logic_error& e) {
    gErr = e.what();
    return -3; // RuntimeError (logic_error)
  } catch (const std::exception& e) {
    gErr = e.what();
    return -4;
  }
}

// CTC target graph exactly as the reference benchmark builds it
// (benchmarks/ctc.cpp:40-58), parameterised on the blank index; with
// arc_sort=0 it is the test-suite variant (test/criterion_test.cpp:37-54).
Graph makeCtc(const int* target, int U, int blank, bool arcSort) {
  size_t L = 2 * (size_t)U + 1;
  Graph ctc;
  for (size_t l = 0; l < L; l++) {
    size_t idx = (l - 1) / 2;
    ctc.addNode(l == 0, l == L - 1 || l == L - 2);
    int label = l % 2 ? target[idx] : blank;
    ctc.addArc(l, l, label);
    if (l > 0) {
      ctc.addArc(l - 1, l, label);
    }
    if (l % 2 && l > 1 && label != target[idx - 1]) {
      ctc.addArc(l - 2, l, label);
    }
  }
  if (arcSort) {
    ctc.arcSort();
  }
  return ctc;
}

// ASG transitions graph (test/criterion_test.cpp:244-254): node 0 start,
// nodes 1..C accept; arcs 0->i+1 label i (C of them, weights w[0..C)), then
// for i, for j: arc j+1 -> i+1 label i, weight w[C + i*C + j].
Graph makeTransitions(int C, const float* w, bool calcGrad) {
  Graph tr(calcGrad);
  tr.addNode(true);
  for (int i = 1; i <= C; i++) {
    tr.addNode(false, true);
    tr.addArc(0, i, i - 1, i - 1, w[i - 1]);
  }
  for (int i = 0; i < C; i++) {
    for (int j = 0; j < C; j++) {
      tr.addArc(j + 1, i + 1, i, i, w[C + i * C + j]);
    }
  }
  return tr;
}

Graph makeFal(const int* target, int U) {
  Graph fal;
  fal.addNode(true);
  for (int l = 1; l <= U; l++) {
    fal.addNode(false, l == U);
    fal.addArc(l - 1, l, target[l - 1]);
    fal.addArc(l, l, target[l - 1]);
  }
  return fal;
}

double now() {
  return std::chrono::duration<double>(
             std::chrono::steady_clock::now().time_since_epoch())
      .count();
}

} // namespace

extern "C" {

const char* ref_last_error() {
  return gErr.c_str();
}

int ref_hardware_threads() {
  return (int)std::thread::hardware_concurrency();
}

/* ---------------- generic handle API ---------------- */

int ref_graph_new(
    int numNodes,
    const uint8_t* flags, /* bit0 start, bit1 accept */
    int numArcs,
    const int* src,
    const int* dst,
    const int* ilabel,
    const int* olabel,
    const float* w,
    int calcGrad) {
  return guard([&] {
    Graph g(calcGrad != 0);
    for (int n = 0; n < numNodes; n++) {
      g.addNode(flags[n] & 1, flags[n] & 2);
    }
    for (int a = 0; a < numArcs; a++) {
      g.addArc(src[a], dst[a], ilabel[a], olabel[a], w ? w[a] : 0.0f);
    }
    return put(g);
  });
}

int ref_linear_graph(int T, int C, int calcGrad) {
  return guard([&] { return put(linearGraph(T, C, calcGrad != 0)); });
}

int ref_scalar_graph(float v, int calcGrad) {
  return guard([&] { return put(scalarGraph(v, calcGrad != 0)); });
}

int ref_graph_free(int h) {
  std::lock_guard<std::mutex> l(gMutex);
  gTable.erase(h);
  return 0;
}

int ref_graph_free_all() {
  std::lock_guard<std::mutex> l(gMutex);
  gTable.clear();
  return 0;
}

int ref_graph_num_nodes(int h) {
  return guard([&] { return (int)get(h).numNodes(); });
}
int ref_graph_num_arcs(int h) {
  return guard([&] { return (int)get(h).numArcs(); });
}

int ref_graph_set_weights(int h, const float* w) {
  return guard([&] {
    get(h).setWeights(w);
    return 0;
  });
}

int ref_graph_arc_sort(int h, int olabel) {
  return guard([&] {
    get(h).arcSort(olabel != 0);
    return 0;
  });
}

/* Dump a graph: any pointer may be NULL. */
int ref_graph_get(
    int h,
    uint8_t* flags,
    int* src,
    int* dst,
    int* ilabel,
    int* olabel,
    float* w) {
  return guard([&] {
    Graph g = get(h);
    for (size_t n = 0; flags && n < g.numNodes(); n++) {
      flags[n] = (g.isStart(n) ? 1 : 0) | (g.isAccept(n) ? 2 : 0);
    }
    for (size_t a = 0; a < g.numArcs(); a++) {
      if (src) src[a] = g.srcNode(a);
      if (dst) dst[a] = g.dstNode(a);
      if (ilabel) ilabel[a] = g.ilabel(a);
      if (olabel) olabel[a] = g.olabel(a);
      if (w) w[a] = g.weight(a);
    }
    return 0;
  });
}

/* In-/out-arc lists in the reference's per-node order (after any arcSort). */
int ref_graph_get_adj(int h, int* inPtr, int* inArcs, int* outPtr, int* outArcs) {
  return guard([&] {
    Graph g = get(h);
    int pi = 0, po = 0;
    for (size_t n = 0; n < g.numNodes(); n++) {
      inPtr[n] = pi;
      outPtr[n] = po;
      for (auto a : g.in(n)) inArcs[pi++] = a;
      for (auto a : g.out(n)) outArcs[po++] = a;
    }
    inPtr[g.numNodes()] = pi;
    outPtr[g.numNodes()] = po;
    return 0;
  });
}

int ref_graph_item(int h, float* out) {
  return guard([&] {
    *out = get(h).item();
    return 0;
  });
}

int ref_graph_grad(int h) {
  return guard([&] { return put(get(h).grad()); });
}

int ref_graph_zero_grad(int h) {
  return guard([&] {
    get(h).zeroGrad();
    return 0;
  });
}

int ref_compose(int a, int b) {
  return guard([&] { return put(compose(get(a), get(b))); });
}
int ref_intersect(int a, int b) {
  return guard([&] { return put(intersect(get(a), get(b))); });
}
int ref_forward_score(int a) {
  return guard([&] { return put(forwardScore(get(a))); });
}
int ref_viterbi_score(int a) {
  return guard([&] { return put(viterbiScore(get(a))); });
}
int ref_viterbi_path(int a) {
  return guard([&] { return put(viterbiPath(get(a))); });
}
int ref_negate(int a) {
  return guard([&] { return put(negate(get(a))); });
}
int ref_add(int a, int b) {
  return guard([&] { return put(add(get(a), get(b))); });
}
int ref_subtract(int a, int b) {
  return guard([&] { return put(subtract(get(a), get(b))); });
}
int ref_backward(int a, int retain) {
  return guard([&] {
    backward(get(a), retain != 0);
    return 0;
  });
}
/* backward with an explicit seed graph (autograd.h:37) */
int ref_backward_with(int a, int seed, int retain) {
  return guard([&] {
    backward(get(a), get(seed), retain != 0);
    return 0;
  });
}

/* ---------------- host API: rational ops, utils, rand ---------------- */

int ref_concat(const int* hs, int n) {
  return guard([&] {
    std::vector<Graph> v;
    for (int i = 0; i < n; i++) v.push_back(get(hs[i]));
    return put(concat(v));
  });
}
int ref_union(const int* hs, int n) {
  return guard([&] {
    std::vector<Graph> v;
    for (int i = 0; i < n; i++) v.push_back(get(hs[i]));
    return put(union_(v));
  });
}
int ref_closure(int a) {
  return guard([&] { return put(closure(get(a))); });
}
int ref_remove(int a, int ilabel, int olabel) {
  return guard([&] { return put(remove(get(a), ilabel, olabel)); });
}
int ref_clone(int a, int projection) {
  return guard([&] { return put(clone(get(a), static_cast<Projection>(projection))); });
}
int ref_equal(int a, int b) {
  return guard([&] { return equal(get(a), get(b)) ? 1 : 0; });
}
int ref_isomorphic(int a, int b) {
  return guard([&] { return isomorphic(get(a), get(b)) ? 1 : 0; });
}

/* serialisers: kind 0 = binary save(), 1 = saveTxt(), 2 = operator<<, 3 = draw() without symbol
 * maps.  Returns the byte count (copied into buf when it fits in cap). */
long long ref_serialise(int h, int kind, char* buf, long long cap) {
  long long n = -1;
  int rc = guard([&] {
    std::ostringstream os;
    Graph g = get(h);
    if (kind == 0) {
      save(os, g);
    } else if (kind == 1) {
      saveTxt(os, g);
    } else if (kind == 2) {
      os << g;
    } else {
      draw(g, os);
    }
    std::string str = os.str();
    n = (long long)str.size();
    if (buf && n <= cap) std::memcpy(buf, str.data(), str.size());
    return 0;
  });
  return rc < 0 ? rc : n;
}
/* kind 0 = binary load(), 1 = loadTxt() */
int ref_parse(int kind, const char* buf, long long n) {
  return guard([&] {
    std::istringstream in(std::string(buf, (size_t)n));
    return put(kind == 0 ? load(in) : loadTxt(in));
  });
}
/* sample() after std::srand(seed) (rand.cpp:14-72) */
int ref_sample(int h, unsigned seed, long long maxLength) {
  return guard([&] {
    std::srand(seed);
    return put(sample(get(h), (size_t)maxLength));
  });
}
int ref_rand_equivalent(int a, int b, unsigned seed, int numSamples) {
  return guard([&] {
    std::srand(seed);
    return randEquivalent(get(a), get(b), (size_t)numSamples) ? 1 : 0;
  });
}

/* ---------------- batched criteria (benchmark workloads) ---------------- */

/*
 * Batched CTC fwd(+bwd) exactly as benchmarks/ctc.cpp:150-165:
 *   loss_b = forwardScore(emissions_b) - forwardScore(intersect(ctc_b, emissions_b))
 *   parallelMap(fwd) then parallelMap(bwd).
 * emissions [B][T][C]; targets concatenated, targetLens[B].
 * grads (nullable) [B][T][C].  Returns wall seconds of fwd(+bwd) (excluding
 * result extraction), negative on error.
 */
double ref_ctc_batch(
    int B,
    int T,
    int C,
    const float* emissions,
    const int* targets,
    const int* targetLens,
    int blank,
    int arcSortCtc,
    float* losses,
    float* grads) {
  double elapsed = -1;
  int rc = guard([&] {
    std::vector<std::vector<int>> tg(B);
    std::vector<const float*> em(B);
    size_t off = 0;
    for (int b = 0; b < B; b++) {
      tg[b].assign(targets + off, targets + off + targetLens[b]);
      off += targetLens[b];
      em[b] = emissions + (size_t)b * T * C;
    }
    std::vector<Graph> emGraphs(B);
    std::mutex m;
    std::vector<int> idx(B);
    for (int b = 0; b < B; b++) idx[b] = b;
    auto fwd = [&](int b) {
      auto ctc = makeCtc(tg[b].data(), (int)tg[b].size(), blank, arcSortCtc);
      auto e = linearGraph(T, C);
      e.setWeights(em[b]);
      emGraphs[b] = e;
      return subtract(forwardScore(e), forwardScore(intersect(ctc, e)));
    };
    auto bwd = [](const Graph& g) { backward(g); };
    double t0 = now();
    auto lossGraphs = parallelMap(fwd, idx);
    if (grads) {
      parallelMap(bwd, lossGraphs);
    }
    elapsed = now() - t0;
    for (int b = 0; b < B; b++) {
      if (losses) losses[b] = lossGraphs[b].item();
      if (grads) {
        std::memcpy(
            grads + (size_t)b * T * C,
            emGraphs[b].grad().weights(),
            sizeof(float) * T * C);
      }
    }
    return 0;
  });
  return rc < 0 ? (double)rc : elapsed;
}

/*
 * Batched ASG (test/criterion_test.cpp:244-305, examples/asg.cpp:59-81):
 *   loss_b = forwardScore(compose(e_b, trans))
 *          - forwardScore(compose(compose(fal_b, trans), e_b))
 * trans shared across the batch; transW [C + C*C]; transGrad (nullable) same
 * shape, accumulated over the batch.
 */
double ref_asg_batch(
    int B,
    int T,
    int C,
    const float* emissions,
    const float* transW,
    const int* targets,
    const int* targetLens,
    float* losses,
    float* grads,
    float* transGrad) {
  double elapsed = -1;
  int rc = guard([&] {
    Graph trans = makeTransitions(C, transW, true);
    std::vector<std::vector<int>> tg(B);
    size_t off = 0;
    for (int b = 0; b < B; b++) {
      tg[b].assign(targets + off, targets + off + targetLens[b]);
      off += targetLens[b];
    }
    std::vector<Graph> emGraphs(B);
    std::vector<int> idx(B);
    for (int b = 0; b < B; b++) idx[b] = b;
    auto fwd = [&](int b) {
      auto fal = makeFal(tg[b].data(), (int)tg[b].size());
      auto e = linearGraph(T, C);
      e.setWeights(emissions + (size_t)b * T * C);
      emGraphs[b] = e;
      return subtract(
          forwardScore(compose(e, trans)),
          forwardScore(compose(compose(fal, trans), e)));
    };
    auto bwd = [](const Graph& g) { backward(g); };
    double t0 = now();
    auto lossGraphs = parallelMap(fwd, idx);
    if (grads || transGrad) {
      parallelMap(bwd, lossGraphs);
    }
    elapsed = now() - t0;
    for (int b = 0; b < B; b++) {
      if (losses) losses[b] = lossGraphs[b].item();
      if (grads) {
        std::memcpy(
            grads + (size_t)b * T * C,
            emGraphs[b].grad().weights(),
            sizeof(float) * T * C);
      }
    }
    if (transGrad) {
      std::memcpy(
          transGrad, trans.grad().weights(), sizeof(float) * (C + C * C));
    }
    return 0;
  });
  return rc < 0 ? (double)rc : elapsed;
}

/*
 * Batched Viterbi decode over the dense trellis
 * (test/criterion_test.cpp:328-344): viterbiPath(compose(e_b, trans)).
 * paths [B][T] (ilabels of the best path), scores[B] = viterbiScore.
 */
double ref_viterbi_dense_batch(
    int B,
    int T,
    int C,
    const float* emissions,
    const float* transW,
    int* paths,
    float* scores) {
  double elapsed = -1;
  int rc = guard([&] {
    Graph trans = makeTransitions(C, transW, false);
    std::vector<int> idx(B);
    for (int b = 0; b < B; b++) idx[b] = b;
    auto dec = [&](int b) {
      auto e = linearGraph(T, C, false);
      e.setWeights(emissions + (size_t)b * T * C);
      auto lattice = compose(e, trans);
      auto path = viterbiPath(lattice);
      if (scores) scores[b] = viterbiScore(lattice).item();
      auto labels = path.labelsToVector();
      for (int t = 0; t < T; t++) {
        paths[(size_t)b * T + t] = t < (int)labels.size() ? labels[t] : -1;
      }
      return 0;
    };
    double t0 = now();
    parallelMap(dec, idx);
    elapsed = now() - t0;
    return 0;
  });
  return rc < 0 ? (double)rc : elapsed;
}

/*
 * Batched CTC forced alignment: viterbiPath(intersect(ctc_b, e_b)).
 */
double ref_viterbi_ctc_batch(
    int B,
    int T,
    int C,
    const float* emissions,
    const int* targets,
    const int* targetLens,
    int blank,
    int arcSortCtc,
    int* paths,
    float* scores) {
  double elapsed = -1;
  int rc = guard([&] {
    std::vector<std::vector<int>> tg(B);
    size_t off = 0;
    for (int b = 0; b < B; b++) {
      tg[b].assign(targets + off, targets + off + targetLens[b]);
      off += targetLens[b];
    }
    std::vector<int> idx(B);
    for (int b = 0; b < B; b++) idx[b] = b;
    auto dec = [&](int b) {
      auto ctc = makeCtc(tg[b].data(), (int)tg[b].size(), blank, arcSortCtc);
      auto e = linearGraph(T, C, false);
      e.setWeights(emissions + (size_t)b * T * C);
      auto lattice = intersect(ctc, e);
      auto path = viterbiPath(lattice);
      if (scores) scores[b] = viterbiScore(lattice).item();
      auto labels = path.labelsToVector();
      for (int t = 0; t < T; t++) {
        paths[(size_t)b * T + t] = t < (int)labels.size() ? labels[t] : -1;
      }
      return 0;
    };
    double t0 = now();
    parallelMap(dec, idx);
    elapsed = now() - t0;
    return 0;
  });
  return rc < 0 ? (double)rc : elapsed;
}

} // extern "C"

/*
 * gtn/rational.cpp -- the rational operations of the reference (concat, closure, union_,
 * remove; gtn/functions.cpp:93-223, :253-318) for the host side of the gtn:: surface.
 *
 * These are graph *construction*: they run once per criterion graph (lexicons, label
 * alternatives), never per frame, and their outputs feed compose / forwardScore, which is
 * where the device work starts.  Node and arc numbering of
 * every output is the reference's (its gradFuncs slice `deltas` by arc position, and user
 * code indexes the results), pinned by tests/test_host_api.py against the compiled reference.
 */
#include <queue>
#include <stdexcept>

#include "gtn/functions.h"

namespace gtn {

namespace {

/* Copy `g` into `out` with node ids shifted by `shift`; start / accept as given. */
void appendShifted(Graph& out, const Graph& g, size_t shift, bool keepStart, bool keepAccept) {
  for (size_t n = 0; n < g.numNodes(); ++n) {
    out.addNode(keepStart && g.isStart(n), keepAccept && g.isAccept(n));
  }
  for (size_t a = 0; a < g.numArcs(); ++a) {
    out.addArc(shift + g.srcNode(a), shift + g.dstNode(a), g.ilabel(a), g.olabel(a), g.weight(a));
  }
}

std::vector<Graph> weightless(const std::vector<Graph>& graphs) {
  std::vector<Graph> v;
  v.reserve(graphs.size());
  for (auto& g : graphs) v.push_back(g.withoutWeights());
  return v;
}

} // namespace

Graph concat(const Graph& g1, const Graph& g2) {
  return concat(std::vector<Graph>{g1, g2});
}

/* functions.cpp:97-155: output arcs = [arcs of g_0][arcs of g_1][eps arcs accept(g_0) x start(g_1)]
 * [arcs of g_2][eps arcs accept(g_1) x start(g_2)]...  The gradient of input i is the slice of
 * `deltas` holding its own arcs. */
Graph concat(const std::vector<Graph>& graphs) {
  auto gradFunc = [](std::vector<Graph>& inputs, Graph& deltas) {
    const float* d = deltas.weights();
    size_t at = 0;
    for (size_t i = 0; i < inputs.size(); ++i) {
      const size_t na = inputs[i].numArcs();
      if (inputs[i].calcGrad()) {
        inputs[i].addGrad(std::vector<float>(d + at, d + at + na));
      }
      at += na;
      if (i > 0) at += inputs[i - 1].numAccept() * inputs[i].numStart(); // the connecting epsilons
    }
  };
  Graph out(gradFunc, weightless(graphs));
  if (graphs.empty()) { // a^0 accepts the empty string
    out.addNode(true, true);
    return out;
  }
  size_t shift = 0;
  for (size_t i = 0; i < graphs.size(); ++i) {
    const Graph& g = graphs[i];
    appendShifted(out, g, shift, i == 0, i + 1 == graphs.size());
    if (i > 0) {
      const Graph& p = graphs[i - 1];
      const size_t pshift = shift - p.numNodes();
      for (int a : p.accept()) {
        for (int s : g.start()) out.addArc(pshift + a, shift + s, epsilon);
      }
    }
    shift += g.numNodes();
  }
  return out;
}

/* functions.cpp:157-190: node 0 is a new start+accept node; arcs of g first (same order), then
 * epsilons 0 -> old starts, then old accepts -> 0. */
Graph closure(const Graph& g) {
  auto gradFunc = [](std::vector<Graph>& inputs, Graph& deltas) {
    const float* d = deltas.weights();
    inputs[0].addGrad(std::vector<float>(d, d + inputs[0].numArcs()));
  };
  Graph out(gradFunc, {g.withoutWeights()});
  out.addNode(true, true);
  appendShifted(out, g, 1, false, false);
  for (int s : g.start()) out.addArc(0, s + 1, epsilon);
  for (int a : g.accept()) out.addArc(a + 1, 0, epsilon);
  return out;
}

/* functions.cpp:192-223: disjoint union, inputs laid out one after the other. */
Graph union_(const std::vector<Graph>& graphs) {
  auto gradFunc = [](std::vector<Graph>& inputs, Graph& deltas) {
    const float* d = deltas.weights();
    size_t at = 0;
    for (auto& in : inputs) {
      if (in.calcGrad()) in.addGrad(std::vector<float>(d + at, d + at + in.numArcs()));
      at += in.numArcs();
    }
  };
  Graph out(gradFunc, weightless(graphs));
  size_t shift = 0;
  for (auto& g : graphs) {
    appendShifted(out, g, shift, true, true);
    shift += g.numNodes();
  }
  return out;
}

Graph remove(const Graph& g, int label /* = epsilon */) {
  return remove(g, label, label);
}

/* functions.cpp:257-318.  A node survives iff it is a start node or has an in-arc that is
 * not (ilabel, olabel); from every survivor the closure over matching arcs is explored
 * breadth-first, copying the non-matching arcs met on the way (weights are dropped, as in the
 * reference) and inheriting acceptance.  No gradient. */
Graph remove(const Graph& g, int ilabel, int olabel) {
  auto gradFunc = [](std::vector<Graph>&, Graph&) {
    throw std::logic_error("[gtn::remove] gradient compuation not implemented");
  };
  auto matches = [&](int a) { return g.ilabel(a) == ilabel && g.olabel(a) == olabel; };

  const int N = static_cast<int>(g.numNodes());
  std::vector<int> newId(N, -1);
  Graph out(gradFunc, {g});
  for (int n = 0; n < N; ++n) {
    bool keep = g.isStart(n);
    for (size_t k = 0; !keep && k < g.numIn(n); ++k) keep = !matches(g.in(n, k));
    if (keep) newId[n] = out.addNode(g.isStart(n));
  }
  std::vector<int> seenAt(N, -1); // last root from which the node was reached
  std::queue<int> todo;
  for (int root = 0; root < N; ++root) {
    if (newId[root] < 0) continue;
    todo.push(root);
    seenAt[root] = root;
    while (!todo.empty()) {
      const int n = todo.front();
      todo.pop();
      if (g.isAccept(n)) out.makeAccept(newId[root]);
      for (int a : g.out(n)) {
        const int dn = g.dstNode(a);
        if (matches(a)) {
          if (seenAt[dn] != root) {
            seenAt[dn] = root;
            todo.push(dn);
          }
        } else {
          out.addArc(newId[root], newId[dn], g.ilabel(a), g.olabel(a));
        }
      }
    }
  }
  return out;
}

} // namespace gtn

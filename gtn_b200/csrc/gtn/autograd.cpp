/*
 * gtn/autograd.cpp -- the autograd tape driver (reference: gtn/autograd.cpp:17-67).
 * Host C++ on purpose: the tape is a handful of nodes; the gradFuncs it calls are
 * what launch the device kernels.
 */
#include "gtn/autograd.h"

#include <functional>
#include <stdexcept>
#include <unordered_set>

namespace gtn {

namespace {

void backwardImpl(Graph g, bool retainGraph) {
  // depth-first topological order over inputs(), keyed by Graph::id()
  std::unordered_set<std::uintptr_t> seen;
  std::vector<Graph> tape;
  std::function<void(Graph&)> visit = [&](Graph& node) {
    auto id = node.id();
    if (seen.count(id)) return;
    for (auto& in : node.inputs()) visit(in);
    seen.insert(id);
    tape.push_back(node);
  };
  visit(g);

  for (auto it = tape.rbegin(); it != tape.rend(); ++it) {
    if (!it->gradFunc()) continue;
    if (it->inputs().empty()) {
      throw std::invalid_argument(
          "[autograd::backward] Cannot Backward twice without retaining the graph.");
    }
    it->gradFunc()(it->inputs(), it->grad());
    if (!retainGraph) {
      it->inputs().clear();
      *it = Graph{};
    }
  }
}

} // namespace

void backward(Graph g, bool retainGraph /* = false */) {
  g.addGrad(std::vector<float>(g.numArcs(), 1.0f));
  backwardImpl(g, retainGraph);
}

void backward(Graph g, const Graph& grad, bool retainGraph /* = false */) {
  g.addGrad(grad);
  backwardImpl(g, retainGraph);
}

} // namespace gtn

/*
 * gtn/rand.cpp -- reference: gtn/rand.cpp:14-126.  Random walks draw from std::rand() in the
 * reference's order (one draw for the start node, one per step), so a program that seeds
 * std::srand gets the same sample from either library.
 */
#include "gtn/rand.h"

#include <cmath>
#include <cstdlib>

#include "gtn/functions.h"
#include "gtn/utils.h"

namespace gtn {

Graph sample(const Graph& g, size_t maxLength /* = 1000 */) {
  if (g.numStart() == 0 || g.numAccept() == 0) return Graph{};

  size_t node = g.start()[std::rand() % g.numStart()];
  std::vector<int> walk;
  size_t keep = 0; // 1 + number of arcs of the longest prefix that ends in an accept node
  for (size_t len = 0; len <= maxLength; len++) {
    const size_t choices = g.numOut(node) + (g.isAccept(node) ? 1 : 0);
    if (g.isAccept(node)) keep = len + 1;
    if (choices == 0) return Graph{}; // dead end
    const size_t pick = static_cast<size_t>(std::rand() % choices);
    if (pick == g.numOut(node)) break; // stop here (only offered on accept nodes)
    const int arc = g.out(node, pick);
    walk.push_back(arc);
    node = g.dstNode(arc);
  }
  if (keep == 0) return Graph{}; // never saw an accept node
  walk.resize(keep - 1);

  auto gradFunc = [walk](std::vector<Graph>& inputs, Graph& deltas) {
    if (!inputs[0].calcGrad()) return;
    std::vector<float> grad(inputs[0].numArcs(), 0.0f);
    for (size_t a = 0; a < deltas.numArcs(); a++) grad[walk[a]] += deltas.weight(a);
    inputs[0].addGrad(std::move(grad));
  };
  Graph path(gradFunc, {g});
  path.addNode(true, keep == 1);
  for (size_t i = 1; i < keep; i++) {
    path.addNode(false, i + 1 == keep);
    const int arc = walk[i - 1];
    path.addArc(i - 1, i, g.ilabel(arc), g.olabel(arc), g.weight(arc));
  }
  return path;
}

bool randEquivalent(const Graph& g1, const Graph& g2, size_t numSamples /* = 100 */, double tol /* = 1e-4 */,
                    size_t maxLength /* = 1000 */) {
  const Graph empty;
  // score of the sampled path's label sequences in g: forwardScore(in o g o out), or "no such path"
  auto restrictTo = [](const Graph& in, const Graph& g, const Graph& out) {
    Graph left = compose(in, g);
    left.setCalcGrad(false);
    return compose(left, out);
  };
  for (size_t i = 0; i < numSamples; i++) {
    Graph path = sample(std::rand() % 2 ? g1 : g2, maxLength);
    path.setCalcGrad(false);
    if (equal(path, empty)) continue;
    const Graph in = projectInput(path), out = projectOutput(path);
    const Graph c1 = restrictTo(in, g1, out), c2 = restrictTo(in, g2, out);
    const bool none1 = equal(c1, empty), none2 = equal(c2, empty);
    if (none1 != none2) return false;
    if (none1) continue;
    if (std::fabs(forwardScore(c1).item() - forwardScore(c2).item()) > tol) return false;
  }
  return true;
}

} // namespace gtn

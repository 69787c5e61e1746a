/* gtn/creations.cpp -- reference: gtn/creations.cpp:12-33. */
#include "gtn/creations.h"

namespace gtn {

Graph scalarGraph(float val, bool calcGrad) {
  Graph g(calcGrad);
  g.addNode(true);
  g.addNode(false, true);
  g.addArc(0, 1, epsilon, epsilon, val);
  return g;
}

Graph linearGraph(int M, int N, bool calcGrad /* = true */) {
  Graph g(calcGrad);
  if (M <= 0) {
    // no frames: a lone start node that is NOT accepting (creations.cpp:22): forwardScore is -inf.  It has no
    // accept node and takes the generic path like any other graph.
    g.addNode(true);
    g.markArcSorted();
    g.markArcSorted(true);
    return g;
  }
  // The chain of creations.cpp:20-33 -- node m-1 -> m by N arcs labelled 0..N-1, arc id (m-1)*N + n, weights
  // = the row-major [M][N] matrix, both sorted flags set -- but NOT built yet: the device paths only need
  // (M, N) and the weights, and a T = 1000, C = 64 chain is 64,000 addArc calls per utterance (about a
  // millisecond of host time the reference spends inside its timed loop).  The host topology appears the first
  // time somebody looks at it (Graph::materialize).
  g.markArcSorted();
  g.markArcSorted(true);
  g.sharedGraph_->linearFrames = M;
  g.sharedGraph_->linearLabels = N;
  g.sharedGraph_->hostReady = false;
  const size_t n = (size_t)M * (size_t)N;
  g.sharedWeights_->lazyFetch = [n](std::vector<float>& h) { h.assign(n, 0.0f); }; // weights start at 0
  return g;
}

} // namespace gtn

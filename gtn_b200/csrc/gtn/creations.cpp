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
  g.addNode(true); // not accepting, even for M == 0 (creations.cpp:22): forwardScore(linearGraph(0, N)) is -inf
  for (int m = 1; m <= M; ++m) {
    g.addNode(false, m == M);
    for (int n = 0; n < N; ++n) {
      g.addArc(m - 1, m, n); // arc id (m-1)*N + n: weights are the row-major [M][N] matrix
    }
  }
  g.markArcSorted();
  g.markArcSorted(true);
  // remember the structure: compose / forwardScore take the frame-synchronous device path
  // (M == 0 has no accept node: it takes the generic path like any other graph)
  if (M > 0) {
    g.sharedGraph_->linearFrames = M;
    g.sharedGraph_->linearLabels = N;
  }
  return g;
}

} // namespace gtn

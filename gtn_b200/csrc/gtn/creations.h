/* gtn/creations.h -- reference: gtn/creations.h:25,32. */
#pragma once

#include "gtn/graph.h"

namespace gtn {

/** A graph with one epsilon arc between two nodes carrying `val`. */
Graph scalarGraph(float val, bool calcGrad = true);

/** The emissions chain: M+1 nodes, N arcs (labels 0..N-1) between consecutive nodes. */
Graph linearGraph(int M, int N, bool calcGrad = true);

} // namespace gtn

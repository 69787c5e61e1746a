/* gtn/autograd.h -- reference: gtn/autograd.h:27,37. */
#pragma once

#include <vector>

#include "gtn/graph.h"

namespace gtn {

/** Compute the gradients of every graph `g` depends on, seeding with ones. */
void backward(Graph g, bool retainGraph = false);

/** As above, seeding with `grad` (same structure as `g`). */
void backward(Graph g, const Graph& grad, bool retainGraph = false);

/** The list form (bindings/python/gtn/_autograd.cpp:19-62); see gtn/batched.cpp. */
void backward(const std::vector<Graph>& graphs, bool retainGraph = false);

} // namespace gtn

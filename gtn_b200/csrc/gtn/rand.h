/* gtn/rand.h -- reference: gtn/rand.h:22-40. */
#pragma once

#include <cstddef>

#include "gtn/graph.h"

namespace gtn {

/** A uniformly random walk from a start node to an accept node as a chain graph (empty graph if none found). */
Graph sample(const Graph& g, size_t maxLength = 1000);

/** Monte-Carlo equivalence: sampled paths of either graph get the same score in both. */
bool randEquivalent(const Graph& g1, const Graph& g2, size_t numSamples = 100, double tol = 1e-4,
                    size_t maxLength = 1000);

} // namespace gtn

/*
 * gtn/utils.h -- the subset of the reference's utils the hot path's tests use
 * (reference: gtn/utils.h:23-153): exact graph equality and the text format.
 * isomorphic / binary save-load / draw are outside the scope of this build
 * (SURVEY.md section 2).
 */
#pragma once

#include <iostream>
#include <string>

#include "gtn/graph.h"

namespace gtn {

/** Exact equality: same nodes, same arcs per node (order-insensitive), same weights. */
bool equal(const Graph& a, const Graph& b);

/** Text format: "start ids\naccept ids\nsrc dst ilabel [olabel [weight]]\n..." */
Graph loadTxt(std::istream& in);
Graph loadTxt(const std::string& fileName);
void saveTxt(std::ostream& out, const Graph& g);

} // namespace gtn

/*
 * gtn/utils.h -- the reference's gtn/utils.h:23-153 surface: graph comparison (equal,
 * isomorphic), the binary and text wire formats, stream output and Graphviz export.
 * Host-side tooling around the hot path: the formats are byte/character compatible with
 * the reference so that graphs saved by either library load in the other
 * (tests/test_host_api.py round-trips them through the compiled reference).
 */
#pragma once

#include <iostream>
#include <string>
#include <unordered_map>

#include "gtn/graph.h"

namespace gtn {

/** Exact equality: same nodes, same arcs per node (order-insensitive), same weights. */
bool equal(const Graph& g1, const Graph& g2);

/** Equality up to a renumbering of the nodes (labels, weights, start/accept must agree). */
bool isomorphic(const Graph& g1, const Graph& g2);

/** Binary format (utils.cpp:152-225): int32 counts, start ids, accept ids, 4 x int32 per arc, float32 weights. */
void save(const std::string& fileName, const Graph& g);
void save(std::ostream& out, const Graph& g);
Graph load(const std::string& fileName);
Graph load(std::istream& in);
Graph load(std::istream&& in);

/** Text format: "start ids\naccept ids\nsrc dst ilabel [olabel [weight]]\n..." (utils.cpp:227-345). */
void saveTxt(const std::string& fileName, const Graph& g);
void saveTxt(std::ostream& out, const Graph& g);
Graph loadTxt(const std::string& fileName);
Graph loadTxt(std::istream& in);
Graph loadTxt(std::istream&& in);

/** Text format, abbreviated to 10 entries per section for graphs with more than 20 nodes or arcs. */
std::ostream& operator<<(std::ostream& out, const Graph& g);

/** label -> printable symbol */
using SymbolMap = std::unordered_map<int, std::string>;

/** Graphviz DOT (utils.cpp:383-454). */
void draw(const Graph& g, std::ostream& out, const SymbolMap& isymbols = SymbolMap(),
          const SymbolMap& osymbols = SymbolMap());
void draw(const Graph& g, const std::string& filename, const SymbolMap& isymbols = SymbolMap(),
          const SymbolMap& osymbols = SymbolMap());

} // namespace gtn

/*
 * gtn/utils.cpp -- reference behaviour: gtn/utils.cpp:45-77 (equal), :79-150 (isomorphic),
 * :152-225 (binary format), :227-345 (text format), :378-454 (stream output, Graphviz).
 * Host-side tooling; the wire formats are compatible with the reference's in both directions.
 */
#include "gtn/utils.h"

#include <algorithm>
#include <cstdint>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <tuple>
#include <unordered_map>

namespace gtn {

/* ---- comparison ----------------------------------------------------------- */

bool equal(const Graph& a, const Graph& b) {
  if (a.numNodes() != b.numNodes() || a.numArcs() != b.numArcs() || a.numStart() != b.numStart() ||
      a.numAccept() != b.numAccept()) {
    return false;
  }
  for (size_t n = 0; n < a.numNodes(); n++) {
    if (a.isStart(n) != b.isStart(n) || a.isAccept(n) != b.isAccept(n) || a.numOut(n) != b.numOut(n) ||
        a.numIn(n) != b.numIn(n)) {
      return false;
    }
    // the out-arcs of the node as multisets of (destination, labels, weight)
    auto key = [](const Graph& g, int arc) {
      return std::make_tuple(g.dstNode(arc), g.ilabel(arc), g.olabel(arc), g.weight(arc));
    };
    std::vector<std::tuple<int, int, int, float>> ka, kb;
    for (auto arc : a.out(n)) ka.push_back(key(a, arc));
    for (auto arc : b.out(n)) kb.push_back(key(b, arc));
    std::sort(ka.begin(), ka.end());
    std::sort(kb.begin(), kb.end());
    if (ka != kb) return false;
  }
  return true;
}

namespace {

/*
 * Backtracking search for a label/weight-preserving pairing of out-arcs, node pair by node
 * pair, memoising the verdict per pair like the reference does (utils.cpp:79-131: a pair is
 * assumed good while it is on the search stack, which is what makes cycles terminate).
 */
class IsoSearch {
 public:
  IsoSearch(const Graph& a, const Graph& b) : a_(a), b_(b) {}

  bool pair(int na, int nb) {
    const std::uint64_t key = (static_cast<std::uint64_t>(static_cast<std::uint32_t>(na)) << 32) |
        static_cast<std::uint32_t>(nb);
    auto ins = verdict_.emplace(key, true);
    if (!ins.second) return ins.first->second;
    auto fail = [&]() {
      verdict_[key] = false;
      return false;
    };
    if (a_.numIn(na) != b_.numIn(nb) || a_.numOut(na) != b_.numOut(nb) || a_.isStart(na) != b_.isStart(nb) ||
        a_.isAccept(na) != b_.isAccept(nb)) {
      return fail();
    }
    std::vector<int> candidates(b_.out(nb).begin(), b_.out(nb).end());
    for (int arcA : a_.out(na)) {
      size_t hit = candidates.size();
      for (size_t k = 0; k < candidates.size(); k++) {
        const int arcB = candidates[k];
        if (a_.ilabel(arcA) != b_.ilabel(arcB) || a_.olabel(arcA) != b_.olabel(arcB) ||
            a_.weight(arcA) != b_.weight(arcB)) {
          continue;
        }
        if (pair(a_.dstNode(arcA), b_.dstNode(arcB))) {
          hit = k;
          break;
        }
      }
      if (hit == candidates.size()) return fail();
      candidates.erase(candidates.begin() + hit);
    }
    return true;
  }

 private:
  const Graph& a_;
  const Graph& b_;
  std::unordered_map<std::uint64_t, bool> verdict_;
};

} // namespace

bool isomorphic(const Graph& a, const Graph& b) {
  if (a.numNodes() != b.numNodes() || a.numArcs() != b.numArcs() || a.numStart() != b.numStart() ||
      a.numAccept() != b.numAccept()) {
    return false;
  }
  IsoSearch search(a, b);
  std::vector<int> free(b.start().begin(), b.start().end());
  for (int sa : a.start()) { // every start node of a needs its own start node of b
    size_t hit = free.size();
    for (size_t k = 0; k < free.size() && hit == free.size(); k++) {
      if (search.pair(sa, free[k])) hit = k;
    }
    if (hit == free.size()) return false;
    free.erase(free.begin() + hit);
  }
  return true;
}

/* ---- binary format ---------------------------------------------------------- */

namespace {
template <typename T>
void put(std::ostream& out, const T* p, size_t n) {
  out.write(reinterpret_cast<const char*>(p), static_cast<std::streamsize>(n * sizeof(T)));
}
template <typename T>
void get(std::istream& in, T* p, size_t n) {
  in.read(reinterpret_cast<char*>(p), static_cast<std::streamsize>(n * sizeof(T)));
}
} // namespace

void save(std::ostream& out, const Graph& g) {
  const std::int32_t head[4] = {static_cast<std::int32_t>(g.numNodes()), static_cast<std::int32_t>(g.numArcs()),
                                static_cast<std::int32_t>(g.numStart()), static_cast<std::int32_t>(g.numAccept())};
  put(out, head, 4);
  put(out, g.start().data(), g.start().size());
  put(out, g.accept().data(), g.accept().size());
  std::vector<std::int32_t> rec(4 * g.numArcs());
  for (size_t a = 0; a < g.numArcs(); a++) {
    rec[4 * a + 0] = g.srcNode(a);
    rec[4 * a + 1] = g.dstNode(a);
    rec[4 * a + 2] = g.ilabel(a);
    rec[4 * a + 3] = g.olabel(a);
  }
  put(out, rec.data(), rec.size());
  put(out, g.weights(), g.numArcs());
}

Graph load(std::istream& in) {
  std::int32_t head[4] = {0, 0, 0, 0};
  get(in, head, 4);
  const int numNodes = head[0], numArcs = head[1], numStart = head[2], numAccept = head[3];
  if (!in || numNodes < 0 || numArcs < 0 || numStart < 0 || numAccept < 0) {
    throw std::invalid_argument("[gtn::load] not a graph in binary format");
  }
  std::vector<std::int32_t> start(numStart), accept(numAccept);
  get(in, start.data(), start.size());
  get(in, accept.data(), accept.size());
  std::vector<std::uint8_t> flags(numNodes, 0);
  for (int s : start) {
    if (s >= 0 && s < numNodes) flags[s] |= 1;
  }
  for (int a : accept) {
    if (a >= 0 && a < numNodes) flags[a] |= 2;
  }
  Graph g;
  for (int n = 0; n < numNodes; n++) g.addNode(flags[n] & 1, flags[n] & 2);
  std::vector<std::int32_t> rec(4 * static_cast<size_t>(numArcs));
  get(in, rec.data(), rec.size());
  std::vector<float> w(numArcs);
  get(in, w.data(), w.size());
  if (!in) throw std::invalid_argument("[gtn::load] truncated graph file");
  for (int a = 0; a < numArcs; a++) {
    g.addArc(rec[4 * a], rec[4 * a + 1], rec[4 * a + 2], rec[4 * a + 3], w[a]);
  }
  return g;
}

void save(const std::string& fileName, const Graph& g) {
  std::ofstream out(fileName, std::ios::binary);
  save(out, g);
}

Graph load(const std::string& fileName) {
  std::ifstream in(fileName, std::ios::binary);
  if (!in) throw std::invalid_argument("Couldn't find graph file to load. '" + fileName + "'");
  return load(in);
}

Graph load(std::istream&& in) {
  return load(in);
}

/* ---- text format ------------------------------------------------------------ */

namespace {

/* fields separated by single spaces, exactly (an empty line is one empty field: it fails to
 * parse as a node id, which is how the reference rejects graphs without start/accept nodes) */
std::vector<std::string> fields(const std::string& line) {
  std::vector<std::string> f;
  size_t at = 0;
  for (;;) {
    const size_t sp = line.find(' ', at);
    if (sp == std::string::npos) break;
    f.push_back(line.substr(at, sp - at));
    at = sp + 1;
  }
  f.push_back(line.substr(at));
  return f;
}

std::vector<int> nodeList(const std::string& line, const char* what) {
  std::vector<int> ids;
  for (auto& f : fields(line)) ids.push_back(std::stoi(f)); // std::invalid_argument on junk
  auto sorted = ids;
  std::sort(sorted.begin(), sorted.end());
  if (std::adjacent_find(sorted.begin(), sorted.end()) != sorted.end()) {
    throw std::invalid_argument(std::string("Repeat ") + what + " node detected.");
  }
  return ids;
}

constexpr size_t kSummaryItems = 10; // entries per section in the abbreviated form
constexpr size_t kSummaryAbove = 20; // graphs larger than this are abbreviated by operator<<

void writeTxt(std::ostream& out, const Graph& g, bool brief) {
  auto idLine = [&](const std::vector<int>& ids) {
    for (size_t i = 0; i < ids.size(); i++) {
      if (brief && i >= kSummaryItems) {
        out << " ...";
        break;
      }
      out << (i ? " " : "") << ids[i];
    }
    out << "\n";
  };
  idLine(g.start());
  idLine(g.accept());
  for (size_t a = 0; a < g.numArcs(); a++) {
    if (brief && a >= kSummaryItems) {
      out << "...\n";
      break;
    }
    out << g.srcNode(a) << " " << g.dstNode(a) << " " << g.ilabel(a) << " " << g.olabel(a) << " "
        << g.weight(a) << "\n";
  }
}

} // namespace

Graph loadTxt(std::istream& in) {
  std::string line;
  if (!std::getline(in, line)) throw std::invalid_argument("Must specify start node(s).");
  const auto starts = nodeList(line, "start");
  if (!std::getline(in, line)) throw std::invalid_argument("Must specify accept node(s).");
  const auto accepts = nodeList(line, "accept");

  int maxNode = -1;
  for (int s : starts) maxNode = std::max(maxNode, s);
  for (int a : accepts) maxNode = std::max(maxNode, a);
  std::vector<std::uint8_t> flags(maxNode + 1, 0);
  for (int s : starts) {
    if (s >= 0) flags[s] |= 1;
  }
  for (int a : accepts) {
    if (a >= 0) flags[a] |= 2;
  }
  Graph g;
  for (int n = 0; n <= maxNode; n++) g.addNode(flags[n] & 1, flags[n] & 2);

  while (std::getline(in, line)) {
    const auto f = fields(line);
    if (f.size() < 3 || f.size() > 5) throw std::invalid_argument("Bad line for loading arc.");
    const int src = std::stoi(f[0]), dst = std::stoi(f[1]);
    for (; maxNode < std::max(src, dst); maxNode++) g.addNode();
    const int il = std::stoi(f[2]);
    if (f.size() == 3) {
      g.addArc(src, dst, il);
    } else if (f.size() == 4) {
      g.addArc(src, dst, il, std::stoi(f[3]));
    } else {
      g.addArc(src, dst, il, std::stoi(f[3]), std::stof(f[4]));
    }
  }
  return g;
}

Graph loadTxt(const std::string& fileName) {
  std::ifstream in(fileName);
  if (!in) throw std::invalid_argument("Couldn't find graph file to load. '" + fileName + "'");
  return loadTxt(in);
}

Graph loadTxt(std::istream&& in) {
  return loadTxt(in);
}

void saveTxt(std::ostream& out, const Graph& g) {
  writeTxt(out, g, false);
}

void saveTxt(const std::string& fileName, const Graph& g) {
  std::ofstream out(fileName);
  saveTxt(out, g);
}

std::ostream& operator<<(std::ostream& out, const Graph& g) {
  writeTxt(out, g, std::max(g.numArcs(), g.numNodes()) > kSummaryAbove);
  return out;
}

/* ---- Graphviz ----------------------------------------------------------------- */

void draw(const Graph& g, std::ostream& out, const SymbolMap& isymbols, const SymbolMap& osymbols) {
  auto symbol = [](const SymbolMap& map, int label) -> std::string {
    if (label == epsilon) return "\xCE\xB5"; // U+03B5
    if (map.empty()) return std::to_string(label);
    return map.at(label);
  };
  auto emit = [&](int n) {
    out << "  " << n << " [label = \"" << n << "\", shape = " << (g.isAccept(n) ? "doublecircle" : "circle")
        << ", penwidth = " << (g.isStart(n) ? "2.0" : "1.0") << ", fontsize = 14];\n";
    for (int a : g.out(n)) {
      out << "  " << g.srcNode(a) << " -> " << g.dstNode(a) << " [label = \"" << symbol(isymbols, g.ilabel(a));
      if (!osymbols.empty()) out << ":" << symbol(osymbols, g.olabel(a));
      out << "/" << g.weight(a) << "\", fontsize = 14];\n";
    }
  };
  out << "digraph FST {\n  margin = 0;\n  rankdir = LR;\n  label = \"\";\n"
      << "  center = 1;\n  ranksep = \"0.4\";\n  nodesep = \"0.25\";\n";
  // start nodes first, pure accept nodes last: helps the left-to-right layout
  for (int n : g.start()) emit(n);
  for (size_t n = 0; n < g.numNodes(); n++) {
    if (!g.isStart(n) && !g.isAccept(n)) emit(static_cast<int>(n));
  }
  for (int n : g.accept()) {
    if (!g.isStart(n)) emit(n);
  }
  out << "}";
}

void draw(const Graph& g, const std::string& filename, const SymbolMap& isymbols, const SymbolMap& osymbols) {
  std::ofstream out(filename);
  if (!out.is_open()) throw std::runtime_error("Could not open file [" + filename + "]");
  draw(g, out, isymbols, osymbols);
}

} // namespace gtn

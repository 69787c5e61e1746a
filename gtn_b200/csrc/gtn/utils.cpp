/* gtn/utils.cpp -- reference behaviour: gtn/utils.cpp:45-77 (equal), :227-345 (text format). */
#include "gtn/utils.h"

#include <algorithm>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <tuple>

namespace gtn {

bool equal(const Graph& a, const Graph& b) {
  if (a.numNodes() != b.numNodes() || a.numArcs() != b.numArcs() || a.numStart() != b.numStart() ||
      a.numAccept() != b.numAccept()) {
    return false;
  }
  for (size_t n = 0; n < a.numNodes(); n++) {
    if (a.isStart(n) != b.isStart(n) || a.isAccept(n) != b.isAccept(n) || a.numOut(n) != b.numOut(n)) {
      return false;
    }
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
std::vector<int> intsOf(const std::string& line) {
  std::vector<int> v;
  std::istringstream ss(line);
  int x;
  while (ss >> x) v.push_back(x);
  return v;
}
} // namespace

Graph loadTxt(std::istream& in) {
  std::string line;
  if (!std::getline(in, line)) throw std::logic_error("[gtn::loadTxt] missing start node line");
  auto starts = intsOf(line);
  if (!std::getline(in, line)) throw std::logic_error("[gtn::loadTxt] missing accept node line");
  auto accepts = intsOf(line);
  struct Row {
    int s, d, il, ol;
    float w;
  };
  std::vector<Row> rows;
  int maxNode = -1;
  for (int s : starts) maxNode = std::max(maxNode, s);
  for (int a : accepts) maxNode = std::max(maxNode, a);
  while (std::getline(in, line)) {
    std::istringstream ss(line);
    std::vector<std::string> tok;
    std::string t;
    while (ss >> t) tok.push_back(t);
    if (tok.empty()) continue;
    if (tok.size() < 3 || tok.size() > 5) throw std::logic_error("[gtn::loadTxt] bad arc line");
    Row r;
    r.s = std::stoi(tok[0]);
    r.d = std::stoi(tok[1]);
    r.il = std::stoi(tok[2]);
    r.ol = tok.size() > 3 ? std::stoi(tok[3]) : r.il;
    r.w = tok.size() > 4 ? std::stof(tok[4]) : 0.0f;
    maxNode = std::max({maxNode, r.s, r.d});
    rows.push_back(r);
  }
  Graph g;
  std::vector<uint8_t> flags(maxNode + 1, 0);
  for (int s : starts) flags[s] |= 1;
  for (int a : accepts) flags[a] |= 2;
  for (int n = 0; n <= maxNode; n++) g.addNode(flags[n] & 1, flags[n] & 2);
  for (auto& r : rows) g.addArc(r.s, r.d, r.il, r.ol, r.w);
  return g;
}

Graph loadTxt(const std::string& fileName) {
  std::ifstream in(fileName);
  if (!in.is_open()) throw std::logic_error("[gtn::loadTxt] Can't open file " + fileName);
  return loadTxt(in);
}

void saveTxt(std::ostream& out, const Graph& g) {
  auto list = [&](const std::vector<int>& v) {
    for (size_t i = 0; i < v.size(); i++) out << (i ? " " : "") << v[i];
    out << "\n";
  };
  list(g.start());
  list(g.accept());
  for (size_t a = 0; a < g.numArcs(); a++) {
    out << g.srcNode(a) << " " << g.dstNode(a) << " " << g.ilabel(a) << " " << g.olabel(a) << " "
        << g.weight(a) << "\n";
  }
}

} // namespace gtn

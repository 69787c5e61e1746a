/*
 * gtn/pybind.cpp -- Python bindings with the reference's surface
 * (bindings/python/gtn/_graph.cpp, _functions.cpp, _autograd.cpp, _creations.cpp,
 * _parallel.cpp; flattened into one module like gtn/__init__.py:13-20 does): every
 * function has a single-graph and a list-of-graphs overload, calls release the GIL.
 */
#include <pybind11/functional.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstring>

#include "gtn/gtn.h"

using namespace gtn;
namespace py = pybind11;
using namespace py::literals;

namespace {

template <typename F>
auto unary(F f) {
  return [f](const Graph& g) {
    py::gil_scoped_release release;
    return f(g);
  };
}
template <typename F>
auto unaryList(F f) {
  return [f](const std::vector<Graph>& gs) {
    py::gil_scoped_release release;
    return parallelMap(f, gs);
  };
}
template <typename F>
auto binary(F f) {
  return [f](const Graph& a, const Graph& b) {
    py::gil_scoped_release release;
    return f(a, b);
  };
}
template <typename F>
auto binaryList(F f) {
  return [f](const std::vector<Graph>& a, const std::vector<Graph>& b) {
    py::gil_scoped_release release;
    return parallelMap(f, a, b);
  };
}

} // namespace

PYBIND11_MODULE(_gtn, m) {
  m.attr("epsilon") = epsilon;
  m.attr("__version__") = "0.0.0+b200";

  py::enum_<Projection>(m, "Projection")
      .value("NONE", Projection::NONE)
      .value("INPUT", Projection::INPUT)
      .value("OUTPUT", Projection::OUTPUT);

  py::class_<Graph>(m, "Graph")
      .def(py::init<bool>(), "calc_grad"_a = true)
      .def("add_node", &Graph::addNode, "start"_a = false, "accept"_a = false)
      .def("add_arc", py::overload_cast<size_t, size_t, int>(&Graph::addArc), "src_node"_a, "dst_node"_a, "label"_a)
      .def(
          "add_arc", py::overload_cast<size_t, size_t, int, int, float>(&Graph::addArc), "src_node"_a,
          "dst_node"_a, "ilabel"_a, "olabel"_a, "weight"_a = 0.0)
      .def("num_arcs", &Graph::numArcs)
      .def("num_nodes", &Graph::numNodes)
      .def("num_start", &Graph::numStart)
      .def("num_accept", &Graph::numAccept)
      .def("item", &Graph::item)
      .def("arc_sort", &Graph::arcSort, "olabel"_a = false)
      .def("mark_arc_sorted", &Graph::markArcSorted, "olabel"_a = false)
      .def("ilabel_sorted", &Graph::ilabelSorted)
      .def("olabel_sorted", &Graph::olabelSorted)
      .def("weights", [](Graph& g) { return reinterpret_cast<std::uintptr_t>(g.weights()); })
      .def("weights_to_list", [](Graph& g) { return std::vector<float>(g.weights(), g.weights() + g.numArcs()); })
      .def(
          "weights_to_numpy",
          [](Graph& g) {
            py::array_t<float> out(g.numArcs());
            std::memcpy(out.mutable_data(), g.weights(), sizeof(float) * g.numArcs());
            return out;
          })
      .def(
          "set_weights",
          [](Graph& g, std::uintptr_t ptr) { g.setWeights(reinterpret_cast<const float*>(ptr)); }, "weights"_a)
      .def(
          "set_weights",
          [](Graph& g, const std::vector<float>& w) {
            if (w.size() != g.numArcs()) throw std::invalid_argument("[Graph.set_weights] wrong number of weights");
            g.setWeights(w.data());
          },
          "weights"_a)
      .def(
          "set_weights",
          [](Graph& g, py::array_t<float, py::array::c_style | py::array::forcecast> w) {
            if ((size_t)w.size() != g.numArcs())
              throw std::invalid_argument("[Graph.set_weights] wrong number of weights");
            g.setWeights(w.data());
          },
          "weights"_a)
      .def("labels_to_list", &Graph::labelsToVector, "ilabel"_a = true)
      .def("calc_grad", &Graph::calcGrad)
      .def_property("calc_grad", &Graph::calcGrad, &Graph::setCalcGrad)
      .def("is_grad_available", &Graph::isGradAvailable)
      .def("grad", (Graph & (Graph::*)()) & Graph::grad, py::return_value_policy::reference_internal)
      .def("zero_grad", &Graph::zeroGrad)
      .def("is_start", &Graph::isStart)
      .def("is_accept", &Graph::isAccept)
      .def("make_accept", &Graph::makeAccept)
      .def("src_node", &Graph::srcNode)
      .def("dst_node", &Graph::dstNode)
      .def("ilabel", &Graph::ilabel)
      .def("olabel", &Graph::olabel)
      .def("label", &Graph::label)
      .def("weight", &Graph::weight)
      .def("set_weight", &Graph::setWeight)
      .def("num_in", &Graph::numIn)
      .def("num_out", &Graph::numOut)
      .def("is_linear", &Graph::isLinear)
      .def("is_device_resident", &Graph::isDeviceResident)
      .def("__str__", [](const Graph& g) {
        std::ostringstream os;
        saveTxt(os, g);
        return os.str();
      });

  // functions: single graph + list overloads (bindings/python/gtn/_functions.cpp)
  m.def("negate", unary(negate), "g"_a);
  m.def("negate", unaryList(negate), "graphs"_a);
  m.def("add", binary(add), "g1"_a, "g2"_a);
  m.def("add", binaryList(add), "graphs1"_a, "graphs2"_a);
  m.def("subtract", binary(subtract), "g1"_a, "g2"_a);
  m.def("subtract", binaryList(subtract), "graphs1"_a, "graphs2"_a);
  m.def("compose", binary(compose), "g1"_a, "g2"_a);
  m.def("compose", binaryList(compose), "graphs1"_a, "graphs2"_a);
  m.def("intersect", binary(intersect), "g1"_a, "g2"_a);
  m.def("intersect", binaryList(intersect), "graphs1"_a, "graphs2"_a);
  m.def("forward_score", unary(forwardScore), "g"_a);
  m.def("forward_score", unaryList(forwardScore), "graphs"_a);
  m.def("viterbi_score", unary(viterbiScore), "g"_a);
  m.def("viterbi_score", unaryList(viterbiScore), "graphs"_a);
  m.def("viterbi_path", unary(viterbiPath), "g"_a);
  m.def("viterbi_path", unaryList(viterbiPath), "graphs"_a);
  m.def("project_input", unary(projectInput), "g"_a);
  m.def("project_output", unary(projectOutput), "g"_a);
  m.def(
      "clone",
      [](const Graph& g, Projection p) {
        py::gil_scoped_release release;
        return clone(g, p);
      },
      "g"_a, "projection"_a = Projection::NONE);

  // autograd (bindings/python/gtn/_autograd.cpp)
  m.def(
      "backward",
      [](Graph g, bool retain) {
        py::gil_scoped_release release;
        backward(g, retain);
      },
      "g"_a, "retain_graph"_a = false);
  m.def(
      "backward",
      [](Graph g, const Graph& grad, bool retain) {
        py::gil_scoped_release release;
        backward(g, grad, retain);
      },
      "g"_a, "grad"_a, "retain_graph"_a = false);
  m.def(
      "backward",
      [](const std::vector<Graph>& gs, const std::vector<bool>& retain) {
        py::gil_scoped_release release;
        std::vector<int> idx(gs.size());
        for (size_t i = 0; i < gs.size(); i++) idx[i] = (int)i;
        auto one = [&](int i) { backward(gs[i], retain.size() == 1 ? retain[0] : (bool)retain[i]); };
        parallelMap(one, idx);
      },
      "graphs"_a, "retain_graph"_a = std::vector<bool>({false}));

  // creations (bindings/python/gtn/_creations.cpp)
  m.def("scalar_graph", &scalarGraph, "val"_a, "calc_grad"_a = true);
  m.def("linear_graph", &linearGraph, "M"_a, "N"_a, "calc_grad"_a = true);

  // parallel (bindings/python/gtn/_parallel.cpp)
  m.def(
      "parallel_for",
      [](const std::function<void(int)>& fn, const std::vector<int>& ints) {
        py::gil_scoped_release release;
        parallelMap(fn, ints);
      },
      "function"_a, "int_list"_a);

  // utils subset
  m.def("equal", &equal, "g1"_a, "g2"_a);
  m.def("loads", [](const std::string& s) {
    std::istringstream in(s);
    return loadTxt(in);
  });
}

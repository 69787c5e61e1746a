"""General composition (k_gcompose.cu) against the host construction of the gtn:: layer: wall time per call of
gtn.intersect / gtn.compose for a few operand shapes, single pairs and lists.

    python scripts/bench_gcompose.py > gpurun_out/gcompose.json
"""
import json
import sys
import time

sys.path.insert(0, '.')
import numpy as np

import gtn_b200.gtn as gtn
from tests.test_gpu_gcompose import ctc_graph, transitions_graph


def timed(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


def lexicon(rng, words, letters, max_len):
    """A trie-less lexicon transducer: one chain per word from the shared start node, letters in, word id out on
    the first arc, epsilon out on the rest; every chain returns to the start node with an epsilon:epsilon arc."""
    g = gtn.Graph(False)
    g.add_node(True, True)
    for wd in range(words):
        prev = 0
        n = int(rng.integers(2, max_len + 1))
        for k in range(n):
            cur = g.add_node()
            g.add_arc(prev, cur, int(rng.integers(0, letters)), wd if k == 0 else gtn.epsilon, 0.0)
            prev = cur
        g.add_arc(prev, 0, gtn.epsilon, gtn.epsilon, 0.0)
    return g


def word_lm(rng, words, fanout):
    g = gtn.Graph(False)
    n = 64
    for i in range(n):
        g.add_node(i == 0, True)
    for i in range(n):
        for wd in rng.choice(words, fanout, replace=False):
            g.add_arc(i, int(rng.integers(0, n)), int(wd), int(wd), float(rng.uniform(-1, 0)))
    g.arc_sort()
    return g


rng = np.random.default_rng(0)  # (module level: scripts/prof_gcompose.py imports `cases`)
cases = {}
ctc10 = ctc_graph(gtn, rng.integers(1, 30, 10))
cases["ngram_ctc U=10 x bigram M=30 (benchmarks/ctc.cpp:107-123)"] = (ctc10, transitions_graph(gtn, 30, 2, rng), True)
ctc100 = ctc_graph(gtn, rng.integers(1, 30, 100))
cases["ctc U=100 x trigram M=30 (900 states, 27000 arcs)"] = (ctc100, transitions_graph(gtn, 30, 3, rng), True)
lex = lexicon(rng, 2000, 28, 8)
cases["lexicon 2000 words x word LM (64 states, fan-out 200)"] = (lex, word_lm(rng, 2000, 200), False)

def eps_graph(rng, n, C):
    """a left-to-right graph with self-loops and a few epsilon-output arcs (not frame-synchronous)"""
    g = gtn.Graph(False)
    for i in range(n):
        g.add_node(i == 0, i == n - 1)
    for i in range(n - 1):
        g.add_arc(i, i + 1, int(rng.integers(0, C)), int(rng.integers(0, C)), 0.0)
        g.add_arc(i, i, int(rng.integers(0, C)), int(rng.integers(0, C)), 0.0)
        if i % 5 == 0:
            g.add_arc(i, i + 1, int(rng.integers(0, C)), gtn.epsilon, 0.0)
    return g


em = gtn.linear_graph(1000, 64, False)
em.set_weights(rng.uniform(-5, 5, 64000).astype(np.float32).tolist())
cases["201-node graph with epsilon outputs x emissions T=1000 C=64"] = (eps_graph(rng, 201, 64), em, False)


def main():
  out = {}
  for name, (a, b, inter) in cases.items():
      op = gtn.intersect if inter else gtn.compose
      row = {"states": a.num_nodes() * b.num_nodes()}
      for pol, key in ((2, "host_ms"), (1, "device_ms")):
          gtn.set_compose_device_policy(pol)
          reps = 20 if key == "device_ms" else 3
          row[key] = timed(lambda: op(a, b), reps)
          r = op(a, b)
          row["nodes"], row["arcs"] = r.num_nodes(), r.num_arcs()
      # a list of 32 such pairs: one batched device call against parallelMap of the host construction
      for pol, key in ((2, "host_list32_ms"), (1, "device_list32_ms")):
          gtn.set_compose_device_policy(pol)
          row[key] = timed(lambda: op([a] * 32, [b] * 32), 3)
      out[name] = row
      print(name, row, file=sys.stderr)
  gtn.set_compose_device_policy(0)
  print(json.dumps(out, indent=1))


if __name__ == '__main__':
  main()

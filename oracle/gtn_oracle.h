/*
 * oracle/gtn_oracle.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Plain-C restatement of the reference's hot path so the CUDA path can be
 * checked on the GPU box (where /root/reference does not exist):
 *   compose / intersect        gtn/functions/compose.cpp:377-522, functions.cpp:225-251
 *   shortestDistance (+grad)   gtn/functions/shortest.cpp:33-188
 *   shortestPath               gtn/functions/shortest.cpp:190-272
 *   arcSort / linearGraph      gtn/graph.cpp:162-177, gtn/creations.cpp:20-33
 *   CTC / ASG criteria         benchmarks/ctc.cpp:40-58,150-160, test/criterion_test.cpp:244-278
 *
 * Parity pinned: tests/test_oracle.py checks this file against the
 * real reference (oracle/_ref/libgtn_ref.so) on the reference's own known-
 * answer tests (test/criterion_test.cpp, test/functions_test.cpp,
 * test/autograd_test.cpp) and on random graphs; tests/golden/ holds fixtures
 * generated from the real reference.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 */
#ifndef GTN_ORACLE_H
#define GTN_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OG_EPSILON (-1)
#define OG_OK 0
#define OG_ERR_INVALID (-2) /* std::invalid_argument in the reference */
#define OG_ERR_LOGIC (-3) /* std::logic_error */

typedef struct og_graph {
  int num_nodes, num_arcs;
  int cap_nodes, cap_arcs;
  uint8_t* flags; /* bit0 start, bit1 accept */
  int *src, *dst, *ilabel, *olabel;
  float* w;
  /* per-node adjacency in the reference's order (insertion, then arcSort) */
  int **in, **out;
  int *nin, *nout, *cin, *cout;
  int *start, num_start, cap_start;
  int *accept, num_accept, cap_accept;
  int ilabel_sorted, olabel_sorted;
  /* compose provenance (gradInfo, compose.cpp:407,445); NULL otherwise */
  int *gi1, *gi2;
} og_graph;

og_graph* og_new(void);
void og_free(og_graph* g);
int og_add_node(og_graph* g, int start, int accept);
int og_add_arc(og_graph* g, int src, int dst, int il, int ol, float w);
og_graph* og_from_arrays(
    int num_nodes,
    const uint8_t* flags,
    int num_arcs,
    const int* src,
    const int* dst,
    const int* il,
    const int* ol,
    const float* w);
og_graph* og_linear(int T, int C, const float* w);
void og_arc_sort(og_graph* g, int olabel);
/* flat CSR views of the adjacency, for ctypes */
void og_get_adj(const og_graph* g, int* in_ptr, int* in_arcs, int* out_ptr, int* out_arcs);

/* mode 0 = compose (functions.cpp:225-237), 1 = intersect (:239-251) */
og_graph* og_compose(const og_graph* a, const og_graph* b, int mode);

/*
 * shortestDistance.  scores[N], max_cache[N+1], arg_cache[N+1] (arc index,
 * -1 when unset / start; last entry = best accept NODE) are caller buffers
 * (any may be NULL).  Returns OG_OK or OG_ERR_INVALID.
 */
int og_shortest_distance(
    const og_graph* g,
    int tropical,
    float* out,
    float* scores,
    float* max_cache,
    int64_t* arg_cache);

/* shortestDistanceGrad: arc_grads[A] written (not accumulated). */
int og_shortest_distance_grad(
    const og_graph* g,
    int tropical,
    float output,
    float delta,
    const float* scores,
    const float* max_cache,
    const int64_t* arg_cache,
    float* arc_grads);

/*
 * shortestPath: path_arcs (start -> end order) must hold num_nodes entries.
 * *path_len = -1 when there is no accepting path (empty result graph).
 */
int og_shortest_path(const og_graph* g, int* path_arcs, int* path_len);

/* compose gradFunc (compose.cpp:496-518): scatter deltas[A] via gradInfo. */
void og_compose_grad(
    const og_graph* composed,
    const float* deltas,
    float* grad1, /* nullable, accumulated into */
    float* grad2);

/* criteria built from the primitives above */
og_graph* og_ctc_graph(const int* target, int U, int blank, int arc_sort);
og_graph* og_transitions_graph(int C, const float* w);
og_graph* og_fal_graph(const int* target, int U);

/* loss = fwd(emissions) - fwd(intersect(ctc, emissions)); grad nullable */
int og_ctc_loss(
    int T,
    int C,
    const float* emissions,
    const int* target,
    int U,
    int blank,
    int arc_sort,
    float* loss,
    float* grad);

/* loss = fwd(compose(e,trans)) - fwd(compose(compose(fal,trans),e)) */
int og_asg_loss(
    int T,
    int C,
    const float* emissions,
    const float* trans_w,
    const int* target,
    int U,
    float* loss,
    float* grad,
    float* trans_grad /* accumulated into, nullable */);

/* viterbiPath(compose(e, trans)) ilabels -> path[T]; score = viterbiScore */
int og_viterbi_dense(
    int T,
    int C,
    const float* emissions,
    const float* trans_w,
    int* path,
    float* score);

/* viterbiPath(intersect(ctc, e)) ilabels -> path[T] */
int og_viterbi_ctc(
    int T,
    int C,
    const float* emissions,
    const int* target,
    int U,
    int blank,
    int arc_sort,
    int* path,
    float* score);

#ifdef __cplusplus
}
#endif
#endif

/*
 * oracle/gtn_oracle.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 * See gtn_oracle.h for scope, provenance and who may load this.
 *
 * Every function cites the reference file:line whose algorithm it restates.
 * Arithmetic is IEEE float32 with libm expf/log1pf, like the reference's
 * std::exp(float)/std::log1p(float) (shortest.cpp:111,113).
 */
#include "gtn_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* containers                                                          */
/* ------------------------------------------------------------------ */

static void* xrealloc(void* p, size_t n) {
  void* q = realloc(p, n ? n : 1);
  if (!q) abort();
  return q;
}

static void push_int(int** v, int* n, int* cap, int x) {
  if (*n == *cap) {
    *cap = *cap ? *cap * 2 : 4;
    *v = (int*)xrealloc(*v, sizeof(int) * (size_t)*cap);
  }
  (*v)[(*n)++] = x;
}

og_graph* og_new(void) {
  return (og_graph*)calloc(1, sizeof(og_graph));
}

void og_free(og_graph* g) {
  if (!g) return;
  for (int n = 0; n < g->num_nodes; n++) {
    free(g->in[n]);
    free(g->out[n]);
  }
  free(g->flags);
  free(g->src);
  free(g->dst);
  free(g->ilabel);
  free(g->olabel);
  free(g->w);
  free(g->in);
  free(g->out);
  free(g->nin);
  free(g->nout);
  free(g->cin);
  free(g->cout);
  free(g->start);
  free(g->accept);
  free(g->gi1);
  free(g->gi2);
  free(g);
}

/* Graph::addNode, graph.cpp:33-45 */
int og_add_node(og_graph* g, int start, int accept) {
  int idx = g->num_nodes;
  if (idx == g->cap_nodes) {
    int c = g->cap_nodes ? g->cap_nodes * 2 : 16;
    g->flags = (uint8_t*)xrealloc(g->flags, (size_t)c);
    g->in = (int**)xrealloc(g->in, sizeof(int*) * (size_t)c);
    g->out = (int**)xrealloc(g->out, sizeof(int*) * (size_t)c);
    g->nin = (int*)xrealloc(g->nin, sizeof(int) * (size_t)c);
    g->nout = (int*)xrealloc(g->nout, sizeof(int) * (size_t)c);
    g->cin = (int*)xrealloc(g->cin, sizeof(int) * (size_t)c);
    g->cout = (int*)xrealloc(g->cout, sizeof(int) * (size_t)c);
    g->cap_nodes = c;
  }
  g->flags[idx] = (uint8_t)((start ? 1 : 0) | (accept ? 2 : 0));
  g->in[idx] = g->out[idx] = NULL;
  g->nin[idx] = g->nout[idx] = g->cin[idx] = g->cout[idx] = 0;
  g->num_nodes++;
  if (start) push_int(&g->start, &g->num_start, &g->cap_start, idx);
  if (accept) push_int(&g->accept, &g->num_accept, &g->cap_accept, idx);
  g->ilabel_sorted = g->olabel_sorted = 0;
  return idx;
}

/* Graph::addArc, graph.cpp:51-67 */
int og_add_arc(og_graph* g, int src, int dst, int il, int ol, float w) {
  int idx = g->num_arcs;
  if (idx == g->cap_arcs) {
    int c = g->cap_arcs ? g->cap_arcs * 2 : 64;
    g->src = (int*)xrealloc(g->src, sizeof(int) * (size_t)c);
    g->dst = (int*)xrealloc(g->dst, sizeof(int) * (size_t)c);
    g->ilabel = (int*)xrealloc(g->ilabel, sizeof(int) * (size_t)c);
    g->olabel = (int*)xrealloc(g->olabel, sizeof(int) * (size_t)c);
    g->w = (float*)xrealloc(g->w, sizeof(float) * (size_t)c);
    g->cap_arcs = c;
  }
  g->src[idx] = src;
  g->dst[idx] = dst;
  g->ilabel[idx] = il;
  g->olabel[idx] = ol;
  g->w[idx] = w;
  g->num_arcs++;
  push_int(&g->out[src], &g->nout[src], &g->cout[src], idx);
  push_int(&g->in[dst], &g->nin[dst], &g->cin[dst], idx);
  g->ilabel_sorted = g->olabel_sorted = 0;
  return idx;
}

og_graph* og_from_arrays(
    int num_nodes,
    const uint8_t* flags,
    int num_arcs,
    const int* src,
    const int* dst,
    const int* il,
    const int* ol,
    const float* w) {
  og_graph* g = og_new();
  for (int n = 0; n < num_nodes; n++) og_add_node(g, flags[n] & 1, flags[n] & 2);
  for (int a = 0; a < num_arcs; a++)
    og_add_arc(g, src[a], dst[a], il[a], ol[a], w ? w[a] : 0.0f);
  return g;
}

/* linearGraph, creations.cpp:20-33: arc (m-1)*N + n, label n; both sorted
 * flags set.  Weights (nullable) are the row-major [T][C] matrix
 * (Graph::setWeights, graph.cpp:179-181). */
og_graph* og_linear(int T, int C, const float* w) {
  og_graph* g = og_new();
  og_add_node(g, 1, T == 0);
  for (int m = 1; m <= T; m++) {
    og_add_node(g, 0, m == T);
    for (int n = 0; n < C; n++)
      og_add_arc(g, m - 1, m, n, n, w ? w[(size_t)(m - 1) * C + n] : 0.0f);
  }
  g->ilabel_sorted = g->olabel_sorted = 1;
  return g;
}

/* Graph::arcSort, graph.cpp:162-177.  The reference uses std::sort; for the
 * <=16-element lists of every graph on the hot path libstdc++'s std::sort is
 * a plain insertion sort, which is what is restated here (stable). */
static void sort_by_label(int* v, int n, const int* label) {
  for (int i = 1; i < n; i++) {
    int x = v[i], j = i - 1;
    while (j >= 0 && label[v[j]] > label[x]) {
      v[j + 1] = v[j];
      j--;
    }
    v[j + 1] = x;
  }
}

void og_arc_sort(og_graph* g, int olabel) {
  if ((olabel && g->olabel_sorted) || (!olabel && g->ilabel_sorted)) return;
  g->olabel_sorted = olabel ? 1 : 0;
  g->ilabel_sorted = olabel ? 0 : 1;
  const int* lab = olabel ? g->olabel : g->ilabel;
  for (int n = 0; n < g->num_nodes; n++) {
    sort_by_label(g->in[n], g->nin[n], lab);
    sort_by_label(g->out[n], g->nout[n], lab);
  }
}

void og_get_adj(const og_graph* g, int* in_ptr, int* in_arcs, int* out_ptr, int* out_arcs) {
  int pi = 0, po = 0;
  for (int n = 0; n < g->num_nodes; n++) {
    in_ptr[n] = pi;
    out_ptr[n] = po;
    for (int k = 0; k < g->nin[n]; k++) in_arcs[pi++] = g->in[n][k];
    for (int k = 0; k < g->nout[n]; k++) out_arcs[po++] = g->out[n][k];
  }
  in_ptr[g->num_nodes] = pi;
  out_ptr[g->num_nodes] = po;
}

/* ------------------------------------------------------------------ */
/* arc matchers, compose.cpp:211-374                                   */
/* ------------------------------------------------------------------ */

enum { M_UNSORTED = 0, M_SINGLY = 1, M_DOUBLY = 2 };

typedef struct {
  const og_graph *g1, *g2;
  int kind, search_g1_cfg, search_g1;
  /* unsorted */
  const int *lv, *rv;
  int nl, nr, li, ri;
  /* sorted: query list, search list, cursors */
  const int *q, *s;
  int nq, ns, qi, si, sb;
} matcher;

static int qlabel(const matcher* m, int arc) {
  return m->search_g1 ? m->g2->ilabel[arc] : m->g1->olabel[arc];
}
static int slabel(const matcher* m, int arc) {
  return m->search_g1 ? m->g1->olabel[arc] : m->g2->ilabel[arc];
}

/* std::lower_bound over s[lo, hi) on the search label */
static int lower_bound_label(const matcher* m, int lo, int hi, int val) {
  while (lo < hi) {
    int mid = lo + (hi - lo) / 2;
    if (slabel(m, m->s[mid]) < val)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}

static void m_match(matcher* m, int lnode, int rnode, int match_in) {
  const int* lv = match_in ? m->g1->in[lnode] : m->g1->out[lnode];
  int nl = match_in ? m->g1->nin[lnode] : m->g1->nout[lnode];
  const int* rv = match_in ? m->g2->in[rnode] : m->g2->out[rnode];
  int nr = match_in ? m->g2->nin[rnode] : m->g2->nout[rnode];
  if (m->kind == M_UNSORTED) { /* compose.cpp:211-218 */
    m->lv = lv;
    m->nl = nl;
    m->rv = rv;
    m->nr = nr;
    m->li = m->ri = 0;
    return;
  }
  /* compose.cpp:244-259 (singly), :307-323 (doubly) */
  m->search_g1 = (m->kind == M_DOUBLY) ? (nl > nr) : m->search_g1_cfg;
  if (m->search_g1) {
    m->s = lv;
    m->ns = nl;
    m->q = rv;
    m->nq = nr;
  } else {
    m->s = rv;
    m->ns = nr;
    m->q = lv;
    m->nq = nl;
  }
  m->qi = m->si = m->sb = 0;
}

static int m_has_next(matcher* m) {
  if (m->kind == M_UNSORTED) { /* compose.cpp:220-230 */
    for (; m->li < m->nl; m->li++) {
      for (; m->ri < m->nr; m->ri++) {
        if (m->g1->olabel[m->lv[m->li]] == m->g2->ilabel[m->rv[m->ri]]) return 1;
      }
      m->ri = 0;
    }
    return 0;
  }
  /* compose.cpp:261-296 (singly), :325-366 (doubly) */
  if (m->qi == m->nq) return 0;
  if (m->si != m->ns) {
    if (qlabel(m, m->q[m->qi]) == slabel(m, m->s[m->si])) return 1;
  }
  if (m->si != m->sb) m->qi++;
  for (; m->qi < m->nq; m->qi++) {
    int ql = qlabel(m, m->q[m->qi]);
    if (m->kind == M_SINGLY) {
      m->si = lower_bound_label(m, m->sb, m->ns, ql);
      if (m->si == m->ns) continue;
      if (slabel(m, m->s[m->si]) == ql) return 1;
    } else {
      m->sb = lower_bound_label(m, m->sb, m->ns, ql);
      if (m->sb == m->ns) return 0;
      if (slabel(m, m->s[m->sb]) == ql) {
        m->si = m->sb;
        return 1;
      }
    }
  }
  return 0;
}

static void m_next(matcher* m, int* i, int* j) {
  if (m->kind == M_UNSORTED) { /* compose.cpp:232-234 */
    *i = m->lv[m->li];
    *j = m->rv[m->ri++];
    return;
  }
  if (m->search_g1) { /* compose.cpp:298-304, :368-374 */
    *i = m->s[m->si++];
    *j = m->q[m->qi];
  } else {
    *i = m->q[m->qi];
    *j = m->s[m->si++];
  }
}

/* ------------------------------------------------------------------ */
/* compose, compose.cpp:377-522                                        */
/* ------------------------------------------------------------------ */

typedef struct {
  int *a, *b;
  size_t head, n, cap;
} pairq;

static void pq_push(pairq* q, int x, int y) {
  if (q->n == q->cap) {
    q->cap = q->cap ? q->cap * 2 : 256;
    q->a = (int*)xrealloc(q->a, sizeof(int) * q->cap);
    q->b = (int*)xrealloc(q->b, sizeof(int) * q->cap);
  }
  q->a[q->n] = x;
  q->b[q->n] = y;
  q->n++;
}

/* epsilonReachable, compose.cpp:21-53 */
static void eps_reachable(
    int second_or_first,
    const og_graph* first,
    const og_graph* second,
    int n1,
    int n2,
    uint8_t* reach,
    pairq* q) {
  const int* edges = second_or_first ? second->in[n2] : first->in[n1];
  int ne = second_or_first ? second->nin[n2] : first->nin[n1];
  int sorted = second_or_first ? second->ilabel_sorted : first->olabel_sorted;
  for (int k = 0; k < ne; k++) {
    int i = edges[k];
    int label = second_or_first ? second->ilabel[i] : first->olabel[i];
    if (label != OG_EPSILON) {
      if (sorted) break;
      continue;
    }
    int un = second_or_first ? second->src[i] : first->src[i];
    size_t idx = second_or_first
        ? (size_t)n1 + (size_t)first->num_nodes * (size_t)un
        : (size_t)un + (size_t)first->num_nodes * (size_t)n2;
    if (!reach[idx]) {
      if (second_or_first)
        pq_push(q, n1, un);
      else
        pq_push(q, un, n2);
    }
    reach[idx] = 1;
  }
}

/* addReachableNodeAndArc, compose.cpp:108-136 */
static int add_reachable(
    const og_graph* first,
    const og_graph* second,
    int cur,
    int d1,
    int d2,
    float w,
    int il,
    int ol,
    const uint8_t* reach,
    pairq* q,
    int* new_nodes,
    og_graph* ng) {
  size_t idx = (size_t)d1 + (size_t)first->num_nodes * (size_t)d2;
  if (reach[idx]) {
    if (new_nodes[idx] < 0) {
      new_nodes[idx] = og_add_node(
          ng,
          (first->flags[d1] & 1) && (second->flags[d2] & 1),
          (first->flags[d1] & 2) && (second->flags[d2] & 2));
      pq_push(q, d1, d2);
    }
    og_add_arc(ng, cur, new_nodes[idx], il, ol, w);
  }
  return reach[idx];
}

typedef struct {
  int *a, *b, n, cap;
} gradinfo;

static void gi_push(gradinfo* g, int i, int j) {
  if (g->n == g->cap) {
    g->cap = g->cap ? g->cap * 2 : 256;
    g->a = (int*)xrealloc(g->a, sizeof(int) * (size_t)g->cap);
    g->b = (int*)xrealloc(g->b, sizeof(int) * (size_t)g->cap);
  }
  g->a[g->n] = i;
  g->b[g->n] = j;
  g->n++;
}

/* addEpsilonReachableNodes, compose.cpp:146-208 */
static void add_eps_nodes(
    int second_or_first,
    const og_graph* first,
    const og_graph* second,
    int cur,
    int n1,
    int n2,
    const uint8_t* reach,
    pairq* q,
    int* new_nodes,
    og_graph* ng,
    gradinfo* gi) {
  const int* edges = second_or_first ? second->out[n2] : first->out[n1];
  int ne = second_or_first ? second->nout[n2] : first->nout[n1];
  int sorted = second_or_first ? second->ilabel_sorted : first->olabel_sorted;
  for (int k = 0; k < ne; k++) {
    int i = edges[k];
    int label = second_or_first ? second->ilabel[i] : first->olabel[i];
    if (label != OG_EPSILON) {
      if (sorted) break;
      continue;
    }
    int ok = add_reachable(
        first,
        second,
        cur,
        second_or_first ? n1 : first->dst[i],
        second_or_first ? second->dst[i] : n2,
        second_or_first ? second->w[i] : first->w[i],
        second_or_first ? OG_EPSILON : first->ilabel[i],
        second_or_first ? second->olabel[i] : OG_EPSILON,
        reach,
        q,
        new_nodes,
        ng);
    if (ok) {
      if (second_or_first)
        gi_push(gi, -1, i);
      else
        gi_push(gi, i, -1);
    }
  }
}

og_graph* og_compose(const og_graph* first, const og_graph* second, int mode) {
  matcher m;
  memset(&m, 0, sizeof(m));
  m.g1 = first;
  m.g2 = second;
  /* matcher choice: functions.cpp:225-237 (compose), :239-251 (intersect) */
  int s1 = mode ? (first->ilabel_sorted || first->olabel_sorted) : first->olabel_sorted;
  int s2 = mode ? (second->ilabel_sorted || second->olabel_sorted) : second->ilabel_sorted;
  if (s1 && s2) {
    m.kind = M_DOUBLY;
  } else if (s1 || s2) {
    m.kind = M_SINGLY;
    m.search_g1_cfg = s1;
  } else {
    m.kind = M_UNSORTED;
  }

  size_t np = (size_t)first->num_nodes * (size_t)second->num_nodes;
  uint8_t* reach = (uint8_t*)calloc(np ? np : 1, 1);
  pairq q;
  memset(&q, 0, sizeof(q));

  /* findReachable, compose.cpp:64-104 */
  for (int f = 0; f < first->num_accept; f++) {
    for (int s = 0; s < second->num_accept; s++) {
      int a1 = first->accept[f], a2 = second->accept[s];
      pq_push(&q, a1, a2);
      reach[(size_t)a1 + (size_t)first->num_nodes * (size_t)a2] = 1;
    }
  }
  while (q.head < q.n) {
    int c1 = q.a[q.head], c2 = q.b[q.head];
    q.head++;
    m_match(&m, c1, c2, 1);
    while (m_has_next(&m)) {
      int i, j;
      m_next(&m, &i, &j);
      int u1 = first->src[i], u2 = second->src[j];
      size_t idx = (size_t)u1 + (size_t)first->num_nodes * (size_t)u2;
      if (!reach[idx]) pq_push(&q, u1, u2);
      reach[idx] = 1;
    }
    eps_reachable(0, first, second, c1, c2, reach, &q);
    eps_reachable(1, first, second, c1, c2, reach, &q);
  }

  /* forward construction, compose.cpp:389-489 */
  og_graph* ng = og_new();
  int* new_nodes = (int*)malloc(sizeof(int) * (np ? np : 1));
  for (size_t i = 0; i < np; i++) new_nodes[i] = -1;
  q.head = q.n = 0;
  for (int a = 0; a < first->num_start; a++) {
    for (int b = 0; b < second->num_start; b++) {
      int s1n = first->start[a], s2n = second->start[b];
      size_t idx = (size_t)s1n + (size_t)first->num_nodes * (size_t)s2n;
      if (reach[idx]) {
        new_nodes[idx] =
            og_add_node(ng, 1, (first->flags[s1n] & 2) && (second->flags[s2n] & 2));
        pq_push(&q, s1n, s2n);
      }
    }
  }
  gradinfo gi;
  memset(&gi, 0, sizeof(gi));
  while (q.head < q.n) {
    int c1 = q.a[q.head], c2 = q.b[q.head];
    q.head++;
    int cur = new_nodes[(size_t)c1 + (size_t)first->num_nodes * (size_t)c2];
    int eps_matched = 0;
    m_match(&m, c1, c2, 0);
    while (m_has_next(&m)) {
      int i, j;
      m_next(&m, &i, &j);
      if (first->olabel[i] == OG_EPSILON) {
        eps_matched = 1;
        continue;
      }
      int ok = add_reachable(
          first,
          second,
          cur,
          first->dst[i],
          second->dst[j],
          first->w[i] + second->w[j],
          first->ilabel[i],
          second->olabel[j],
          reach,
          &q,
          new_nodes,
          ng);
      if (ok) gi_push(&gi, i, j);
    }
    if (!eps_matched || (second->flags[c2] & 2) || !(first->flags[c1] & 2)) {
      add_eps_nodes(0, first, second, cur, c1, c2, reach, &q, new_nodes, ng, &gi);
    }
    if (!eps_matched || (first->flags[c1] & 2)) {
      add_eps_nodes(1, first, second, cur, c1, c2, reach, &q, new_nodes, ng, &gi);
    }
  }
  ng->gi1 = gi.a;
  ng->gi2 = gi.b;
  if (!gi.a) {
    ng->gi1 = (int*)calloc(1, sizeof(int));
    ng->gi2 = (int*)calloc(1, sizeof(int));
  }
  free(reach);
  free(new_nodes);
  free(q.a);
  free(q.b);
  return ng;
}

/* compose gradFunc, compose.cpp:496-518 */
void og_compose_grad(const og_graph* c, const float* deltas, float* grad1, float* grad2) {
  for (int k = 0; k < c->num_arcs; k++) {
    if (grad1 && c->gi1[k] >= 0) grad1[c->gi1[k]] += deltas[k];
    if (grad2 && c->gi2[k] >= 0) grad2[c->gi2[k]] += deltas[k];
  }
}

/* ------------------------------------------------------------------ */
/* shortest distance / path, shortest.cpp                              */
/* ------------------------------------------------------------------ */

/* getScore, shortest.cpp:102-114 */
static float get_score(const float* in, int n, float max_score, int tropical) {
  if (n == 0) return -INFINITY;
  if (tropical || max_score == INFINITY || max_score == -INFINITY) return max_score;
  float score = -1.0f;
  for (int i = 0; i < n; i++) score += expf(in[i] - max_score);
  return max_score + log1pf(score);
}

/* shortestDistance, shortest.cpp:86-188 */
int og_shortest_distance(
    const og_graph* g,
    int tropical,
    float* out,
    float* scores_out,
    float* max_cache_out,
    int64_t* arg_cache_out) {
  int N = g->num_nodes;
  float* scores = (float*)calloc((size_t)N + 1, sizeof(float));
  float* maxc = (float*)malloc(sizeof(float) * ((size_t)N + 1));
  int64_t* argc = (int64_t*)malloc(sizeof(int64_t) * ((size_t)N + 1));
  int* deg = (int*)malloc(sizeof(int) * ((size_t)N + 1));
  int* queue = (int*)malloc(sizeof(int) * ((size_t)N + 1));
  int maxin = 1;
  for (int n = 0; n < N; n++) {
    maxc[n] = -INFINITY;
    argc[n] = -1;
    deg[n] = g->nin[n];
    if (g->nin[n] + 1 > maxin) maxin = g->nin[n] + 1;
  }
  maxc[N] = -INFINITY;
  argc[N] = -1;
  if (g->num_accept + 1 > maxin) maxin = g->num_accept + 1;
  float* ins = (float*)malloc(sizeof(float) * (size_t)maxin);
  int qh = 0, qn = 0;
  for (int k = 0; k < g->num_start; k++)
    if (g->nin[g->start[k]] == 0) queue[qn++] = g->start[k];

  while (qh < qn) {
    int n = queue[qh++];
    int ni = 0;
    for (int k = 0; k < g->nin[n]; k++) {
      int a = g->in[n][k];
      ins[ni] = scores[g->src[a]] + g->w[a];
      if (ins[ni] > maxc[n]) {
        maxc[n] = ins[ni];
        argc[n] = a;
      }
      ni++;
    }
    if (g->flags[n] & 1) {
      ins[ni] = 0.0f;
      if (ins[ni] > maxc[n]) {
        maxc[n] = ins[ni];
        argc[n] = -1;
      }
      ni++;
    }
    scores[n] = get_score(ins, ni, maxc[n], tropical);
    for (int k = 0; k < g->nout[n]; k++) {
      int dn = g->dst[g->out[n][k]];
      if (--deg[dn] == 0) queue[qn++] = dn;
    }
  }

  int rc = OG_OK;
  int ni = 0;
  for (int k = 0; k < g->num_accept; k++) {
    int n = g->accept[k];
    if (deg[n] > 0) {
      rc = OG_ERR_INVALID; /* shortest.cpp:149-152 */
      break;
    }
    ins[ni] = scores[n];
    if (ins[ni] > maxc[N]) {
      maxc[N] = ins[ni];
      argc[N] = n;
    }
    ni++;
  }
  if (rc == OG_OK) {
    *out = get_score(ins, ni, maxc[N], tropical);
    if (scores_out) memcpy(scores_out, scores, sizeof(float) * (size_t)N);
    if (max_cache_out) memcpy(max_cache_out, maxc, sizeof(float) * ((size_t)N + 1));
    if (arg_cache_out) memcpy(arg_cache_out, argc, sizeof(int64_t) * ((size_t)N + 1));
  }
  free(scores);
  free(maxc);
  free(argc);
  free(deg);
  free(queue);
  free(ins);
  return rc;
}

/* shortestDistanceGrad, shortest.cpp:33-82 */
int og_shortest_distance_grad(
    const og_graph* g,
    int tropical,
    float output,
    float delta,
    const float* scores,
    const float* maxc,
    const int64_t* argc,
    float* arc_grads) {
  int N = g->num_nodes;
  int* deg = (int*)malloc(sizeof(int) * ((size_t)N + 1));
  int* queue = (int*)malloc(sizeof(int) * ((size_t)N + 1));
  float* ngrad = (float*)calloc((size_t)N + 1, sizeof(float));
  for (int a = 0; a < g->num_arcs; a++) arc_grads[a] = 0.0f;
  for (int n = 0; n < N; n++) deg[n] = g->nout[n];
  int qh = 0, qn = 0;
  float cur = 0.0f;
  float denom = tropical ? 0.0f : expf(output - maxc[N]);
  for (int k = 0; k < g->num_accept; k++) {
    int n = g->accept[k];
    if (g->nout[n] == 0) queue[qn++] = n;
    if (tropical)
      cur = ((int64_t)n == argc[N]) ? 1.0f : 0.0f;
    else
      cur = expf(scores[n] - maxc[N]) / denom;
    ngrad[n] += cur;
  }
  while (qh < qn) {
    int n = queue[qh++];
    denom = tropical ? 0.0f : expf(scores[n] - maxc[n]);
    for (int k = 0; k < g->nin[n]; k++) {
      int a = g->in[n][k];
      int un = g->src[a];
      if (tropical)
        cur = ((int64_t)a == argc[n]) ? ngrad[n] : 0.0f;
      else
        cur = ngrad[n] * expf(scores[un] + g->w[a] - maxc[n]) / denom;
      ngrad[un] += cur;
      arc_grads[a] = cur * delta;
      if (--deg[un] == 0) queue[qn++] = un;
    }
  }
  free(deg);
  free(queue);
  free(ngrad);
  return OG_OK;
}

/* shortestPath, shortest.cpp:190-272 */
int og_shortest_path(const og_graph* g, int* path_arcs, int* path_len) {
  int N = g->num_nodes;
  int* deg = (int*)malloc(sizeof(int) * ((size_t)N + 1));
  int* bp = (int*)calloc((size_t)N + 1, sizeof(int));
  float* scores = (float*)malloc(sizeof(float) * ((size_t)N + 1));
  int* queue = (int*)malloc(sizeof(int) * ((size_t)N + 1));
  for (int n = 0; n < N; n++) {
    deg[n] = g->nin[n];
    scores[n] = -INFINITY;
  }
  int qh = 0, qn = 0;
  for (int k = 0; k < g->num_start; k++) {
    int n = g->start[k];
    scores[n] = 0.0f;
    bp[n] = -1;
    if (g->nin[n] == 0) queue[qn++] = n;
  }
  while (qh < qn) {
    int n = queue[qh++];
    float s = scores[n];
    for (int k = 0; k < g->nout[n]; k++) {
      int a = g->out[n][k];
      int dn = g->dst[a];
      float ns = s + g->w[a];
      if (ns > scores[dn]) {
        scores[dn] = ns;
        bp[dn] = a;
      }
      if (--deg[dn] == 0) queue[qn++] = dn;
    }
  }
  int rc = OG_OK;
  float best_s = -INFINITY;
  int best = -1;
  for (int k = 0; k < g->num_accept; k++) {
    int n = g->accept[k];
    if (deg[n] > 0) {
      rc = OG_ERR_INVALID;
      break;
    }
    if (scores[n] > best_s) {
      best_s = scores[n];
      best = n;
    }
  }
  if (rc == OG_OK) {
    int len = 0;
    int had = best != -1;
    while (best != -1 && bp[best] != -1) {
      int a = bp[best];
      best = g->src[a];
      path_arcs[len++] = a;
    }
    for (int i = 0; i < len / 2; i++) {
      int t = path_arcs[i];
      path_arcs[i] = path_arcs[len - 1 - i];
      path_arcs[len - 1 - i] = t;
    }
    *path_len = had ? len : -1;
  }
  free(deg);
  free(bp);
  free(scores);
  free(queue);
  return rc;
}

/* ------------------------------------------------------------------ */
/* criteria                                                            */
/* ------------------------------------------------------------------ */

/* benchmarks/ctc.cpp:40-58 (arc_sort=1), test/criterion_test.cpp:37-54 (=0) */
og_graph* og_ctc_graph(const int* target, int U, int blank, int arc_sort) {
  int L = 2 * U + 1;
  og_graph* g = og_new();
  for (int l = 0; l < L; l++) {
    int idx = (l - 1) / 2;
    og_add_node(g, l == 0, l == L - 1 || l == L - 2);
    int label = (l % 2) ? target[idx] : blank;
    og_add_arc(g, l, l, label, label, 0.0f);
    if (l > 0) og_add_arc(g, l - 1, l, label, label, 0.0f);
    if ((l % 2) && l > 1 && label != target[idx - 1])
      og_add_arc(g, l - 2, l, label, label, 0.0f);
  }
  if (arc_sort) og_arc_sort(g, 0);
  return g;
}

/* test/criterion_test.cpp:244-254 */
og_graph* og_transitions_graph(int C, const float* w) {
  og_graph* g = og_new();
  og_add_node(g, 1, 0);
  for (int i = 1; i <= C; i++) {
    og_add_node(g, 0, 1);
    og_add_arc(g, 0, i, i - 1, i - 1, w ? w[i - 1] : 0.0f);
  }
  for (int i = 0; i < C; i++)
    for (int j = 0; j < C; j++)
      og_add_arc(g, j + 1, i + 1, i, i, w ? w[C + i * C + j] : 0.0f);
  return g;
}

/* test/criterion_test.cpp:260-267 */
og_graph* og_fal_graph(const int* target, int U) {
  og_graph* g = og_new();
  og_add_node(g, 1, U == 0);
  for (int l = 1; l <= U; l++) {
    og_add_node(g, 0, l == U);
    og_add_arc(g, l - 1, l, target[l - 1], target[l - 1], 0.0f);
    og_add_arc(g, l, l, target[l - 1], target[l - 1], 0.0f);
  }
  return g;
}

/* forwardScore(g) and, when arc_grads != NULL, its gradient scaled by delta */
static int fwd_and_grad(const og_graph* g, float delta, float* score, float* arc_grads) {
  int N = g->num_nodes;
  float* sc = (float*)malloc(sizeof(float) * ((size_t)N + 1));
  float* mc = (float*)malloc(sizeof(float) * ((size_t)N + 1));
  int rc = og_shortest_distance(g, 0, score, sc, mc, NULL);
  if (rc == OG_OK && arc_grads)
    rc = og_shortest_distance_grad(g, 0, *score, delta, sc, mc, NULL, arc_grads);
  free(sc);
  free(mc);
  return rc;
}

/* benchmarks/ctc.cpp:150-160 */
int og_ctc_loss(
    int T,
    int C,
    const float* emissions,
    const int* target,
    int U,
    int blank,
    int arc_sort,
    float* loss,
    float* grad) {
  og_graph* ctc = og_ctc_graph(target, U, blank, arc_sort);
  og_graph* em = og_linear(T, C, emissions);
  og_graph* lat = og_compose(ctc, em, 1);
  float z = 0, s = 0;
  float* gz = grad ? (float*)malloc(sizeof(float) * (size_t)em->num_arcs) : NULL;
  float* gl = grad ? (float*)malloc(sizeof(float) * ((size_t)lat->num_arcs + 1)) : NULL;
  int rc = fwd_and_grad(em, 1.0f, &z, gz);
  if (rc == OG_OK) rc = fwd_and_grad(lat, -1.0f, &s, gl);
  if (rc == OG_OK) {
    *loss = z - s; /* subtract, functions.cpp:48-64 */
    if (grad) {
      memcpy(grad, gz, sizeof(float) * (size_t)em->num_arcs);
      og_compose_grad(lat, gl, NULL, grad);
    }
  }
  free(gz);
  free(gl);
  og_free(ctc);
  og_free(em);
  og_free(lat);
  return rc;
}

/* test/criterion_test.cpp:271-278 */
int og_asg_loss(
    int T,
    int C,
    const float* emissions,
    const float* trans_w,
    const int* target,
    int U,
    float* loss,
    float* grad,
    float* trans_grad) {
  og_graph* tr = og_transitions_graph(C, trans_w);
  og_graph* fal = og_fal_graph(target, U);
  og_graph* em = og_linear(T, C, emissions);
  og_graph* den = og_compose(em, tr, 0);
  og_graph* ft = og_compose(fal, tr, 0);
  og_graph* num = og_compose(ft, em, 0);
  int want = grad || trans_grad;
  float d = 0, n = 0;
  float* gd = want ? (float*)malloc(sizeof(float) * ((size_t)den->num_arcs + 1)) : NULL;
  float* gn = want ? (float*)malloc(sizeof(float) * ((size_t)num->num_arcs + 1)) : NULL;
  int rc = fwd_and_grad(den, 1.0f, &d, gd);
  if (rc == OG_OK) rc = fwd_and_grad(num, -1.0f, &n, gn);
  if (rc == OG_OK) {
    *loss = d - n;
    if (want) {
      float* ge = (float*)calloc((size_t)em->num_arcs, sizeof(float));
      float* gft = (float*)calloc((size_t)ft->num_arcs + 1, sizeof(float));
      float* gt = (float*)calloc((size_t)tr->num_arcs, sizeof(float));
      og_compose_grad(den, gd, ge, gt);
      og_compose_grad(num, gn, gft, ge);
      og_compose_grad(ft, gft, NULL, gt);
      if (grad) memcpy(grad, ge, sizeof(float) * (size_t)em->num_arcs);
      if (trans_grad)
        for (int a = 0; a < tr->num_arcs; a++) trans_grad[a] += gt[a];
      free(ge);
      free(gft);
      free(gt);
    }
  }
  free(gd);
  free(gn);
  og_free(tr);
  og_free(fal);
  og_free(em);
  og_free(den);
  og_free(ft);
  og_free(num);
  return rc;
}

static int path_labels(const og_graph* lat, int T, int* path, float* score) {
  int* arcs = (int*)malloc(sizeof(int) * ((size_t)lat->num_nodes + 1));
  int len = 0;
  int rc = og_shortest_path(lat, arcs, &len);
  if (rc == OG_OK) {
    for (int t = 0; t < T; t++) path[t] = (t < len) ? lat->ilabel[arcs[t]] : -1;
    if (score) rc = og_shortest_distance(lat, 1, score, NULL, NULL, NULL);
  }
  free(arcs);
  return rc;
}

/* test/criterion_test.cpp:328-344 */
int og_viterbi_dense(
    int T,
    int C,
    const float* emissions,
    const float* trans_w,
    int* path,
    float* score) {
  og_graph* tr = og_transitions_graph(C, trans_w);
  og_graph* em = og_linear(T, C, emissions);
  og_graph* lat = og_compose(em, tr, 0);
  int rc = path_labels(lat, T, path, score);
  og_free(tr);
  og_free(em);
  og_free(lat);
  return rc;
}

int og_viterbi_ctc(
    int T,
    int C,
    const float* emissions,
    const int* target,
    int U,
    int blank,
    int arc_sort,
    int* path,
    float* score) {
  og_graph* ctc = og_ctc_graph(target, U, blank, arc_sort);
  og_graph* em = og_linear(T, C, emissions);
  og_graph* lat = og_compose(ctc, em, 1);
  int rc = path_labels(lat, T, path, score);
  og_free(ctc);
  og_free(em);
  og_free(lat);
  return rc;
}

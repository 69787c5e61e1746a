/*
 * tests/emu/gcompose_emu.cpp -- TEST INFRASTRUCTURE: the SOURCE of gtn_b200/csrc/k_gcompose.cu (general
 * device composition: co-reachability search, ordered forward construction, arc fill) on the CPU through
 * tests/emu/simt_emu.h.  The scratch layout and the two launches of gtnb_compose_graphs are restated here.
 */
#define GTNB_HOST_EMU 1
#include "../../gtn_b200/csrc/k_gcompose.cu"

#include <cstring>
#include <vector>

extern "C" {

struct EmuGraph {
  int32_t N, A;
  const uint8_t* flags;
  const int32_t *src, *dst, *il, *ol;
  const float* w;
  const int32_t *in_ptr, *in_arcs, *out_ptr, *out_arcs, *start;
  int32_t n_start;
  const int32_t* accept;
  int32_t n_accept;
};

/* composes one pair; outputs up to cap_nodes / cap_arcs entries, the true sizes in n_nodes / n_arcs */
int emu_gcompose(
    const EmuGraph* a, const EmuGraph* b, int kind, int reach_ctas, int cap_nodes, int cap_arcs, int32_t* n_nodes,
    int32_t* n_arcs,
    uint8_t* flags, int32_t* src, int32_t* dst, int32_t* il, int32_t* ol, float* w, int32_t* gi1, int32_t* gi2) {
  using namespace gtnb;
  auto view = [](const EmuGraph* g) {
    gc::GView v;
    v.N = g->N;
    v.A = g->A;
    v.flags = g->flags;
    v.src = g->src;
    v.dst = g->dst;
    v.il = g->il;
    v.ol = g->ol;
    v.w = g->w;
    v.in_ptr = g->in_ptr;
    v.in_arcs = g->in_arcs;
    v.out_ptr = g->out_ptr;
    v.out_arcs = g->out_arcs;
    v.start = g->start;
    v.n_start = g->n_start;
    v.accept = g->accept;
    v.n_accept = g->n_accept;
    return v;
  };
  const long long S = (long long)a->N * b->N;
  std::vector<uint32_t> reach((size_t)(S + 31) / 32 + 4, 0u);
  std::vector<int32_t> ids((size_t)S + 4, gc::kUndiscovered), queue((size_t)S + 4, 0), arc_off((size_t)S + 4, 0);
  int32_t counts[4] = {0, 0, 0, 0};
  gc::PairDev P;
  std::memset(&P, 0, sizeof(P));
  P.g1 = view(a);
  P.g2 = view(b);
  P.kind = kind;
  P.reach = reach.data();
  P.ids = ids.data();
  P.queue = queue.data();
  P.arc_off = arc_off.data();
  P.counts = counts;
  int32_t sync[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  P.sync = sync;
  if (reach_ctas > 1) {
    // gtnb_compose_graphs' cooperative launch: the G CTAs of the pair run concurrently
    P.reach_done = 1;
    emu::launch_clusters(reach_ctas, reach_ctas, gc::kThreads, 0, [&] { gc::gcompose_reach_kernel(&P, reach_ctas); });
    if (emu::g_launch_failed.exchange(false)) return emu::kEmuNoThreads;
  }
  emu::launch(1, gc::kThreads, 0, [&] { gc::gcompose_search_kernel(&P); });
  if (emu::g_launch_failed.exchange(false)) return emu::kEmuNoThreads;
  *n_nodes = counts[0];
  *n_arcs = counts[1];
  if (counts[0] > cap_nodes || counts[1] > cap_arcs) return 2;
  P.out_flags = flags;
  P.out_src = src;
  P.out_dst = dst;
  P.out_il = il;
  P.out_ol = ol;
  P.out_w = w;
  P.out_gi1 = gi1;
  P.out_gi2 = gi2;
  const int bpp = (counts[0] + gc::kThreads - 1) / gc::kThreads;
  if (bpp > 0) emu::launch(bpp, gc::kThreads, 0, [&] { gc::gcompose_fill_kernel(&P, bpp); });
  if (emu::g_launch_failed.exchange(false)) return emu::kEmuNoThreads;
  return 0;
}

} // extern "C"

/*
 * gtnb_gcompose.h -- what k_gcompose.cu's kernels read: device views of the two operands of a pair and the
 * pair's scratch / output arrays.  Plain data (also included by the host emulation of tests/emu).
 */
#pragma once

#include <cstdint>

namespace gtnb {
namespace gc {

constexpr int32_t kUndiscovered = (int32_t)0x80808080; // ids[]: what cudaMemset(.., 0x80, ..) leaves

/* one operand: gtn::Graph's arrays (graph.h:56-465), arc ids and node ids as the Graph numbers them */
struct GView {
  int32_t N, A;
  const uint8_t* flags; // bit0 start, bit1 accept
  const int32_t *src, *dst, *il, *ol;
  const float* w;
  const int32_t *in_ptr, *in_arcs; // g.in(n), in the Graph's current order
  const int32_t *out_ptr, *out_arcs; // g.out(n)
  const int32_t* start; // g.start()
  int32_t n_start;
  const int32_t* accept; // g.accept()
  int32_t n_accept;
};

struct PairDev {
  GView g1, g2;
  int32_t kind; // the reference's matcher: 0 unsorted, 1 g1's lists sorted on the matched label, 2 g2's, 3 both
  uint32_t* reach; // [ceil(N1 * N2 / 32)] zeroed: co-reachable product states
  int32_t* ids; // [N1 * N2] kUndiscovered: node id of a product state
  int32_t* queue; // [N1 * N2]: search queue, then product state of every output node
  int32_t* arc_off; // [nodes + 1] first arc id of every output node
  int32_t* sync; // [8] zeroed: gcompose_reach_kernel's queue counters and grid barrier (several CTAs per pair)
  int32_t reach_done; // the co-reachable set and nothing else is in place: gcompose_search_kernel skips its step 1
  int32_t* counts; // [4]: nodes, arcs, microseconds of the co-reachability search and of the forward construction
  // written by gcompose_fill_kernel (allocated once the counts are known)
  uint8_t* out_flags;
  int32_t *out_src, *out_dst, *out_il, *out_ol, *out_gi1, *out_gi2;
  float* out_w;
};

} // namespace gc
} // namespace gtnb

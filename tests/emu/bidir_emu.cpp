/*
 * tests/emu/bidir_emu.cpp -- TEST INFRASTRUCTURE: runs the SOURCE of gtn_b200/csrc/k_bidir.cu (the
 * criterion's bidirectional meet-in-the-middle kernel: two-CTA clusters, TMA-staged emission / score
 * blocks, helper warps) on the CPU through tests/emu/simt_emu.h, fed by k_ctc.cu's ctc_build_kernel
 * exactly as gtnb_ctc_loss feeds it.  Only the launch geometry and the device tables are restated here.
 */
#define GTNB_HOST_EMU 1
#include "../../gtn_b200/csrc/k_ctc.cu"
#include "../../gtn_b200/csrc/k_bidir.cu"

#include <vector>

using gtnb::GraphMeta;

extern "C" {

/*
 * targets concatenated, target_lens[B]; emissions [B][T][C]; input_lens[B] or NULL.
 * out_scores[B] = forwardScore(intersect(ctc_b, e_b)); znorm[B] = forwardScore(e_b);
 * grad [B][T][C] (NULL: loss only) = d(znorm - out) / d e; status[B].
 */
int emu_bidir_ctc(
    int B, int T, int C, const float* emissions, const int32_t* input_lens, const int32_t* targets,
    const int32_t* target_lens, int blank, int zero_w, float* out_scores, float* znorm, float* grad, int32_t* status) {
  using namespace gtnb;
  std::vector<GraphMeta> meta(B);
  std::vector<int32_t> tgt_off(B);
  long long nodes = 0, arcs = 0, scores_len = 0, acc = 0, toff = 0;
  int maxN = 0;
  const bool fx = C == 64 || C == 128; // launch_bidir_ctc's rule: fixed row pitch for the FX kernels
  for (int b = 0; b < B; b++) {
    const int L = 2 * target_lens[b] + 1;
    GraphMeta& m = meta[b];
    std::memset(&m, 0, sizeof(m));
    m.sg_node_base = nodes;
    m.sg_arc_base = arcs;
    m.acc_base = acc;
    m.emis_off = (long long)b * T * C;
    m.node_base = scores_len;
    m.T = input_lens ? input_lens[b] : T;
    m.sg_N = L;
    tgt_off[b] = (int32_t)toff;
    toff += target_lens[b];
    nodes += L + 1;
    arcs += 3 * L;
    acc += 2;
    scores_len += (long long)(T + 1) * (fx ? bidir::kFixedPitch : ((L + 3) & ~3));
    maxN = std::max(maxN, L);
  }
  std::vector<uint8_t> flags(nodes, 0);
  std::vector<int32_t> in_ptr(nodes, 0), in_src(arcs, 0), in_label(arcs, 0), in_arc(arcs, 0), il(arcs, 0),
      ol(arcs, 0), acc_nodes(acc, 0);
  std::vector<float> in_w(arcs, 0.0f);
  emu::launch(B, 256, 0, [&] {
    ctc_build_kernel(meta.data(), targets, tgt_off.data(), target_lens, blank, C, flags.data(), in_ptr.data(),
                     in_src.data(), in_label.data(), in_arc.data(), in_w.data(), il.data(), ol.data(),
                     acc_nodes.data());
  });
  for (int b = 0; b < B; b++) meta[b].sg_A = in_ptr[meta[b].sg_node_base + meta[b].sg_N];
  // the bulk copies need 16-byte aligned sources: the vectors' data() is (malloc), T * C * 4 % 16 == 0 is
  // the caller's business (C % 4 == 0)
  std::vector<float> saved((size_t)scores_len + 16, 0.0f);
  std::vector<float> zparts(2 * (size_t)B * bidir::kHelpers, 0.0f);
  const int nblk_cap = std::max(1, (T + bidir::kBlk - 1) / bidir::kBlk);
  std::vector<float> boff((size_t)B * 2 * nblk_cap * 4 + 16, 0.0f);
  for (int b = 0; b < B; b++) status[b] = 0;
  bidir::Params P;
  P.meta = meta.data();
  P.sg_flags = flags.data();
  P.sg_in_ptr = in_ptr.data();
  P.sg_in_src = in_src.data();
  P.sg_in_label = in_label.data();
  P.sg_in_w = in_w.data();
  P.emissions = emissions;
  P.saved = saved.data();
  P.boff = boff.data();
  P.nblk_cap = nblk_cap;
  P.out_scores = out_scores;
  P.zparts = zparts.data();
  P.status = status;
  P.grad = grad;
  P.grad_stride = (long long)T * C;
  P.C = C;
  P.zero_w = zero_w;
  // launch_bidir_ctc's rule: the PAIR kernels (two nodes per thread) for gtnb_ctc_loss's own graphs; zero_w = 2
  // asks for the general FX kernel on the same graphs
  const bool ctc_ok = fx && (zero_w == 1 || zero_w == 3) && maxN <= 2 * 32 * bidir::kMaxPairWarps - 1;
  const bool pair = ctc_ok && zero_w == 1, quad = ctc_ok && zero_w == 3; // 1: gtnb_ctc_loss's default, 3: "bidir_mode" 2
  P.ctc_blank = blank;
  P.targets = targets;
  for (int b = 0; b < B; b++) meta[b].tgt_off = tgt_off[b];
  P.nwn = quad ? 1 : pair ? std::max(1, ((maxN + 1) / 2 + 31) / 32) : std::max(1, (maxN + 31) / 32);
  P.lay = bidir::make_layout(C, fx ? bidir::kFixedPitch : ((maxN + 3) & ~3));
  const unsigned block = 32 * (P.nwn + 1 + bidir::kHelpers + (quad ? bidir::kQuadSpare : 0));
  emu::launch_clusters(2 * B, 2, block, P.lay.total, [&] {
    // ctc_build_kernel writes weight 0 on every arc: the zero-weight variant, as gtnb_ctc_loss launches it;
    // zero_w = 0 exercises the general one on the same graphs
    if (quad)
      C == 64 ? bidir::bidir_ctc_kernel<1, true, true, 2>(P) : bidir::bidir_ctc_kernel<2, true, true, 2>(P);
    else if (pair)
      C == 64 ? bidir::bidir_ctc_kernel<1, true, true, 1>(P) : bidir::bidir_ctc_kernel<2, true, true, 1>(P);
    else if (fx && C == 64)
      zero_w ? bidir::bidir_ctc_kernel<1, true, true, 0>(P) : bidir::bidir_ctc_kernel<1, false, true, 0>(P);
    else if (fx)
      zero_w ? bidir::bidir_ctc_kernel<2, true, true, 0>(P) : bidir::bidir_ctc_kernel<2, false, true, 0>(P);
    else if (C <= 64)
      zero_w ? bidir::bidir_ctc_kernel<1, true, false, 0>(P) : bidir::bidir_ctc_kernel<1, false, false, 0>(P);
    else
      zero_w ? bidir::bidir_ctc_kernel<2, true, false, 0>(P) : bidir::bidir_ctc_kernel<2, false, false, 0>(P);
  });
  if (emu::g_launch_failed.exchange(false)) return emu::kEmuNoThreads;
  for (int b = 0; b < B; b++) {
    znorm[b] = 0.0f;
    for (int k = 0; k < 2 * bidir::kHelpers; k++) znorm[b] += zparts[2 * bidir::kHelpers * b + k];
  }
  return 0;
}

} // extern "C"

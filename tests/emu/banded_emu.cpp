/*
 * tests/emu/banded_emu.cpp -- TEST INFRASTRUCTURE: runs the SOURCE of gtn_b200/csrc/k_banded.cu
 * (the temporally blocked CTC sweeps) on the CPU through tests/emu/simt_emu.h, one std::thread per
 * CUDA thread, so that tests/test_banded_emulation.py can compare it with the oracle without a GPU.
 * Only the launch geometry and the device tables (what k_ctc.cu's ctc_build_kernel writes) are
 * restated here; the kernels are the product's.
 */
#define GTNB_HOST_EMU 1
#include "../../gtn_b200/csrc/k_banded.cu"

#include <vector>

using gtnb::GraphMeta;

extern "C" {

/*
 * CTC forward (+ backward with deltas = -1, as gtnb_ctc_loss seeds it) of B utterances.
 *   emissions [B][T][C], input_lens[B] (<= T), targets concatenated, target_lens[B], blank
 *   out_scores[B]   forwardScore(intersect(ctc, emissions))
 *   grad [B][T][C]  must be zero on entry: receives -d score / d emissions
 *   status[B]       bit 0 non-finite weight, bit 1 (value 2) graph not band shaped
 * K: frames per barrier (1, 2, 4 or 8).  Returns 0, or 1 when a graph needs more than 16 warps.
 */
int emu_banded_ctc(
    int B, int T, int C, const float* emissions, const int32_t* input_lens, const int32_t* targets,
    const int32_t* target_lens, int blank, int K, float* out_scores, float* grad, int32_t* status) {
  if (K != 1 && K != 2 && K != 4 && K != 8) return 2;
  std::vector<GraphMeta> meta(B);
  std::vector<uint8_t> flags;
  std::vector<int32_t> in_ptr, in_src, in_label;
  std::vector<float> in_w;
  long long node_base = 0, tgt_off = 0;
  int maxN = 0;
  for (int b = 0; b < B; b++) {
    const int U = target_lens[b], L = 2 * U + 1;
    const int32_t* tg = targets + tgt_off;
    GraphMeta& m = meta[b];
    std::memset(&m, 0, sizeof(m));
    m.sg_node_base = (long long)flags.size();
    m.sg_arc_base = (long long)in_src.size();
    m.emis_off = (long long)b * T * C;
    m.node_base = node_base;
    m.T = input_lens ? input_lens[b] : T;
    m.sg_N = L;
    // the tables of k_ctc.cu:ctc_build_kernel: in-arcs of node l in the order skip, step, self
    for (int l = 0; l < L; l++) {
      const int idx = (l - 1) / 2;
      const int label = (l & 1) ? tg[idx] : blank;
      const bool skip = (l & 1) && l > 1 && label != tg[idx - 1];
      flags.push_back((uint8_t)((l == 0 ? 1 : 0) | ((l == L - 1 || l == L - 2) ? 2 : 0)));
      in_ptr.push_back((int32_t)(in_src.size() - m.sg_arc_base));
      auto arc = [&](int src) {
        in_src.push_back(src);
        in_label.push_back((label >= 0 && label < C) ? label : -1);
        in_w.push_back(0.0f);
      };
      if (skip) arc(l - 2);
      if (l > 0) arc(l - 1);
      arc(l);
    }
    in_ptr.push_back((int32_t)(in_src.size() - m.sg_arc_base));
    m.sg_A = (int)(in_src.size() - m.sg_arc_base);
    const int pitch = (L + 3) & ~3;
    node_base += (long long)(T + 1) * pitch;
    tgt_off += U;
    maxN = std::max(maxN, L);
  }
  // sg_in_ptr has N+1 entries per graph: the slabs of consecutive graphs must not share the +1
  // (sg_node_base indexes flags AND in_ptr): rebuild in_ptr with one extra slot per graph
  {
    std::vector<int32_t> ip2;
    std::vector<uint8_t> fl2;
    size_t at_ptr = 0, at_fl = 0;
    for (int b = 0; b < B; b++) {
      const int L = meta[b].sg_N;
      meta[b].sg_node_base = (long long)ip2.size();
      for (int l = 0; l <= L; l++) ip2.push_back(in_ptr[at_ptr + l]);
      for (int l = 0; l < L; l++) fl2.push_back(flags[at_fl + l]);
      fl2.push_back(0);
      at_ptr += L + 1;
      at_fl += L;
    }
    in_ptr.swap(ip2);
    flags.swap(fl2);
  }
  std::vector<float> scores((size_t)node_base + 16, 0.0f);
  for (int b = 0; b < B; b++) status[b] = 0;

  const int own_f = 32 - 2 * K, own_b = 30 - 2 * K;
  const int warps_f = std::max(1, (maxN + own_f - 1) / own_f), warps_b = std::max(1, (maxN + own_b - 1) / own_b);
  if (32 * warps_b > gtnb::kBandMaxThreads) return 1;
  const int pitch_f = (warps_f * own_f + 3) & ~3, pitch_b = (warps_b * own_b + 3) & ~3;
  std::vector<float> deltas(B, -1.0f); // subtract's gradFunc (functions.cpp:53-58)
#define RUN_K(KK)                                                                                             \
  {                                                                                                           \
    emu::launch(B, 32 * warps_f, sizeof(float) * (2 * pitch_f + 32), [&] {                                    \
      gtnb::banded_forward_kernel<KK>(meta.data(), flags.data(), in_ptr.data(), in_src.data(),                \
                                      in_label.data(), in_w.data(), emissions, scores.data(), out_scores,     \
                                      status, C, pitch_f);                                                    \
    });                                                                                                       \
    emu::launch(B, 32 * warps_b, sizeof(float) * (2 * pitch_b + 32), [&] {                                    \
      gtnb::banded_backward_kernel<KK>(meta.data(), flags.data(), in_ptr.data(), in_src.data(),               \
                                       in_label.data(), in_w.data(), emissions, scores.data(), out_scores,    \
                                       deltas.data(), grad, (long long)T * C, C, pitch_b);                    \
    });                                                                                                       \
  }
  switch (K) {
    case 1: RUN_K(1) break;
    case 2: RUN_K(2) break;
    case 8: RUN_K(8) break;
    default: RUN_K(4) break;
  }
#undef RUN_K
  if (emu::g_launch_failed.exchange(false)) return emu::kEmuNoThreads; // the host refused the threads
  return 0;
}

} // extern "C"

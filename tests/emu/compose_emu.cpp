/*
 * tests/emu/compose_emu.cpp -- TEST INFRASTRUCTURE: the SOURCE of gtn_b200/csrc/k_compose.cu
 * (frame-synchronous device composition: alive sets, counts, scans, CSR emission, and compose's
 * gradFunc) on the CPU through tests/emu/simt_emu.h.  The slab layout of gtnb_capi.cu's
 * composed_alloc and the launch geometry of launch_compose are restated here.
 */
#define GTNB_HOST_EMU 1
#include "../../gtn_b200/csrc/k_compose.cu"

#include <vector>

using gtnb::GraphMeta;

namespace {
long long align_up(long long x, long long a) {
  return (x + a - 1) / a * a;
}
} // namespace

extern "C" {

/*
 * compose(g_b, linearGraph(T_b, C) with the given emissions) for B graphs given as concatenated
 * CSR-by-destination tables (n_nodes[B], in_ptr with N_b + 1 entries per graph, offsets local):
 * flags bit 0 start / bit 1 accept, in_src / in_label (-1: cannot match) / in_arc (Graph arc id) / in_w,
 * accept lists acc (n_acc[B]).  Outputs per graph b (capacities (T+1) N_b + 1 nodes, T A_b arcs,
 * concatenated at node_off[b] / arc_off[b] which the caller computes the same way):
 *   out_N[b], out_A[b], row_ptr, arc_src, arc_w, gi_graph, gi_emis, lvl_node_ptr (T_b + 2 per graph at
 *   lvl_off[b]), accept nodes of the lattice (acc_out, n_acc_out[b]).
 * arc_grad (nullable, indexed like the arcs): if given, compose's gradFunc is run into grad_graph
 * (A_b per graph, concatenated) and grad_emis [B][T][C] (both zero on entry).
 */
int emu_compose(
    int B, int T, int C, const float* emissions, const int32_t* lens, const int32_t* n_nodes, const uint8_t* flags,
    const int32_t* in_ptr, const int32_t* in_src, const int32_t* in_label, const int32_t* in_arc, const float* in_w,
    const int32_t* n_acc, const int32_t* acc, int32_t* out_N, int32_t* out_A, uint32_t* row_ptr, int32_t* arc_src,
    float* arc_w, int32_t* gi_graph, int32_t* gi_emis, int32_t* lvl_node_ptr, int32_t* acc_out, int32_t* n_acc_out,
    const float* arc_grad, float* grad_graph, float* grad_emis) {
  using namespace gtnb;
  std::vector<GraphMeta> meta(B);
  std::vector<uint8_t> sg_flags;
  std::vector<int32_t> sg_ptr, sg_src, sg_lab, sg_arc;
  std::vector<float> sg_w;
  long long tn = 0, ta = 0, tl = 0, tc = 0, gg = 0, nb = 0, pb = 0, ab = 0, accb = 0;
  int maxN = 0, maxT = 0;
  std::vector<int32_t> acc_stage;
  for (int b = 0; b < B; b++) {
    const int N = n_nodes[b], A = in_ptr[pb + N], Tb = lens ? lens[b] : T;
    GraphMeta& m = meta[b];
    std::memset(&m, 0, sizeof(m));
    m.sg_node_base = (long long)sg_ptr.size();
    m.sg_arc_base = (long long)sg_src.size();
    int all_valid = 1, uniform = 1;
    for (int n = 0; n <= N; n++) sg_ptr.push_back(in_ptr[pb + n]);
    for (int n = 0; n < N; n++) sg_flags.push_back(flags[nb + n]);
    sg_flags.push_back(0);
    while (sg_ptr.size() % kAlign) sg_ptr.push_back(0), sg_flags.push_back(0);
    for (int a = 0; a < A; a++) {
      sg_src.push_back(in_src[ab + a]);
      sg_lab.push_back(in_label[ab + a]);
      sg_arc.push_back(in_arc[ab + a]);
      sg_w.push_back(in_w[ab + a]);
      all_valid &= in_label[ab + a] >= 0;
    }
    for (int n = 0; n < N; n++)
      for (int a = in_ptr[pb + n]; a < in_ptr[pb + n + 1]; a++) uniform &= in_label[ab + a] == in_label[ab + in_ptr[pb + n]];
    while (sg_src.size() % kAlign || sg_src.empty())
      sg_src.push_back(0), sg_lab.push_back(0), sg_arc.push_back(0), sg_w.push_back(0.0f);
    m.node_base = tn;
    m.arc_base = ta;
    m.lvl_base = tl;
    m.acc_base = tc;
    m.emis_off = (long long)b * T * C;
    m.grad_graph_off = gg;
    gg += A;
    m.L = Tb + 1;
    m.n_accept = n_acc[b];
    m.LB = -1;
    m.T = Tb;
    m.sg_N = N;
    m.sg_A = A;
    m.sg_all_valid = all_valid;
    m.sg_uniform = uniform;
    const long long capN = align_up((long long)(Tb + 1) * N + 1, kAlign), capA = align_up(std::max<long long>((long long)Tb * A, 1), kAlign);
    m.cap_N = (int)capN;
    m.cap_A = (int)capA;
    m.cap_L = (int)align_up(Tb + 2, kAlign);
    tn += capN;
    ta += capA;
    tl += m.cap_L;
    const long long capAcc = align_up(std::max(n_acc[b], 1), kAlign);
    acc_stage.resize(tc + capAcc, 0);
    for (int i = 0; i < n_acc[b]; i++) acc_stage[tc + i] = acc[accb + i];
    tc += capAcc;
    maxN = std::max(maxN, N);
    maxT = std::max(maxT, Tb);
    nb += N;
    pb += N + 1;
    ab += A;
    accb += n_acc[b];
  }
  const int W = (maxN + 31) / 32;
  if (W > kMaxWords) return 1;
  std::vector<uint32_t> alive((size_t)B * (maxT + 1) * W + 16, 0), rp((size_t)tn + 16, 0);
  std::vector<int32_t> lnp((size_t)tl + 16, 0), lap((size_t)tl + 16, 0);
  std::vector<int2> arcs((size_t)ta + 16), gi((size_t)ta + 16);

  emu::launch(B, 256, 0, [&] {
    compose_alive_kernel(meta.data(), sg_flags.data(), sg_ptr.data(), sg_src.data(), sg_lab.data(), alive.data(), W, maxT);
  });
  const int gx = (maxT + 1 + kWarpsPerBlock - 1) / kWarpsPerBlock;
  for (int by = 0; by < B; by++)
    for (int bx = 0; bx < gx; bx++)
      emu::launch(1, 32 * kWarpsPerBlock, 0, [&] {
        blockIdx.x = bx;
        blockIdx.y = by;
        compose_count_kernel(meta.data(), sg_ptr.data(), sg_src.data(), sg_lab.data(), alive.data(), W, maxT,
                             lnp.data(), lap.data());
      });
  emu::launch(B, 1024, 0, [&] {
    compose_scan_kernel(meta.data(), sg_flags.data(), alive.data(), W, maxT, lnp.data(), lap.data(), acc_stage.data());
  });
  for (int by = 0; by < B; by++)
    for (int bx = 0; bx < gx; bx++)
      emu::launch(1, 32 * kWarpsPerBlock, 0, [&] {
        blockIdx.x = bx;
        blockIdx.y = by;
        compose_emit_kernel(meta.data(), sg_flags.data(), sg_ptr.data(), sg_src.data(), sg_lab.data(), sg_arc.data(),
                            sg_w.data(), alive.data(), W, maxT, C, emissions, lnp.data(), lap.data(), rp.data(),
                            arcs.data(), gi.data());
      });

  // hand the lattices back, graph by graph, in the caller's (unaligned) capacities
  long long no = 0, ao = 0, lo = 0, co = 0;
  for (int b = 0; b < B; b++) {
    const GraphMeta& m = meta[b];
    out_N[b] = m.N;
    out_A[b] = m.A;
    for (int n = 0; n <= m.N; n++) row_ptr[no + n] = rp[m.node_base + n];
    for (int a = 0; a < m.A; a++) {
      arc_src[ao + a] = arcs[m.arc_base + a].x;
      arc_w[ao + a] = __int_as_float(arcs[m.arc_base + a].y);
      gi_graph[ao + a] = gi[m.arc_base + a].x;
      gi_emis[ao + a] = gi[m.arc_base + a].y;
    }
    for (int l = 0; l <= m.T + 1; l++) lvl_node_ptr[lo + l] = lnp[m.lvl_base + l];
    n_acc_out[b] = m.n_accept;
    for (int i = 0; i < m.n_accept; i++) acc_out[co + i] = acc_stage[m.acc_base + i];
    no += (long long)(m.T + 1) * m.sg_N + 1;
    ao += std::max<long long>((long long)m.T * m.sg_A, 1);
    lo += m.T + 2;
    co += n_acc[b];
  }
  if (arc_grad) {
    std::vector<float> ag((size_t)ta + 16, 0.0f);
    long long src_off = 0;
    for (int b = 0; b < B; b++) {
      for (int a = 0; a < meta[b].A; a++) ag[meta[b].arc_base + a] = arc_grad[src_off + a];
      src_off += std::max<long long>((long long)meta[b].T * meta[b].sg_A, 1);
    }
    int capN = 1;
    for (int b = 0; b < B; b++) capN = std::max(capN, meta[b].cap_N);
    const int ggx = std::min((capN + 255) / 256, 4096);
    for (int by = 0; by < B; by++)
      for (int bx = 0; bx < ggx; bx++)
        emu::launch(1, 256, 0, [&] {
          blockIdx.x = bx;
          blockIdx.y = by;
          gridDim.x = ggx;
          compose_grad_kernel(meta.data(), rp.data(), ag.data(), gi.data(), grad_graph, grad_emis, (long long)T * C);
        });
  }
  return 0;
}

} // extern "C"

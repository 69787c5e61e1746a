/*
 * tests/emu/linear_emu.cpp -- TEST INFRASTRUCTURE: the SOURCE of gtn_b200/csrc/k_linear.cu (the CTC
 * normaliser: per-frame logsumexp / max of the emissions chain and its gradient) on the CPU through
 * tests/emu/simt_emu.h.  Only the launch geometry is restated here.
 */
#define GTNB_HOST_EMU 1
#include "../../gtn_b200/csrc/k_linear.cu"

#include <vector>

extern "C" {

/* scores[B] = sum_t lse_c / max_c e[b][t][c]; grad (nullable) [B][T][C]; vec != 0: the float4 kernel
 * (needs C = 4 * 2^k <= 128), else the scalar one.  Returns 0, or 1 when vec is not applicable. */
int emu_linear(int B, int T, int C, const float* emissions, const int32_t* lens, int tropical, int vec, float delta,
               int overwrite, float* scores, float* grad) {
  using namespace gtnb;
  std::vector<float> row((size_t)B * std::max(T, 1), 0.0f);
  const long long stride = (long long)T * C;
#define RUN_VEC(TROP, L)                                                                                     \
  {                                                                                                          \
    const int rows_per_block = kRowWarps * (32 / L) * kVecIter;                                              \
    for (int bx = 0; bx < (T + rows_per_block - 1) / rows_per_block; bx++)                                   \
      for (int by = 0; by < B; by++)                                                                         \
        emu::launch(1, 32 * kRowWarps, 0, [&] {                                                              \
          blockIdx.x = bx;                                                                                   \
          blockIdx.y = by;                                                                                   \
          linear_rows_vec_kernel<TROP, L>(lens, T, C, emissions, stride, row.data(), grad, stride, nullptr,  \
                                          delta, overwrite);                                                 \
        });                                                                                                  \
  }
#define RUN_VEC_L(L)        \
  if (tropical)             \
    RUN_VEC(true, L)        \
  else                      \
    RUN_VEC(false, L)
  if (vec) {
    switch (C) {
      case 4: RUN_VEC_L(1) break;
      case 8: RUN_VEC_L(2) break;
      case 16: RUN_VEC_L(4) break;
      case 32: RUN_VEC_L(8) break;
      case 64: RUN_VEC_L(16) break;
      case 128: RUN_VEC_L(32) break;
      default: return 1;
    }
  } else {
    for (int bx = 0; bx < (T + kRowWarps - 1) / kRowWarps; bx++)
      for (int by = 0; by < B; by++)
        emu::launch(1, 32 * kRowWarps, 0, [&] {
          blockIdx.x = bx;
          blockIdx.y = by;
          if (tropical)
            linear_rows_kernel<true>(lens, T, C, emissions, stride, row.data(), grad, stride, nullptr, delta,
                                     overwrite);
          else
            linear_rows_kernel<false>(lens, T, C, emissions, stride, row.data(), grad, stride, nullptr, delta,
                                      overwrite);
        });
  }
  emu::launch(B, 256, 0, [&] { linear_reduce_kernel(lens, T, row.data(), scores); });
  if (emu::g_launch_failed.exchange(false)) return emu::kEmuNoThreads; // the host refused the threads
  return 0;
}

} // extern "C"

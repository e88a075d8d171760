// TEST INFRASTRUCTURE ONLY.  Thin C wrapper around the REFERENCE's own CPU
// deformable im2col (compiled from /root/reference where it lies, see
// oracle/Makefile): per image it calls modulated_deformable_im2col_cpu and then
// performs the bias + W * columns product of dcn_v2_cpu.cpp:76-103 with a
// plain fp32 loop (the reference uses THFloatBlas_gemm, which no longer exists).
#include <vector>

#include "dcn_v2_im2col_cpu.h"   // the reference's own header, found through -I$(DCN)/cpu

extern "C" void dcn_ref_forward(const float* input, const float* weight, const float* bias, const float* offset,
                                const float* mask, float* output, int B, int C, int H, int W, int Co) {
  const int HW = H * W, K = C * 9;
  std::vector<float> col((size_t)K * HW);
  for (int b = 0; b < B; ++b) {
    modulated_deformable_im2col_cpu(input + (size_t)b * C * HW, offset + (size_t)b * 18 * HW,
                                    mask + (size_t)b * 9 * HW, 1, C, H, W, H, W, 3, 3, 1, 1, 1, 1, 1, 1, 1,
                                    col.data());
    for (int o = 0; o < Co; ++o) {
      float* out = output + ((size_t)b * Co + o) * HW;
      for (int p = 0; p < HW; ++p) out[p] = bias[o];
      for (int k = 0; k < K; ++k) {
        const float w = weight[(size_t)o * K + k];
        const float* cp = col.data() + (size_t)k * HW;
        for (int p = 0; p < HW; ++p) out[p] += w * cp[p];
      }
    }
  }
}

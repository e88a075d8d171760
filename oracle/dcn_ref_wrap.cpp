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

// Backward pass (row f-4): the loop of dcn_v2_cpu.cpp:109-224 around the reference's own
// modulated_deformable_col2im_coord_cpu / col2im_cpu / im2col_cpu, with the three THFloatBlas calls
// (gemm 'n','t' -> columns; gemm 't','n' -> grad_weight; gemv -> grad_bias) written as plain fp32 loops.
// 3x3, stride 1, pad 1, dilation 1, one deformable group.  All gradients are overwritten.
extern "C" void dcn_ref_backward(const float* input, const float* weight, const float* offset, const float* mask,
                                 const float* grad_output, float* grad_input, float* grad_offset, float* grad_mask,
                                 float* grad_weight, float* grad_bias, int B, int C, int H, int W, int Co) {
  const int HW = H * W, K = C * 9;
  std::vector<float> col((size_t)K * HW);
  for (size_t i = 0; i < (size_t)B * C * HW; ++i) grad_input[i] = 0.f;
  for (size_t i = 0; i < (size_t)Co * K; ++i) grad_weight[i] = 0.f;
  for (int o = 0; o < Co; ++o) grad_bias[o] = 0.f;
  for (int b = 0; b < B; ++b) {
    const float* go = grad_output + (size_t)b * Co * HW;
    const float* off = offset + (size_t)b * 18 * HW;
    const float* msk = mask + (size_t)b * 9 * HW;
    // columns[k][p] = sum_o weight[o][k] * grad_output[o][p]
    for (int k = 0; k < K; ++k) {
      float* cp = col.data() + (size_t)k * HW;
      for (int p = 0; p < HW; ++p) cp[p] = 0.f;
      for (int o = 0; o < Co; ++o) {
        const float w = weight[(size_t)o * K + k];
        const float* g = go + (size_t)o * HW;
        for (int p = 0; p < HW; ++p) cp[p] += w * g[p];
      }
    }
    modulated_deformable_col2im_coord_cpu(col.data(), input + (size_t)b * C * HW, off, msk, 1, C, H, W, H, W, 3, 3, 1, 1, 1,
                                          1, 1, 1, 1, grad_offset + (size_t)b * 18 * HW, grad_mask + (size_t)b * 9 * HW);
    modulated_deformable_col2im_cpu(col.data(), off, msk, 1, C, H, W, H, W, 3, 3, 1, 1, 1, 1, 1, 1, 1,
                                    grad_input + (size_t)b * C * HW);
    modulated_deformable_im2col_cpu(input + (size_t)b * C * HW, off, msk, 1, C, H, W, H, W, 3, 3, 1, 1, 1, 1, 1, 1, 1,
                                    col.data());
    for (int o = 0; o < Co; ++o) {
      const float* g = go + (size_t)o * HW;
      float bs = 0.f;
      for (int p = 0; p < HW; ++p) bs += g[p];
      grad_bias[o] += bs;
      for (int k = 0; k < K; ++k) {
        const float* cp = col.data() + (size_t)k * HW;
        float a = 0.f;
        for (int p = 0; p < HW; ++p) a += cp[p] * g[p];
        grad_weight[(size_t)o * K + k] += a;
      }
    }
  }
}

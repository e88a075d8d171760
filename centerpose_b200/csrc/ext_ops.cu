// Stand-alone entry points next to the plan:
//   cp_dcn_v2_forward -- the `_ext.dcn_v2_forward` replacement (NCHW in / out), built from
//                        the same implicit-GEMM deformable kernel the plan uses
//                        (reference: DCNv2/src/cuda/dcn_v2_cuda.cu:42-172).
//   cp_preprocess     -- batched uint8 HWC frames -> normalised fp32 NCHW network input
//                        (reference: detectors/base_detector.py:91-148, fix_res branch).
#include <stdlib.h>

#include "common.cuh"

namespace cp {
namespace {

// cv2.warpAffine(INTER_LINEAR, BORDER_CONSTANT 0) with the reference's isotropic
// fix_res affine, followed by ((v / 255) - mean) / std evaluated in double and
// rounded once to float32 like the numpy expression at base_detector.py:132.
// Source coordinates are quantised to 1/32 pixel and the blended value is rounded
// to an integer, mirroring OpenCV's fixed-point remap (INTER_BITS = 5).
__global__ void preprocess_kernel(const uint8_t* __restrict__ frames, float* __restrict__ out, int B, int sh,
                                  int sw, int dh, int dw, float m0, float m1, float m2, float s0, float s1,
                                  float s2) {
  const double cx = (double)(float)(sw / 2.0), cy = (double)(float)(sh / 2.0);
  const double s = (double)(sh > sw ? sh : sw);
  const double a = s / (double)dw;  // src pixels per dst pixel
  size_t total = (size_t)B * dh * dw;
  const float mean[3] = {m0, m1, m2};
  const float stdv[3] = {s0, s1, s2};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int x = i % dw;
    size_t t = i / dw;
    int y = t % dh;
    int n = t / dh;
    double sx = ((double)x - dw * 0.5) * a + cx;
    double sy = ((double)y - dh * 0.5) * a + cy;
    long qx = lrint(sx * 32.0), qy = lrint(sy * 32.0);
    int ix = (int)(qx >> 5), iy = (int)(qy >> 5);
    float fx = (float)(qx & 31) * (1.0f / 32.0f), fy = (float)(qy & 31) * (1.0f / 32.0f);
    const uint8_t* img = frames + (size_t)n * sh * sw * 3;
    for (int c = 0; c < 3; ++c) {
      float v00 = 0, v01 = 0, v10 = 0, v11 = 0;
      if (iy >= 0 && iy < sh) {
        if (ix >= 0 && ix < sw) v00 = img[((size_t)iy * sw + ix) * 3 + c];
        if (ix + 1 >= 0 && ix + 1 < sw) v01 = img[((size_t)iy * sw + ix + 1) * 3 + c];
      }
      if (iy + 1 >= 0 && iy + 1 < sh) {
        if (ix >= 0 && ix < sw) v10 = img[((size_t)(iy + 1) * sw + ix) * 3 + c];
        if (ix + 1 >= 0 && ix + 1 < sw) v11 = img[((size_t)(iy + 1) * sw + ix + 1) * 3 + c];
      }
      float v = (1.f - fy) * ((1.f - fx) * v00 + fx * v01) + fy * ((1.f - fx) * v10 + fx * v11);
      double u8 = (double)(int)(v + 0.5f);
      double r = (u8 / 255.0 - (double)mean[c]) / (double)stdv[c];
      out[(((size_t)n * 3 + c) * dh + y) * dw + x] = (float)r;
    }
  }
}

}  // namespace
}  // namespace cp

using namespace cp;

extern "C" {

static int prec_code(int32_t precision, int* prec) {
  if (precision == CP_PREC_FP32) *prec = -1;
  else if (precision == CP_PREC_BF16) *prec = 0;
  else if (precision == CP_PREC_TF32X3) *prec = 1;
  else if (precision == CP_PREC_TF32) *prec = 2;
  else return fail(CP_ERR_INVALID, "unknown precision");
  return CP_OK;
}

// run one implicit-GEMM launch with the kernel family selected by `prec` (weights already packed as fp32 [K][CoutPad])
static int run_igemm(IgemmParams& p, int prec, int Kreal, cudaStream_t s) {
  if (prec < 0) return conv3_c16_supported(p) ? launch_conv3_c16(p, s) : launch_igemm_fp32(p, s);
  if (prec == 2 || prec == 1) {
    const int x3 = prec == 1;
    const char* force_gather = getenv("CP_FORCE_GATHER");
    if (p.mode == IGEMM_DCN && dcn_tma_supported(p, x3) && !(force_gather && atoi(force_gather))) {
      void* tiles = nullptr;
      CP_CUDA_CHECK(cudaMallocAsync(&tiles, tma_weight_bytes(p.Cin, 9, p.CoutPad, x3), s));
      alignas(64) unsigned char map[128];
      int rc = dcn_tma_encode(p, p.B, map);
      if (!rc) rc = launch_pack_tma_weight(p.wgt, p.CoutPad, p.Cin, 9, p.Cout, p.CoutPad, 1, x3, 16, tiles, s,
                                           dcn_tma_tile_n(p.CoutPad, x3));
      if (!rc) {
        p.wgt_umma = tiles;
        rc = launch_dcn_tma(p, map, x3, 0, s);
      }
      cudaFreeAsync(tiles, s);
      return rc;
    }
    if (!tma_conv_supported(p, x3) || (force_gather && atoi(force_gather))) {
      prec = 1;     // deformable / strided ops: 3-term split gather kernel
    } else {
      void* tiles = nullptr;
      const int taps = p.kh * p.kw;
      CP_CUDA_CHECK(cudaMallocAsync(&tiles, tma_weight_bytes(p.Cin, taps, p.CoutPad, x3), s));
      alignas(64) unsigned char maps[512];
      int rc = tma_conv_encode(p, p.B, x3, maps);
      if (!rc) rc = launch_pack_tma_weight(p.wgt, p.CoutPad, p.Cin, taps, p.Cout, p.CoutPad, 1, x3, tma_cslab(p, x3), tiles, s);
      if (!rc) {
        p.wgt_umma = tiles;
        rc = launch_conv_tma(p, maps, 0, x3, s);
      }
      cudaFreeAsync(tiles, s);
      return rc;
    }
  }
  if (!umma_supported(p, prec)) return fail(CP_ERR_INVALID, "shape not supported by the tcgen05 kernel");
  void* tiles = nullptr;
  CP_CUDA_CHECK(cudaMallocAsync(&tiles, umma_weight_bytes(Kreal, p.CoutPad, prec), s));
  int rc = launch_pack_umma_weight(p.wgt, p.CoutPad, Kreal, p.Cout, p.CoutPad, prec, tiles, s);
  if (!rc) {
    p.wgt_umma = tiles;
    rc = launch_igemm_umma(p, prec, s);
  }
  cudaFreeAsync(tiles, s);
  return rc;
}

int cp_conv2d(const float* x, const float* weight, const float* bias, const float* residual, float* out, int32_t B,
              int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t k, int32_t stride, int32_t pad, int32_t relu,
              int32_t precision, void* stream_) {
  if (!x || !weight || !out) return fail(CP_ERR_INVALID, "cp_conv2d: null argument");
  if (B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || k <= 0 || stride <= 0 || pad < 0)
    return fail(CP_ERR_INVALID, "cp_conv2d: bad shape");
  if (Cin % 16 || Cout % 4) return fail(CP_ERR_INVALID, "cp_conv2d: Cin must be a multiple of 16 and Cout of 4");
  int prec;
  int rc = prec_code(precision, &prec);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream_;
  const int CoPad = round_up(Cout, Cout > 32 ? 64 : (Cout > 16 ? 32 : 16));
  const int K = k * k * Cin;
  float* scratch = nullptr;
  CP_CUDA_CHECK(cudaMallocAsync(&scratch, ((size_t)K * CoPad + CoPad) * sizeof(float), s));
  float* wp = scratch;
  float* bp = wp + (size_t)K * CoPad;
  do {
    if ((rc = launch_pack_conv_weight(weight, nullptr, wp, Cout, Cin, k, k, CoPad, K, CoPad, 0, s))) break;
    if ((rc = launch_pack_bias(bias, nullptr, nullptr, nullptr, nullptr, nullptr, bp, Cout, CoPad, 0.f, s))) break;
    IgemmParams p{};
    p.nsrc = 1;
    p.src[0] = x;
    p.srcC[0] = Cin;
    p.srcStride[0] = Cin;
    p.B = B;
    p.Hin = H;
    p.Win = W;
    p.Cin = Cin;
    p.kh = p.kw = k;
    p.stride = stride;
    p.pad = pad;
    p.Hout = (H + 2 * pad - k) / stride + 1;
    p.Wout = (W + 2 * pad - k) / stride + 1;
    p.Cout = Cout;
    p.CoutPad = CoPad;
    p.Kpad = K;
    p.wgt = wp;
    p.bias = bp;
    p.residual = residual;
    p.resStride = Cout;
    p.relu = relu;
    p.out = out;
    p.outStride = Cout;
    p.mode = IGEMM_NHWC_VEC;
    rc = run_igemm(p, prec, K, s);
  } while (0);
  cudaFreeAsync(scratch, s);
  return rc;
}

int cp_dcn_v2_forward(const float* input, const float* weight, const float* bias, const float* offset,
                      const float* mask, float* output, int32_t B, int32_t C, int32_t H, int32_t W, int32_t Co,
                      void* stream_) {
  return cp_dcn_v2_forward_ex(input, weight, bias, offset, mask, output, B, C, H, W, Co, CP_PREC_FP32, stream_);
}

int cp_dcn_v2_forward_ex(const float* input, const float* weight, const float* bias, const float* offset,
                         const float* mask, float* output, int32_t B, int32_t C, int32_t H, int32_t W, int32_t Co,
                         int32_t precision, void* stream_) {
  int prec;
  {
    int rcp = prec_code(precision, &prec);
    if (rcp) return rcp;
  }
  if (!input || !weight || !bias || !offset || !mask || !output)
    return fail(CP_ERR_INVALID, "cp_dcn_v2_forward: null argument");
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || Co <= 0) return fail(CP_ERR_INVALID, "cp_dcn_v2_forward: bad shape");
  cudaStream_t s = (cudaStream_t)stream_;
  const int Cp = round_up(C, 16);
  const int CoPad = round_up(Co, Co > 32 ? 64 : (Co > 16 ? 32 : 16));
  const size_t npix = (size_t)B * H * W;
  const size_t n_x = npix * Cp, n_om = npix * 32, n_w = (size_t)9 * Cp * CoPad, n_b = CoPad;
  float* scratch = nullptr;
  CP_CUDA_CHECK(cudaMallocAsync(&scratch, (n_x + n_om + n_w + n_b) * sizeof(float), s));
  float* x = scratch;
  float* om = x + n_x;
  float* wp = om + n_om;
  float* bp = wp + n_w;
  int rc = CP_OK;
  do {
    if (Cp != C && cudaMemsetAsync(x, 0, n_x * sizeof(float), s) != cudaSuccess) {
      rc = fail(CP_ERR_CUDA, "cp_dcn_v2_forward: memset");
      break;
    }
    if ((rc = launch_nchw_to_nhwc(input, x, B, C, H, W, Cp, 0, s))) break;
    if ((rc = launch_nchw_to_nhwc(offset, om, B, 18, H, W, 32, 0, s))) break;
    if ((rc = launch_nchw_to_nhwc(mask, om, B, 9, H, W, 32, 18, s))) break;
    if ((rc = launch_pack_conv_weight(weight, nullptr, wp, Co, C, 3, 3, CoPad, 9 * Cp, CoPad, 0, s, Cp))) break;
    if ((rc = launch_pack_bias(bias, nullptr, nullptr, nullptr, nullptr, nullptr, bp, Co, CoPad, 0.f, s))) break;
    IgemmParams p{};
    p.nsrc = 1;
    p.src[0] = x;
    p.srcC[0] = Cp;
    p.srcStride[0] = Cp;
    p.B = B;
    p.Hin = p.Hout = H;
    p.Win = p.Wout = W;
    p.Cin = Cp;
    p.Cout = Co;
    p.CoutPad = CoPad;
    p.kh = p.kw = 3;
    p.stride = 1;
    p.pad = 1;
    p.Kpad = 9 * Cp;
    p.wgt = wp;
    p.bias = bp;
    p.out = output;
    p.out_nchw = 1;
    p.offmask = om;
    p.omStride = 32;
    p.mask_is_logit = 0;
    p.mode = IGEMM_DCN;
    rc = run_igemm(p, prec, 9 * Cp, s);
  } while (0);
  cudaFreeAsync(scratch, s);
  return rc;
}

int cp_preprocess(const uint8_t* frames, float* out, int32_t B, int32_t src_h, int32_t src_w, int32_t dst_h,
                  int32_t dst_w, const float mean[3], const float stdv[3], void* stream_) {
  if (!frames || !out || !mean || !stdv) return fail(CP_ERR_INVALID, "cp_preprocess: null argument");
  if (B <= 0 || src_h <= 0 || src_w <= 0 || dst_h <= 0 || dst_w <= 0)
    return fail(CP_ERR_INVALID, "cp_preprocess: bad shape");
  size_t total = (size_t)B * dst_h * dst_w;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  preprocess_kernel<<<blocks, 256, 0, (cudaStream_t)stream_>>>(frames, out, B, src_h, src_w, dst_h, dst_w, mean[0],
                                                              mean[1], mean[2], stdv[0], stdv[1], stdv[2]);
  CP_LAUNCH_CHECK("preprocess_kernel");
  return CP_OK;
}

}  // extern "C"

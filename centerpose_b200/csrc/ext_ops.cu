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

// cv2.warpAffine(src, M, dsize, flags=INTER_LINEAR) (BORDER_CONSTANT 0) for 8-bit 3-channel frames, restated bit for bit
// (OpenCV imgwarp.cpp WarpAffineInvoker + remapBilinear, third party: opencv-python >= 4.5.3.56, 4.13.0 in this image,
// checked by tests/test_preprocess_host.py against cv2 itself):
//   * M is inverted on the host exactly like cv::warpAffine does (cp_preprocess_affine below);
//   * AB_BITS = 10: X0 = cvRound((M[1] y + M[2]) * 1024) + 16, adelta[x] = cvRound(M[0] x * 1024) (cvRound = round
//     half to even), X = (X0 + adelta[x]) >> 5 -> integer source pixel X >> 5 and a 1/32-pixel fraction X & 31;
//   * the four bilinear weights are the integers (32 - fy)(32 - fx) * 32, ... (sum 32768, INTER_REMAP_COEF_BITS = 15);
//   * pixel = (sum of weight * source + 16384) >> 15, neighbours outside the image contribute the border value 0;
// followed by ((v / 255.) - mean) / std evaluated in double and rounded once to float32 like the numpy expression at
// base_detector.py:132.  `Minv` = the inverted 2 x 3 matrix (dst -> src).
struct WarpM {
  double m[6];
};

__global__ void preprocess_kernel(const uint8_t* __restrict__ frames, float* __restrict__ out, int B, int sh,
                                  int sw, int dh, int dw, const WarpM W, float m0, float m1, float m2, float s0, float s1,
                                  float s2) {
  size_t total = (size_t)B * dh * dw;
  const float mean[3] = {m0, m1, m2};
  const float stdv[3] = {s0, s1, s2};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % dw);
    const size_t t = i / dw;
    const int y = (int)(t % dh);
    const int n = (int)(t / dh);
    // unfused double arithmetic (the host code OpenCV runs here has no FMA contraction)
    const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(W.m[1], (double)y), W.m[2]), 1024.0)) + 16;
    const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(W.m[4], (double)y), W.m[5]), 1024.0)) + 16;
    const int ad = __double2int_rn(__dmul_rn(__dmul_rn(W.m[0], (double)x), 1024.0));
    const int bd = __double2int_rn(__dmul_rn(__dmul_rn(W.m[3], (double)x), 1024.0));
    const int X = (X0 + ad) >> 5, Y = (Y0 + bd) >> 5;
    int ix = X >> 5, iy = Y >> 5;
    // saturate_cast<short> of the integer coordinates (only matters for absurd scales; keeps the restatement exact)
    ix = max(-32768, min(32767, ix));
    iy = max(-32768, min(32767, iy));
    const int fx = X & 31, fy = Y & 31;
    const int w00 = (32 - fy) * (32 - fx) * 32, w01 = (32 - fy) * fx * 32, w10 = fy * (32 - fx) * 32, w11 = fy * fx * 32;
    const uint8_t* img = frames + (size_t)n * sh * sw * 3;
    const bool y0 = iy >= 0 && iy < sh, y1 = iy + 1 >= 0 && iy + 1 < sh;
    const bool x0 = ix >= 0 && ix < sw, x1 = ix + 1 >= 0 && ix + 1 < sw;
    for (int c = 0; c < 3; ++c) {
      int v00 = 0, v01 = 0, v10 = 0, v11 = 0;
      if (y0) {
        if (x0) v00 = img[((size_t)iy * sw + ix) * 3 + c];
        if (x1) v01 = img[((size_t)iy * sw + ix + 1) * 3 + c];
      }
      if (y1) {
        if (x0) v10 = img[((size_t)(iy + 1) * sw + ix) * 3 + c];
        if (x1) v11 = img[((size_t)(iy + 1) * sw + ix + 1) * 3 + c];
      }
      int u8 = (v00 * w00 + v01 * w01 + v10 * w10 + v11 * w11 + (1 << 14)) >> 15;
      u8 = max(0, min(255, u8));
      const double r = ((double)u8 / 255.0 - (double)mean[c]) / (double)stdv[c];
      out[(((size_t)n * 3 + c) * dh + y) * dw + x] = (float)r;
    }
  }
}

}  // namespace

int dcn_v2_backward_impl(const float* input, const float* weight, const float* offset, const float* mask,
                         const float* grad_output, float* grad_input, float* grad_offset, float* grad_mask,
                         float* grad_weight, float* grad_bias, int B, int C, int H, int W, int Co, int prec,
                         cudaStream_t s);      // dcn_bwd.cu
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

}  // extern "C"
namespace cp {
int run_igemm_dispatch(IgemmParams& p, int prec, int Kreal, cudaStream_t s) { return run_igemm(p, prec, Kreal, s); }
}  // namespace cp
extern "C" {

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

int cp_dcn_v2_backward(const float* input, const float* weight, const float* offset, const float* mask,
                       const float* grad_output, float* grad_input, float* grad_offset, float* grad_mask,
                       float* grad_weight, float* grad_bias, int32_t B, int32_t C, int32_t H, int32_t W, int32_t Co,
                       int32_t precision, void* stream_) {
  int prec;
  {
    int rcp = prec_code(precision, &prec);
    if (rcp) return rcp;
  }
  if (!input || !weight || !offset || !mask || !grad_output || !grad_input || !grad_offset || !grad_mask || !grad_weight ||
      !grad_bias)
    return fail(CP_ERR_INVALID, "cp_dcn_v2_backward: null argument");
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || Co <= 0) return fail(CP_ERR_INVALID, "cp_dcn_v2_backward: bad shape");
  return dcn_v2_backward_impl(input, weight, offset, mask, grad_output, grad_input, grad_offset, grad_mask, grad_weight,
                              grad_bias, B, C, H, W, Co, prec, (cudaStream_t)stream_);
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

int cp_preprocess_affine(const uint8_t* frames, float* out, int32_t B, int32_t src_h, int32_t src_w, int32_t dst_h,
                         int32_t dst_w, const double trans_input[6], const float mean[3], const float stdv[3], void* stream_) {
  if (!frames || !out || !mean || !stdv || !trans_input) return fail(CP_ERR_INVALID, "cp_preprocess: null argument");
  if (B <= 0 || src_h <= 0 || src_w <= 0 || dst_h <= 0 || dst_w <= 0)
    return fail(CP_ERR_INVALID, "cp_preprocess: bad shape");
  // cv::warpAffine inverts the forward matrix like this (imgwarp.cpp), in double, before the fixed-point walk
  WarpM W;
  double* M = W.m;
  for (int i = 0; i < 6; ++i) M[i] = trans_input[i];
  double D = M[0] * M[4] - M[1] * M[3];
  D = D != 0 ? 1. / D : 0;
  const double A11 = M[4] * D, A22 = M[0] * D;
  M[0] = A11;
  M[1] *= -D;
  M[3] *= -D;
  M[4] = A22;
  const double b1 = -M[0] * M[2] - M[1] * M[5];
  const double b2 = -M[3] * M[2] - M[4] * M[5];
  M[2] = b1;
  M[5] = b2;
  size_t total = (size_t)B * dst_h * dst_w;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  preprocess_kernel<<<blocks, 256, 0, (cudaStream_t)stream_>>>(frames, out, B, src_h, src_w, dst_h, dst_w, W, mean[0],
                                                              mean[1], mean[2], stdv[0], stdv[1], stdv[2]);
  CP_LAUNCH_CHECK("preprocess_kernel");
  return CP_OK;
}

// fix_res affine of base_detector.py:109-121 (c = frame centre, s = max side, rot 0) in closed form: the float32 control
// points of utils/image.py:35-68 give an isotropic scale a = dst_w / s.  (cv2.getAffineTransform solves the same three
// point pairs by LU; the two agree to <= 3e-14, which the fixed-point walk cannot see except on exact rounding ties.
// Callers that hold the reference's own `trans_input` pass it to cp_preprocess_affine.)
int cp_preprocess(const uint8_t* frames, float* out, int32_t B, int32_t src_h, int32_t src_w, int32_t dst_h,
                  int32_t dst_w, const float mean[3], const float stdv[3], void* stream_) {
  if (src_h <= 0 || src_w <= 0 || dst_h <= 0 || dst_w <= 0) return fail(CP_ERR_INVALID, "cp_preprocess: bad shape");
  const float cx = (float)(src_w / 2.0), cy = (float)(src_h / 2.0);
  const float s = (float)(src_h > src_w ? src_h : src_w);
  const float src1y = cy + s * -0.5f;                               // float32 control points
  const float dst0x = (float)(dst_w * 0.5), dst0y = (float)(dst_h * 0.5);
  const float dst1y = dst0y + (float)(dst_w * -0.5);
  const double a = ((double)dst1y - (double)dst0y) / ((double)src1y - (double)cy);
  const double T[6] = {a, 0.0, (double)dst0x - a * (double)cx, 0.0, a, (double)dst0y - a * (double)cy};
  return cp_preprocess_affine(frames, out, B, src_h, src_w, dst_h, dst_w, T, mean, stdv, stream_);
}

}  // extern "C"

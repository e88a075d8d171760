// Shared declarations for libcenterpose_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/centerpose_b200.h"

namespace cp {

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
extern thread_local long long g_launch_counter;      // kernels launched by this thread (every CP_LAUNCH_CHECK)

#define CP_CUDA_CHECK(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      return ::cp::fail(CP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

#define CP_LAUNCH_CHECK(what)                                                            \
  do {                                                                                   \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess)                                                               \
      return ::cp::fail(CP_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(_e));  \
    ++::cp::g_launch_counter;                                                            \
  } while (0)

// ---------------------------------------------------------------------------
// Implicit-GEMM convolution, fp32 CUDA-core path (parity mode).
//   out[m, n] = epilogue( sum_k A[m, k] * Wp[k, n] )
//   m = (b, oy, ox) output pixel, n = output channel, k = (ky, kx, ci).
// A is never materialised: it is gathered on the fly from up to 4 NHWC sources
// (channel concatenation, Root nodes), from an NCHW tensor (stems), or by
// bilinear deformable sampling driven by an offset/mask tensor (DCNv2).
// ---------------------------------------------------------------------------
enum IgemmMode {
  IGEMM_NCHW_SCALAR = 0,  // generic k -> (tap, c) decode, NCHW input (7x7 stems, Cin = 1/3/8)
  IGEMM_NHWC_VEC = 1,     // Cin of every source % 16 == 0, float4 gathers
  IGEMM_DCN = 2           // 3x3 s1 p1 modulated deformable sampling, single NHWC source
};

struct IgemmParams {
  const float* src[4];
  int srcC[4];       // channels contributed by each source
  int srcStride[4];  // pixel stride (floats) of each NHWC source (>= srcC)
  int nsrc;
  int B, Hin, Win, Cin;
  int Hout, Wout, Cout, CoutPad;
  int kh, kw, stride, pad;
  int Kpad;              // rows of the packed weight matrix (multiple of 16)
  const float* wgt;      // [Kpad][CoutPad], BN scale folded in
  const float* bias;     // [CoutPad], conv bias + BN shift folded
  const float* residual; // NHWC, Cout channels, pixel stride resStride; or null
  int resStride;
  int relu;
  int res_after_relu;    // 1: out = relu(acc + bias) + residual  (tracking stems)
  float* out;
  int outStride;         // NHWC pixel stride
  int out_nchw;          // 1: write NCHW [B, Cout, Hout, Wout]
  const float* offmask;  // DCN: NHWC [.., omStride] raw conv_offset_mask output (27 used)
  int omStride;
  int mask_is_logit;     // 1: apply sigmoid to channels 18..26
  int mode;
  const void* wgt_umma;  // tcgen05 path: pre-swizzled weight tiles (igemm_umma.cu), else null
  // conv_tma only: the per-head 1x1 convolutions fused into the epilogue of the merged heads 3x3 conv.  Head h owns the
  // output columns [h * fuse_hidden, (h + 1) * fuse_hidden); its 1x1 weights are [fuse_hidden][16] fp32 (rows = hidden
  // channel, 16 padded outputs), bias [16], output NCHW [B, fuse_cout[h], Hout, Wout].  fuse_n == 0: not fused.
  // conv_tma split-K workspace (plan-owned; null = no split-K): partial sums
  float* splitk_ws;
  size_t splitk_ws_floats;
  int fuse_n, fuse_hidden;
  const float* fuse_w[16];
  const float* fuse_b[16];
  float* fuse_out[16];
  int fuse_cout[16];
};

int launch_igemm_fp32(const IgemmParams& p, cudaStream_t stream);

// dedicated 7x7 stem kernel (stem_conv.cu); consumes the same packed fp32 weights as the generic kernel
bool conv3_c16_supported(const IgemmParams& p);     // direct 3x3 16 -> 16 NHWC convolution (stem_conv.cu)
int launch_conv3_c16(const IgemmParams& p, cudaStream_t stream);
bool stem_supported(const IgemmParams& p);
int launch_stem_conv(const IgemmParams& p, cudaStream_t s);

// tcgen05 tensor-core path (igemm_umma.cu).  prec: 0 = bf16 (kind::f16), 1 = tf32 x 3 (kind::tf32, fp32-equivalent)
bool umma_supported(const IgemmParams& p, int prec);
size_t umma_weight_bytes(int Kreal, int CoutPad, int prec);
int launch_pack_umma_weight(const float* src_k_by_ld, int ld, int Kreal, int Cout, int CoutPad, int prec, void* dst,
                            cudaStream_t s);
int launch_igemm_umma(const IgemmParams& p, int prec, cudaStream_t stream);

// elementwise / data-movement kernels (elementwise.cu)
int launch_maxpool2(const float* in, float* out, int B, int H, int W, int C, cudaStream_t s);
int launch_upsample_add(const float* in, const float* wgt_kkc, const float* skip, float* out, int B,
                        int Hin, int Win, int C, int f, cudaStream_t s);
// writes the [Kpad x CoutPad] block at column `colOff` of a row-major matrix with leading dimension `ld`
int launch_pack_conv_weight(const float* w_oihw, const float* scale, float* out, int Cout, int Cin,
                            int kh, int kw, int CoutPad, int Kpad, int ld, int colOff, cudaStream_t s,
                            int CinPad = 0);
int launch_pack_bias(const float* conv_bias, const float* bn_w, const float* bn_b, const float* bn_mean,
                     const float* bn_var, float* scale_out, float* bias_out, int C, int CPad, float eps,
                     cudaStream_t s);
int launch_pack_up_weight(const float* w_c1kk, float* out_kkc, int C, int k, cudaStream_t s);
int launch_nchw_to_nhwc(const float* in, float* out, int B, int C, int H, int W, int outStride,
                        int chanOffset, cudaStream_t s);
int launch_nhwc_to_nchw(const float* in, float* out, int B, int C, int H, int W, int inStride, cudaStream_t s);
int launch_group_norm_relu(float* x, const float* gamma, const float* beta, int B, int HW, int C,
                           int stride, int chanOffset, int groups, float eps, float* stats, cudaStream_t s);
int launch_gru_gates(const float* xi, const float* hh, const float* hprev, float* hout, int B_HW, int C,
                     int first_step, cudaStream_t s);

// TMA-fed shifted-window tcgen05 convolution (conv_tma.cu): stride-1 1x1 / 3x3 over NHWC fp32, kind::tf32
// x3 = 1: 3-term split with two-level accumulation (fp32-equivalent); x3 = 0: single tf32 pass
bool tma_conv_supported(const IgemmParams& p, int x3);
size_t tma_weight_bytes(int Cin, int taps, int CoutPad, int x3);
int tma_tile_n(int CoutPad, int x3);             // N tile of conv_tma for this output width
int tma_cslab(const IgemmParams& p, int x3);     // channels per activation slab (32 or 16); needs Cin, kh, Win, CoutPad
int x3_group_blocks();                           // tf32x3: 32-channel K blocks per TMEM accumulation group
int launch_pack_tma_weight(const float* src_k_by_ld, int ld, int Cin, int taps, int Cout, int CoutPad, int round_tf32,
                           int x3, int cslab, void* dst, cudaStream_t s, int bn_override = 0);
int tma_encode_nhwc_box(const float* base, int C, int W, int H, int B, int strideFloats, int boxC, int boxW, int boxH,
                        int swizzle64, void* map_out /* 128 bytes, 64-byte aligned */);

// TMA-staged deformable convolution (dcn_tma.cu): DCNv2 3x3 stride 1 pad 1 over one NHWC fp32 source, kind::tf32.
// x3 = 1: 3-term split + promoted accumulation (fp32-equivalent);  x3 = 0: single pass.
bool dcn_tma_supported(const IgemmParams& p, int x3);
int dcn_tma_tile_n(int CoutPad, int x3);
int dcn_tma_encode(const IgemmParams& p, int Bmax, void* map_out /* 128 bytes, 64-byte aligned */);
int launch_dcn_tma(const IgemmParams& p, const void* map, int x3, int round_out_tf32, cudaStream_t stream);
int tma_conv_encode(const IgemmParams& p, int Bmax, int x3, void* maps_out /* 4 x 128 bytes */);
int launch_conv_tma(const IgemmParams& p, const void* maps, int round_out_tf32, int x3, cudaStream_t stream);

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// ---------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  The forward is ~90 dependent launches; at batch 1 a launch is ~25 us of which the
// launch latency + the fixed prologue of a tcgen05 CTA (barrier init, TMEM allocation, first weight tiles) is a third.
// Every kernel of the forward schedule therefore (a) signals `launch_dependents` at its very start, so the NEXT kernel's
// CTAs are placed on an SM the moment a CTA of this one retires, and (b) executes `griddep_wait()` before its first
// read of an activation / first global write.  What runs before the wait touches only per-plan constants (weights,
// biases) and the CTA's own shared memory / TMEM.  Kernels launched without the attribute see both instructions as no-ops.
// g_pdl is set by run_forward (plan.cu) around the op loop; stand-alone ops (cp_conv2d ...) pack their weights on the
// stream right before the launch and therefore never use it.  CP_NO_PDL=1 disables it (A/B runs).
extern thread_local int g_pdl;
#ifdef __CUDACC__
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif
constexpr size_t kSplitkWsFloats = (size_t)160 * 256 * 128;     // >= (#SMs) tile-splits of 256 positions x 128 columns

// Launch attributes (cudaFuncSetAttribute) and the SM count are properties of the CURRENT DEVICE, not of the calling
// thread: the caches below are indexed by cudaGetDevice() so that a process which runs plans on cuda:0 and cuda:1
// configures the > 48 KB shared-memory kernels on both.  (Two threads racing on the same slot set the same attribute
// twice, which is harmless.)
constexpr int kMaxDevices = 64;
inline int current_device_slot() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDevices) d = 0;
  return d;
}
template <typename T, int N = 1>
struct PerDevice {
  T v[kMaxDevices][N] = {};
  T& here(int slot = 0) { return v[current_device_slot()][slot]; }
};
inline int device_sm_count(int* out) {
  static PerDevice<int> cache;
  int& n = cache.here();
  if (!n) {
    int dev = 0;
    CP_CUDA_CHECK(cudaGetDevice(&dev));
    CP_CUDA_CHECK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  }
  *out = n;
  return CP_OK;
}

}  // namespace cp

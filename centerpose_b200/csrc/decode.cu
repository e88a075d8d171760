// Heat-map decode, keypoint-to-centre grouping, output-map -> image affine,
// score filter + Gaussian soft-NMS and per-object PnP, entirely on the GPU.
//
// Two kernels per batch:
//   peaks_topk_kernel   one CTA per (image, heat-map channel): sigmoid, 3x3
//                       equality NMS and an exact top-K (radix select + bitonic
//                       sort) in shared memory -- each map is read from HBM once.
//   group_pose_kernel   one CTA per image: gathers at the K centres, K x K
//                       nearest-peak match per joint, the decode.py gates, the
//                       post_process.py affine, soft-NMS and the PnP solve
//                       (pose_core.h), writing the fixed-shape pose records.
//
// Reference semantics reproduced (paths relative to /root/reference/src/lib):
//   detectors/object_pose.py:136-138  sigmoid;  models/decode.py:17-23 _nms;
//   :40-68 _topk / _topk_channel;  :72-375 object_pose_decode(Inference=True);
//   utils/post_process.py:12-68;  utils/image.py:23-74 (rot = 0);
//   detectors/object_pose.py:184-197 merge_outputs, :27-124 soft_nms_nvidia;
//   detectors/base_detector.py:548-566 point assembly;  utils/pnp/*.
// The seven-gate test at decode.py:183-188 follows the pinned torch==1.1.0
// semantics (uint8 adds, `== 7` means all gates hold) -- see DESIGN.md.
#include "common.cuh"
#include "pose_core.h"

namespace cp {
namespace {

constexpr float SENT = -10000.0f;
constexpr int TOPK_THREADS = 1024;
constexpr int KM = CP_MAX_K;

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---------------------------------------------------------------------------
// kernel 1: per-channel sigmoid + NMS + top-K
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(TOPK_THREADS, 1)
peaks_topk_kernel(const float* __restrict__ hm, const float* __restrict__ hm_hp, int C_hm, int J, int H, int W,
                  int K, int apply_sigmoid, float* __restrict__ peak_val, int* __restrict__ peak_idx) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int HW = H * W;
  float* raw = reinterpret_cast<float*>(smem_raw);
  float* nv = raw + HW;
  __shared__ unsigned int hist[256];
  __shared__ unsigned int s_prefix, s_need, s_cnt;
  __shared__ int warp_tot[TOPK_THREADS / 32];
  __shared__ unsigned long long keys[KM];

  const int ch = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int CH = C_hm + J;
  const float* src = (ch < C_hm) ? hm + ((size_t)b * C_hm + ch) * HW : hm_hp + ((size_t)b * J + (ch - C_hm)) * HW;

  for (int i = tid; i < HW; i += TOPK_THREADS) {
    float v = __ldg(src + i);
    raw[i] = apply_sigmoid ? sigmoid_acc(v) : v;
  }
  __syncthreads();
  // 3x3 max-pool (stride 1, -inf padding) equality NMS: keep = (hmax == heat)
  for (int i = tid; i < HW; i += TOPK_THREADS) {
    int y = i / W, x = i - y * W;
    float v = raw[i];
    float m = v;
    for (int dy = -1; dy <= 1; ++dy) {
      int yy = y + dy;
      if (yy < 0 || yy >= H) continue;
      for (int dx = -1; dx <= 1; ++dx) {
        int xx = x + dx;
        if (xx < 0 || xx >= W) continue;
        m = fmaxf(m, raw[yy * W + xx]);
      }
    }
    nv[i] = (m == v) ? v : v * 0.0f;
  }
  if (tid == 0) {
    s_prefix = 0;
    s_need = K;
    s_cnt = 0;
  }
  __syncthreads();

  // radix select of the K-th largest value (values are >= 0 so the uint order is the float order)
  unsigned int mask = 0;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = tid; i < 256; i += TOPK_THREADS) hist[i] = 0;
    __syncthreads();
    const unsigned int prefix = s_prefix;
    for (int i = tid; i < HW; i += TOPK_THREADS) {
      unsigned int bits = __float_as_uint(nv[i]);
      if ((bits & mask) == prefix) atomicAdd(&hist[(bits >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      unsigned int need = s_need, cum = 0;
      int d = 255;
      for (; d > 0; --d) {
        if (cum + hist[d] >= need) break;
        cum += hist[d];
      }
      s_need = need - cum;
      s_prefix = prefix | ((unsigned int)d << shift);
    }
    mask |= 255u << shift;
    __syncthreads();
  }
  const unsigned int T = s_prefix;      // bit pattern of the K-th largest value
  const unsigned int need = s_need;     // how many elements equal to T are selected (lowest indices first)
  const unsigned int n_gt = K - need;

  for (int i = tid; i < KM; i += TOPK_THREADS) keys[i] = 0ull;
  __syncthreads();
  // strictly greater: unordered compaction
  for (int i = tid; i < HW; i += TOPK_THREADS) {
    unsigned int bits = __float_as_uint(nv[i]);
    if (bits > T) {
      unsigned int pos = atomicAdd(&s_cnt, 1u);
      keys[pos] = ((unsigned long long)bits << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned int)i);
    }
  }
  // ties: ordered by index -> contiguous chunk per thread + block scan
  const int per = (HW + TOPK_THREADS - 1) / TOPK_THREADS;
  const int i0 = tid * per, i1 = min(HW, i0 + per);
  int mine = 0;
  for (int i = i0; i < i1; ++i) mine += (__float_as_uint(nv[i]) == T);
  int incl = mine;
  const int lane = tid & 31, wid = tid >> 5;
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_tot[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int v = warp_tot[lane];
    int s = v;
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    warp_tot[lane] = s - v;  // exclusive
  }
  __syncthreads();
  int rank = warp_tot[wid] + incl - mine;
  for (int i = i0; i < i1 && rank < (int)need; ++i) {
    if (__float_as_uint(nv[i]) == T) {
      keys[n_gt + rank] = ((unsigned long long)T << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned int)i);
      ++rank;
    }
  }
  __syncthreads();
  // bitonic sort of KM keys, descending (value desc, index asc)
  for (int k2 = 2; k2 <= KM; k2 <<= 1) {
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      if (tid < KM) {
        int ixj = tid ^ j;
        if (ixj > tid) {
          unsigned long long a = keys[tid], c = keys[ixj];
          bool desc = ((tid & k2) == 0);
          if ((a < c) == desc) {
            keys[tid] = c;
            keys[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  if (tid < K) {
    unsigned long long kk = keys[tid];
    peak_val[((size_t)b * CH + ch) * K + tid] = __uint_as_float((unsigned int)(kk >> 32));
    peak_idx[((size_t)b * CH + ch) * K + tid] = (int)(0xFFFFFFFFu - (unsigned int)(kk & 0xFFFFFFFFull));
  }
}

// ---------------------------------------------------------------------------
// kernel 2: grouping + post-process + soft-NMS + PnP
// ---------------------------------------------------------------------------
struct GroupArgs {
  cp_decode_params prm;
  cp_heads h;
  const double* meta;
  const float* peak_val;
  const int* peak_idx;
  float* dets;    // [B,K,CP_DETS_RECORD] (user buffer or workspace)
  float* poses;   // [B,K,CP_POSE_RECORD]
  int* n_valid;
};

__device__ __forceinline__ float gatherf(const float* base, int b, int C, int c, int HW, int ind) {
  return __ldg(base + ((size_t)b * C + c) * HW + ind);
}

// python slice start:stop on a length-n axis
__device__ __forceinline__ void py_slice(int start, int stop, int n, int* s0, int* s1) {
  if (start < 0) start += n;
  if (stop < 0) stop += n;
  start = max(0, min(start, n));
  stop = max(0, min(stop, n));
  *s0 = start;
  *s1 = max(start, stop);
}

// ------------------------------------------------------------------------------------------------------------------
// Warp-cooperative PnP.  One thread per detection (the first version) left ~4 ms of serial double-precision latency on
// a handful of lanes: a 12x12 Jacobi eigen-solve and the LM loop with their matrices in local memory.  Here a WARP
// owns a detection: the matrices live in shared memory, the Jacobi rotations are applied by 12 + 12 lanes, the LM
// Jacobian is one point per lane, and the small dependent pieces (rotation angles, 6x6 Cholesky, Rodrigues) are
// computed redundantly by every lane from identical inputs (bitwise identical results, so control flow stays uniform).
// Same algorithm and iteration order as pose::dlt_init / pose::refine_lm (pose_core.h), which remain the host-tested
// statement of the math.
constexpr int PNP_SCRATCH = 320;      // doubles per warp: [0,288) Jacobi A|V or LM workspace, [288,320) image points

__device__ void dlt_init_warp(const double* X, const double* uv, int n, double fx, double fy, double cx, double cy,
                              double* R, double* t, double* sm, int lane) {
  double* A = sm;
  double* V = sm + 144;
  for (int e = lane; e < 144; e += 32) {
    const int a = e / 12, b = e - a * 12;
    double acc = 0.0;
    for (int i = 0; i < n; ++i) {
      const double x = (uv[2 * i] - cx) / fx, y = (uv[2 * i + 1] - cy) / fy;
      const double h[4] = {X[3 * i], X[3 * i + 1], X[3 * i + 2], 1.0};
      const double r1a = a < 4 ? h[a] : (a < 8 ? 0.0 : -x * h[a - 8]);
      const double r1b = b < 4 ? h[b] : (b < 8 ? 0.0 : -x * h[b - 8]);
      const double r2a = a < 4 ? 0.0 : (a < 8 ? h[a - 4] : -y * h[a - 8]);
      const double r2b = b < 4 ? 0.0 : (b < 8 ? h[b - 4] : -y * h[b - 8]);
      acc += r1a * r1b + r2a * r2b;
    }
    A[e] = acc;
    V[e] = (a == b) ? 1.0 : 0.0;
  }
  __syncwarp();
  const int k = lane < 12 ? lane : lane - 12;       // lanes 0-11 rotate A, lanes 12-23 rotate V
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < 12; ++i) {
      diag += A[i * 12 + i] * A[i * 12 + i];
      for (int j = i + 1; j < 12; ++j) off += A[i * 12 + j] * A[i * 12 + j];
    }
    if (off <= 1e-60 * diag || off == 0.0) break;
    for (int p = 0; p < 11; ++p)
      for (int q = p + 1; q < 12; ++q) {
        const double apq = A[p * 12 + q];
        if (apq == 0.0) continue;
        const double app = A[p * 12 + p], aqq = A[q * 12 + q];
        const double theta = (aqq - app) / (2.0 * apq);
        const double tt = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(tt * tt + 1.0), s = tt * c;
        __syncwarp();                                 // everybody has read A[p][q], A[p][p], A[q][q]
        if (lane < 12) {
          const double akp = A[k * 12 + p], akq = A[k * 12 + q];
          A[k * 12 + p] = c * akp - s * akq;
          A[k * 12 + q] = s * akp + c * akq;
        } else if (lane < 24) {
          const double vkp = V[k * 12 + p], vkq = V[k * 12 + q];
          V[k * 12 + p] = c * vkp - s * vkq;
          V[k * 12 + q] = s * vkp + c * vkq;
        }
        __syncwarp();
        if (lane < 12) {
          const double apk = A[p * 12 + k], aqk = A[q * 12 + k];
          A[p * 12 + k] = c * apk - s * aqk;
          A[q * 12 + k] = s * apk + c * aqk;
        }
        __syncwarp();
      }
  }
  int m = 0;
  for (int i = 1; i < 12; ++i)
    if (A[i * 12 + i] < A[m * 12 + m]) m = i;
  double pv[12];
  for (int i = 0; i < 12; ++i) pv[i] = V[i * 12 + m];
  __syncwarp();
  pose::dlt_finish(pv, R, t);
}

// sum_i |project(R X_i + t) - uv_i|^2: one point per lane, summed in point order by every lane
__device__ double reproj_cost_warp(const double* X, const double* uv, int n, const double* R, const double* t, double fx,
                                   double fy, double cx, double cy, double* part, int lane) {
  __syncwarp();
  if (lane < n) {
    const double* x = X + 3 * lane;
    const double px = R[0] * x[0] + R[1] * x[1] + R[2] * x[2] + t[0];
    const double py = R[3] * x[0] + R[4] * x[1] + R[5] * x[2] + t[1];
    const double pz = R[6] * x[0] + R[7] * x[1] + R[8] * x[2] + t[2];
    const double du = fx * px / pz + cx - uv[2 * lane];
    const double dv = fy * py / pz + cy - uv[2 * lane + 1];
    part[lane] = du * du + dv * dv;
  }
  __syncwarp();
  double c = 0.0;
  for (int i = 0; i < n; ++i) c += part[i];
  return c;
}

__device__ double refine_lm_warp(const double* X, const double* uv, int n, double fx, double fy, double cx, double cy,
                                 double* R, double* t, double* sm, int lane) {
  double* J = sm;            // [16][14]: Ju[6], Jv[6], ru, rv
  double* AG = sm + 224;     // 21 lower-triangle entries of J^T J, then 6 of J^T r
  double* part = sm + 256;   // [16]
  double lam = 1e-3;
  double cost = reproj_cost_warp(X, uv, n, R, t, fx, fy, cx, cy, part, lane);
  for (int iter = 0; iter < 20; ++iter) {       // cv2 TermCriteria MAX_ITER, see pose::refine_lm
    __syncwarp();
    if (lane < n) {
      const double* x = X + 3 * lane;
      const double qx = R[0] * x[0] + R[1] * x[1] + R[2] * x[2];
      const double qy = R[3] * x[0] + R[4] * x[1] + R[5] * x[2];
      const double qz = R[6] * x[0] + R[7] * x[1] + R[8] * x[2];
      const double px = qx + t[0], py = qy + t[1], pz = qz + t[2];
      const double iz = 1.0 / pz;
      const double du[3] = {fx * iz, 0.0, -fx * px * iz * iz};
      const double dv[3] = {0.0, fy * iz, -fy * py * iz * iz};
      double* Jr = J + lane * 14;
      Jr[0] = du[1] * (-qz) + du[2] * qy;
      Jr[1] = du[0] * qz + du[2] * (-qx);
      Jr[2] = du[0] * (-qy) + du[1] * qx;
      Jr[6] = dv[1] * (-qz) + dv[2] * qy;
      Jr[7] = dv[0] * qz + dv[2] * (-qx);
      Jr[8] = dv[0] * (-qy) + dv[1] * qx;
      for (int kk = 0; kk < 3; ++kk) {
        Jr[3 + kk] = du[kk];
        Jr[9 + kk] = dv[kk];
      }
      Jr[12] = fx * px * iz + cx - uv[2 * lane];
      Jr[13] = fy * py * iz + cy - uv[2 * lane + 1];
    }
    __syncwarp();
    if (lane < 27) {
      double val = 0.0;
      if (lane < 21) {
        int a = 0;
        while ((a + 1) * (a + 2) / 2 <= lane) ++a;      // lower-triangle index -> (a, b), b <= a
        const int b = lane - a * (a + 1) / 2;
        for (int i = 0; i < n; ++i) val += J[i * 14 + a] * J[i * 14 + b] + J[i * 14 + 6 + a] * J[i * 14 + 6 + b];
      } else {
        const int a = lane - 21;
        for (int i = 0; i < n; ++i) val += J[i * 14 + a] * J[i * 14 + 12] + J[i * 14 + 6 + a] * J[i * 14 + 13];
      }
      AG[lane] = val;
    }
    __syncwarp();
    double A[36], g[6];
    for (int a = 0; a < 6; ++a) {
      g[a] = AG[21 + a];
      for (int b = 0; b <= a; ++b) {
        const double v = AG[a * (a + 1) / 2 + b];
        A[a * 6 + b] = v;
        A[b * 6 + a] = v;
      }
    }
    bool improved = false;
    double d[6], Rn[9], tn[3], cn = 0.0;
    for (int tr = 0; tr < 30; ++tr) {
      if (pose::solve6(A, g, lam, d)) {
        double E[9];
        pose::rodrigues(d, E);
        pose::mat3_mul(E, R, Rn);
        for (int kk = 0; kk < 3; ++kk) tn[kk] = t[kk] + d[3 + kk];
        cn = reproj_cost_warp(X, uv, n, Rn, tn, fx, fy, cx, cy, part, lane);
        if (cn == cn && cn <= cost && fabs(cn) < 1e300) {
          improved = true;
          break;
        }
      }
      lam *= 10.0;
    }
    if (!improved) break;
    double step = 0.0;
    for (int kk = 0; kk < 6; ++kk) step += d[kk] * d[kk];
    step = sqrt(step);
    for (int kk = 0; kk < 9; ++kk) R[kk] = Rn[kk];
    for (int kk = 0; kk < 3; ++kk) t[kk] = tn[kk];
    const double dec = cost - cn;
    cost = cn;
    lam = lam * 0.1;
    if (lam < 1e-12) lam = 1e-12;
    if (step < 1e-10 || dec <= 1e-28 * (cost > 1e-300 ? cost : 1e-300)) break;
  }
  return cost;
}

// pose::solve_and_shell, executed by a whole warp; every lane ends with the same PnPOut
__device__ void solve_and_shell_warp(const double* pts, int n_in, const float* obj_scale, const double* Kc, double width,
                                     double height, int visible_thresh, int opencv_return, pose::PnPOut* o, double* sm,
                                     int lane) {
  double V[24], X[48], uv[32];
  const int n = pose::pnp_collect(pts, n_in, obj_scale, V, X, uv);
  o->n_pts = n;
  o->status = CP_PNP_FEW_POINTS;
  if (n < 6) return;
  double R[9], t[3];
  dlt_init_warp(X, uv, n, Kc[0], Kc[4], Kc[2], Kc[5], R, t, sm, lane);
  const double cost = refine_lm_warp(X, uv, n, Kc[0], Kc[4], Kc[2], Kc[5], R, t, sm, lane);
  pose::pnp_finish(V, R, t, cost, n, Kc, width, height, visible_thresh, opencv_return, o);
}

__global__ void __launch_bounds__(256, 1) group_pose_kernel(const GroupArgs a) {
  const cp_decode_params& P = a.prm;
  const int b = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
  const int K = P.K, J = P.num_joints, H = P.out_h, W = P.out_w, HW = H * W;
  const int CH = P.num_classes + J;
  const float th = 0.1f;

  __shared__ float c_score[KM];
  __shared__ int c_ind[KM];
  __shared__ float c_bbox[KM][4];
  __shared__ float c_disp[KM][16];
  __shared__ float hmx[8][KM], hmy[8][KM], hms[8][KM];
  __shared__ double nb_bbox[KM][4];
  __shared__ double nb_score[KM];
  __shared__ int nb_perm[KM];
  __shared__ int s_n0, s_n1;

  float* dets = a.dets + (size_t)b * K * CP_DETS_RECORD;
  float* poses = a.poses + (size_t)b * K * CP_POSE_RECORD;
  const double* meta = a.meta + (size_t)b * CP_META_DOUBLES;

  // ---------------- phase A: centres (decode.py:83-109, 304-345)
  for (int k = tid; k < K; k += NT) {
    const int ind = a.peak_idx[((size_t)b * CH + 0) * K + k];
    const float score = a.peak_val[((size_t)b * CH + 0) * K + k];
    const float xs = (float)(ind % W), ys = (float)(ind / W);
    float* d = dets + (size_t)k * CP_DETS_RECORD;
    c_score[k] = score;
    c_ind[k] = ind;
    for (int j = 0; j < 2 * J; ++j) {
      float v = gatherf(a.h.hps, b, 2 * J, j, HW, ind) + ((j & 1) ? ys : xs);
      c_disp[k][j] = v;
      d[CP_D_KPS_DISP_MEAN + j] = v;
    }
    float cx, cy;
    if (a.h.reg) {
      cx = xs + gatherf(a.h.reg, b, 2, 0, HW, ind);
      cy = ys + gatherf(a.h.reg, b, 2, 1, HW, ind);
    } else {
      cx = xs + 0.5f;
      cy = ys + 0.5f;
    }
    float w = gatherf(a.h.wh, b, 2, 0, HW, ind), hgt = gatherf(a.h.wh, b, 2, 1, HW, ind);
    float bb[4] = {cx - w / 2.0f, cy - hgt / 2.0f, cx + w / 2.0f, cy + hgt / 2.0f};
    for (int t = 0; t < 4; ++t) {
      c_bbox[k][t] = bb[t];
      d[CP_D_BBOX + t] = bb[t];
    }
    d[CP_D_SCORE] = score;
    d[CP_D_CLS] = 0.0f;
    d[CP_D_IND] = (float)ind;
    for (int t = 0; t < 3; ++t) {
      d[CP_D_OBJ_SCALE + t] = a.h.scale ? gatherf(a.h.scale, b, 3, t, HW, ind) : 0.0f;
      d[CP_D_OBJ_SCALE_UNC + t] =
          a.h.scale_uncertainty ? sqrtf(expf(gatherf(a.h.scale_uncertainty, b, 3, t, HW, ind))) : 0.0f;
    }
    for (int t = 0; t < 2; ++t) d[CP_D_TRACKING + t] = a.h.tracking ? gatherf(a.h.tracking, b, 2, t, HW, ind) : 0.0f;
    for (int t = 0; t < 2 * J; ++t) {
      d[CP_D_TRACKING_HP + t] = a.h.tracking_hp ? gatherf(a.h.tracking_hp, b, 2 * J, t, HW, ind) : 0.0f;
      d[CP_D_KPS_DISP_STD + t] =
          a.h.hps_uncertainty ? sqrtf(expf(gatherf(a.h.hps_uncertainty, b, 2 * J, t, HW, ind))) * P.balance : 0.0f;
    }
  }
  // ---------------- phase B: per-joint heat-map peaks (decode.py:129-144)
  for (int i = tid; i < J * K; i += NT) {
    const int j = i / K, m = i - j * K;
    const int ind = a.peak_idx[((size_t)b * CH + P.num_classes + j) * K + m];
    float s = a.peak_val[((size_t)b * CH + P.num_classes + j) * K + m];
    float x = (float)(ind % W), y = (float)(ind / W);
    if (a.h.hp_offset) {
      x += gatherf(a.h.hp_offset, b, 2, 0, HW, ind);
      y += gatherf(a.h.hp_offset, b, 2, 1, HW, ind);
    } else {
      x += 0.5f;
      y += 0.5f;
    }
    if (!(s > th)) {
      s = -1.0f;
      x = SENT;
      y = SENT;
    }
    hmx[j][m] = x;
    hmy[j][m] = y;
    hms[j][m] = s;
  }
  __syncthreads();

  // ---------------- phase C: nearest peak per (centre, joint) and the gates (decode.py:147-252)
  for (int i = tid; i < K * J; i += NT) {
    const int k = i / J, j = i - k * J;
    const float rx = c_disp[k][2 * j], ry = c_disp[k][2 * j + 1];
    float best = INFINITY;
    int bi = 0;
    for (int m = 0; m < K; ++m) {
      float dx = rx - hmx[j][m], dy = ry - hmy[j][m];
      float dd = sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
      if (dd < best) {
        best = dd;
        bi = m;
      }
    }
    const float sx = hmx[j][bi], sy = hmy[j][bi], ss = hms[j][bi];
    const float l = c_bbox[k][0], t = c_bbox[k][1], r = c_bbox[k][2], bt = c_bbox[k][3];
    const float size = fmaxf(bt - t, r - l);
    const bool bad = (sx < l) || (sx > r) || (sy < t) || (sy > bt) || (ss < th) || (best > __fmul_rn(size, 0.3f));
    float kx = rx, ky = ry;
    if (P.rep_mode == 4) {
      kx = sx;
      ky = sy;
    } else if (P.rep_mode != 3 && !bad) {
      kx = sx;
      ky = sy;
    }
    float* d = dets + (size_t)k * CP_DETS_RECORD;
    d[CP_D_KPS + 2 * j] = kx;
    d[CP_D_KPS + 2 * j + 1] = ky;
    const bool ok2 = (sx > __fmul_rn(0.8f, l)) && (sx < __fmul_rn(1.2f, r)) && (sy > __fmul_rn(0.8f, t)) &&
                     (sy < __fmul_rn(1.2f, bt)) && (ss > th) && (best < __fmul_rn(size, 0.5f)) && (c_score[k] > th);
    float mean_x = SENT, mean_y = SENT, std_x = SENT, std_y = SENT, height = SENT;
    if ((P.rep_mode == 1 || P.rep_mode == 2) && ok2 && !(sx == SENT || sy == SENT)) {
      const float* hp = a.h.hm_hp + ((size_t)b * J + j) * HW;
      if (P.use_moments) {
        const int ran = 5;
        int r0, r1, q0, q1;
        py_slice((int)sy, (int)(sy + (float)(2 * ran + 1)), H + 2 * ran, &r0, &r1);
        py_slice((int)sx, (int)(sx + (float)(2 * ran + 1)), W + 2 * ran, &q0, &q1);
        const int nr = r1 - r0, nc = q1 - q0;
        double win[121];
        for (int rr = 0; rr < nr; ++rr)
          for (int cc = 0; cc < nc; ++cc) {
            int yy = r0 + rr - ran, xx = q0 + cc - ran;
            double v = 0.0;
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
              float raw = __ldg(hp + yy * W + xx);
              v = (double)(P.apply_sigmoid ? sigmoid_acc(raw) : raw);
            }
            win[rr * nc + cc] = v;
          }
        double hh, mx, my, wx, wy;
        if (pose::moments(win, nr, nc, &hh, &mx, &my, &wx, &wy)) {
          if (wx == 0.0) wx = 1e-10;   // least_squares(max_nfev=1) returns a strictly feasible start point
          if (wy == 0.0) wy = 1e-10;
          mean_x = (float)((double)sx + mx - (double)ran);
          mean_y = (float)((double)sy + my - (double)ran);
          std_x = (float)wx;
          std_y = (float)wy;
          height = (float)hh;
        }
      } else {
        int iy = (int)sy, ix = (int)sx;
        if (iy < 0) iy += H;
        if (ix < 0) ix += W;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) {   // the reference raises IndexError outside
          float raw = __ldg(hp + iy * W + ix);
          height = P.apply_sigmoid ? sigmoid_acc(raw) : raw;
          mean_x = sx;
          mean_y = sy;
          std_x = 1.0f;
          std_y = 1.0f;
        }
      }
    }
    d[CP_D_KPS_HM_MEAN + 2 * j] = mean_x;
    d[CP_D_KPS_HM_MEAN + 2 * j + 1] = mean_y;
    d[CP_D_KPS_HM_STD + 2 * j] = std_x;
    d[CP_D_KPS_HM_STD + 2 * j + 1] = std_y;
    d[CP_D_KPS_HM_HEIGHT + j] = height;
  }
  __syncthreads();

  // ---------------- output-map -> image affine (image.py:35-68 with rot = 0, float32 control points)
  const float cxf = (float)meta[0], cyf = (float)meta[1], swf = (float)meta[2];
  const float dwf = (float)W, dhf = (float)H;
  const float d0x = dwf * 0.5f, d0y = dhf * 0.5f;
  const float d1y = d0y + dwf * -0.5f;
  const float s1y = cyf + swf * -0.5f;
  const double aff = ((double)cyf - (double)s1y) / ((double)d0y - (double)d1y);
  const double tx = (double)cxf - aff * (double)d0x, ty = (double)cyf - aff * (double)d0y;
  const float ratio = (float)(meta[2] / (double)max(W, H));   // s / max(w, h), rounded once to float32
  const double img_w = meta[3], img_h = meta[4];

  // ---------------- phase D: score filter (object_pose.py:188-191); scores are sorted descending
  if (tid == 0) {
    int n0 = 0;
    while (n0 < K && c_score[n0] > P.vis_thresh) ++n0;
    s_n0 = n0;
  }
  __syncthreads();
  const int n0 = s_n0;
  for (int i = tid; i < n0; i += NT) {
    for (int p = 0; p < 2; ++p) {
      float x = c_bbox[i][2 * p], y = c_bbox[i][2 * p + 1];
      if (x == SENT && y == SENT) {
        nb_bbox[i][2 * p] = -10000.0;
        nb_bbox[i][2 * p + 1] = -10000.0;
      } else {
        nb_bbox[i][2 * p] = aff * (double)x + tx;
        nb_bbox[i][2 * p + 1] = aff * (double)y + ty;
      }
    }
    nb_score[i] = (double)c_score[i];
    nb_perm[i] = i;
  }
  __syncthreads();
  // ---------------- phase E: Gaussian soft-NMS (sequential by definition)
  if (tid == 0) {
    int n1 = n0;
    if (P.nms && n0 > 0) n1 = pose::soft_nms(&nb_bbox[0][0], nb_score, nb_perm, n0, (double)P.vis_thresh);
    s_n1 = n1;
    a.n_valid[b] = n1;
  }
  __syncthreads();
  const int n1 = s_n1;

  // ---------------- phase F: records, one thread per surviving detection
  for (int i = tid; i < n1; i += NT) {
    const int k = nb_perm[i];
    const float* d = dets + (size_t)k * CP_DETS_RECORD;
    float* o = poses + (size_t)i * CP_POSE_RECORD;
    o[CP_P_SCORE] = (float)nb_score[i];
    o[CP_P_CLS] = d[CP_D_CLS];
    o[CP_P_SRC_INDEX] = (float)k;
    for (int t = 0; t < 4; ++t) o[CP_P_BBOX + t] = (float)nb_bbox[i][t];
    o[CP_P_CT] = (float)((nb_bbox[i][0] + nb_bbox[i][2]) / 2.0);
    o[CP_P_CT + 1] = (float)((nb_bbox[i][1] + nb_bbox[i][3]) / 2.0);
    double kps[16], dmean[16], hmean[16];
    for (int j = 0; j < J; ++j) {
      const int offs[3] = {CP_D_KPS, CP_D_KPS_DISP_MEAN, CP_D_KPS_HM_MEAN};
      double* dst[3] = {kps, dmean, hmean};
      for (int q = 0; q < 3; ++q) {
        float x = d[offs[q] + 2 * j], y = d[offs[q] + 2 * j + 1];
        if (x == SENT && y == SENT) {
          dst[q][2 * j] = -10000.0;
          dst[q][2 * j + 1] = -10000.0;
        } else {
          dst[q][2 * j] = aff * (double)x + tx;
          dst[q][2 * j + 1] = aff * (double)y + ty;
        }
      }
    }
    for (int t = 0; t < 2 * J; ++t) {
      o[CP_P_KPS + t] = (float)kps[t];
      o[CP_P_KPS_DISP_MEAN + t] = (float)dmean[t];
      o[CP_P_KPS_HM_MEAN + t] = (float)hmean[t];
      o[CP_P_KPS_HM_STD + t] = __fmul_rn(__fmul_rn(d[CP_D_KPS_HM_STD + t], ratio), 0.32f);
      o[CP_P_KPS_DISP_STD + t] = __fmul_rn(__fmul_rn(d[CP_D_KPS_DISP_STD + t], ratio), 0.32f);
      o[CP_P_TRACKING_HP + t] = __fmul_rn(d[CP_D_TRACKING_HP + t], ratio);
    }
    for (int t = 0; t < J; ++t) o[CP_P_KPS_HM_HEIGHT + t] = d[CP_D_KPS_HM_HEIGHT + t];
    for (int t = 0; t < 3; ++t) {
      o[CP_P_OBJ_SCALE + t] = d[CP_D_OBJ_SCALE + t];
      o[CP_P_OBJ_SCALE_UNC + t] = d[CP_D_OBJ_SCALE_UNC + t];
    }
    for (int t = 0; t < 2; ++t) o[CP_P_TRACKING + t] = __fmul_rn(d[CP_D_TRACKING + t], ratio);
  }
  // ---------------- phase G: PnP, one WARP per surviving detection
  {
    extern __shared__ double pnp_scratch[];
    const int warp = tid >> 5, lane = tid & 31, NW = NT >> 5;
    double* sm = pnp_scratch + warp * PNP_SCRATCH;
    double* pts = sm + 288;
    for (int i = warp; i < n1; i += NW) {
      const int k = nb_perm[i];
      const float* d = dets + (size_t)k * CP_DETS_RECORD;
      float* o = poses + (size_t)i * CP_POSE_RECORD;
      pose::PnPOut po;
      po.status = CP_PNP_NOT_RUN;
      po.n_pts = 0;
      if (P.use_pnp) {
        const int n_in = (P.rep_mode == 1) ? 16 : 8;
        __syncwarp();
        if (lane < n_in) {
          // rep_mode 1: point 2j = displacement mean of joint j, point 2j+1 = heat-map mean; else point j = kps[j]
          const int j = (P.rep_mode == 1) ? (lane >> 1) : lane;
          const int off = (P.rep_mode == 1) ? ((lane & 1) ? CP_D_KPS_HM_MEAN : CP_D_KPS_DISP_MEAN) : CP_D_KPS;
          const float x = d[off + 2 * j], y = d[off + 2 * j + 1];
          double X0 = -10000.0, Y0 = -10000.0;
          if (!(x == SENT && y == SENT)) {
            X0 = aff * (double)x + tx;
            Y0 = aff * (double)y + ty;
          }
          pts[2 * lane] = X0;
          pts[2 * lane + 1] = Y0;
        }
        __syncwarp();
        float sc[3] = {d[CP_D_OBJ_SCALE], d[CP_D_OBJ_SCALE + 1], d[CP_D_OBJ_SCALE + 2]};
        solve_and_shell_warp(pts, n_in, sc, meta + 5, img_w, img_h, P.visible_thresh, P.opencv_return, &po, sm, lane);
      }
      if (lane == 0) {
        o[CP_P_STATUS] = (float)po.status;
        o[CP_P_NPTS] = (float)po.n_pts;
        const bool has_pose = (po.status == CP_PNP_OK || po.status == CP_PNP_INVISIBLE);
        const bool has_proj = has_pose || po.status == CP_PNP_BEHIND;
        for (int t = 0; t < 3; ++t) o[CP_P_LOCATION + t] = has_pose ? (float)po.loc[t] : 0.0f;
        for (int t = 0; t < 4; ++t) o[CP_P_QUAT + t] = has_pose ? (float)po.quat[t] : 0.0f;
        o[CP_P_REPROJ] = has_proj ? (float)po.reproj : 0.0f;
        for (int t = 0; t < 16; ++t) o[CP_P_PROJ_CUBOID + t] = has_proj ? (float)po.proj[t] : 0.0f;
        for (int t = 0; t < 27; ++t) o[CP_P_KPS_3D_CAM + t] = has_pose ? (float)po.kps3d[t] : 0.0f;
        for (int t = 0; t < 18; ++t) o[CP_P_KPS_PNP + t] = has_pose ? (float)po.kpspnp[t] : 0.0f;
      }
      __syncwarp();
    }
  }
  // zero the unused slots so the all-gathered tensor is deterministic
  for (int i = n1 * CP_POSE_RECORD + tid; i < K * CP_POSE_RECORD; i += NT) poses[i] = 0.0f;
}

struct WsLayout {
  size_t peak_val, peak_idx, dets, total;
};

WsLayout ws_layout(const cp_decode_params* p) {
  WsLayout w;
  size_t n = (size_t)p->batch * (p->num_classes + p->num_joints) * p->K;
  size_t off = 0;
  w.peak_val = off;
  off += (n * sizeof(float) + 255) / 256 * 256;
  w.peak_idx = off;
  off += (n * sizeof(int) + 255) / 256 * 256;
  w.dets = off;
  off += ((size_t)p->batch * p->K * CP_DETS_RECORD * sizeof(float) + 255) / 256 * 256;
  w.total = off;
  return w;
}

int validate(const cp_decode_params* p) {
  if (!p) return fail(CP_ERR_INVALID, "decode: null params");
  if (p->batch <= 0 || p->out_h <= 0 || p->out_w <= 0) return fail(CP_ERR_INVALID, "decode: bad shape");
  if (p->num_classes != 1) return fail(CP_ERR_INVALID, "decode: num_classes must be 1 (Objectron single-category heads)");
  if (p->num_joints != 8) return fail(CP_ERR_INVALID, "decode: num_joints must be 8");
  if (p->K <= 0 || p->K > CP_MAX_K) return fail(CP_ERR_INVALID, "decode: K must be in 1..128");
  if ((size_t)p->out_h * p->out_w < (size_t)p->K) return fail(CP_ERR_INVALID, "decode: map smaller than K");
  if ((size_t)p->out_h * p->out_w * 8 > 200 * 1024)
    return fail(CP_ERR_INVALID, "decode: head map too large for the shared-memory top-K (max 25600 cells)");
  if (p->rep_mode == 2)
    return fail(CP_ERR_INVALID, "decode: rep_mode 2 (random GMM sampling, base_detector.py:568-650) is not supported");
  if (p->rep_mode < 0 || p->rep_mode > 4) return fail(CP_ERR_INVALID, "decode: rep_mode must be 0, 1, 3 or 4");
  return CP_OK;
}

}  // namespace
}  // namespace cp

using namespace cp;

extern "C" {

size_t cp_decode_workspace_bytes(const cp_decode_params* prm) {
  if (validate(prm)) return 0;
  return ws_layout(prm).total;
}

int cp_decode_pnp(const cp_decode_params* prm, const cp_heads* heads, const double* meta, float* dets, float* poses,
                  int32_t* n_valid, void* workspace, size_t workspace_bytes, void* stream_) {
  int rc = validate(prm);
  if (rc) return rc;
  if (!heads || !meta || !poses || !n_valid || !workspace) return fail(CP_ERR_INVALID, "cp_decode_pnp: null argument");
  if (!heads->hm || !heads->wh || !heads->hps || !heads->hm_hp)
    return fail(CP_ERR_INVALID, "cp_decode_pnp: hm, wh, hps and hm_hp heads are required");
  if (prm->use_pnp && !heads->scale) return fail(CP_ERR_INVALID, "cp_decode_pnp: PnP needs the scale head");
  WsLayout w = ws_layout(prm);
  if (workspace_bytes < w.total) return fail(CP_ERR_INVALID, "cp_decode_pnp: workspace too small");
  cudaStream_t s = (cudaStream_t)stream_;
  char* ws = (char*)workspace;
  float* peak_val = (float*)(ws + w.peak_val);
  int* peak_idx = (int*)(ws + w.peak_idx);
  float* dets_buf = dets ? dets : (float*)(ws + w.dets);

  const int HW = prm->out_h * prm->out_w;
  const size_t smem = (size_t)HW * 2 * sizeof(float);
  static thread_local size_t configured = 0;
  if (smem > configured) {
    CP_CUDA_CHECK(cudaFuncSetAttribute(peaks_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  dim3 g1(prm->num_classes + prm->num_joints, prm->batch);
  peaks_topk_kernel<<<g1, TOPK_THREADS, smem, s>>>(heads->hm, heads->hm_hp, prm->num_classes, prm->num_joints,
                                                   prm->out_h, prm->out_w, prm->K, prm->apply_sigmoid, peak_val,
                                                   peak_idx);
  CP_LAUNCH_CHECK("peaks_topk_kernel");
  GroupArgs ga;
  ga.prm = *prm;
  ga.h = *heads;
  ga.meta = meta;
  ga.peak_val = peak_val;
  ga.peak_idx = peak_idx;
  ga.dets = dets_buf;
  ga.poses = poses;
  ga.n_valid = n_valid;
  static thread_local bool pose_configured = false;
  if (!pose_configured) {      // static (~29 KB) + dynamic (20 KB) shared memory crosses the 48 KB default
    CP_CUDA_CHECK(cudaFuncSetAttribute(group_pose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    pose_configured = true;
  }
  group_pose_kernel<<<prm->batch, 256, 8 * PNP_SCRATCH * sizeof(double), s>>>(ga);
  CP_LAUNCH_CHECK("group_pose_kernel");
  return CP_OK;
}

}  // extern "C"

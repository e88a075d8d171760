// CenterPoseTrack state on the device: Tracker.step (association, 32-state Kalman filter per object as eight 4-state
// filters, scale pool, second PnP with the filtered keypoints) and the rendering of the previous-frame heat maps.
// The logic is track_core.h (host-tested against the unmodified reference tracker, tests/test_track_core_host.py); this
// file is its parallel orchestration: one CTA per video stream.
//
// Reference (relative to /root/reference/src/lib): utils/tracker.py:15-302, detectors/base_detector.py:150-388
// (_get_additional_inputs), :502-544 (gaussian_fusion), :660-665 (tracker.step in run()).
#include <new>

#include "common.cuh"
#include "pnp_warp.cuh"
#include "track_core.h"

using namespace cp;
using namespace cp::track;

struct cp_tracker {
  cp_tracker_config cfg;
  Slot* slots[2] = {nullptr, nullptr};      // [streams][max_tracks], ping-pong: slots[cur] = current tracks
  int* n_tracks[2] = {nullptr, nullptr};    // [streams]
  int* id_count = nullptr;                  // [streams]
  int cur = 0;
};

namespace cp {
namespace {

constexpr int TRK_THREADS = 256;
constexpr int TRK_MAXK = CP_MAX_K;

struct StepArgs {
  Cfg cfg;
  int visible_thresh, opencv_return, T, K;
  const float* poses;       // [B, K, 192]
  const int* n_valid;       // [B]
  const double* meta;       // [B, 16]
  const Slot* old_slots;
  const int* old_n;
  Slot* new_slots;
  int* new_n;
  int* id_count;
  float* tracks_out;        // [B, T, 320]
  int* n_out;               // [B]
};

__global__ void __launch_bounds__(TRK_THREADS, 1) tracker_step_kernel(const StepArgs a) {
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, NW = TRK_THREADS / 32;
  __shared__ Entry entries[TRK_MAXK];
  __shared__ int det_idx[TRK_MAXK];
  __shared__ float fbuf[3 * 2 * TRK_MAXK];
  __shared__ int ibuf[4 * TRK_MAXK];
  __shared__ unsigned char taken[TRK_MAXK];
  __shared__ int s_n;
  __shared__ double pnp_sm[(TRK_THREADS / 32) * PNP_SCRATCH];

  const float* poses = a.poses + (size_t)b * a.K * CP_POSE_RECORD;
  const Slot* old = a.old_slots + (size_t)b * a.T;
  Slot* nxt = a.new_slots + (size_t)b * a.T;
  const int M = a.old_n[b];
  int nv = a.n_valid[b];
  if (nv > a.K) nv = a.K;
  if (nv < 0) nv = 0;

  // Steps 0-1 and the order of `ret` (serial, a few hundred operations)
  if (tid == 0) {
    int idc = a.id_count[b];
    s_n = plan_step(a.cfg, poses, nv, old, M, &idc, entries, a.T, det_idx, fbuf, ibuf, taken);
    a.id_count[b] = idc;
  }
  __syncthreads();
  const int n = s_n;
  // Steps 2-4: one thread per entry (gaussian_fusion, eight 4-state predict/update, scale pool)
  for (int e = tid; e < n; e += TRK_THREADS) {
    const Entry en = entries[e];
    if (en.kind == ENTRY_MATCHED)
      entry_matched(a.cfg, &nxt[e], &old[en.trk], poses + (size_t)en.det * CP_POSE_RECORD);
    else if (en.kind == ENTRY_NEW)
      entry_new(a.cfg, &nxt[e], poses + (size_t)en.det * CP_POSE_RECORD, en.id);
    else
      entry_lost(&nxt[e], &old[en.trk]);
  }
  __syncthreads();
  // Steps 5-6: read-out + second PnP, one WARP per entry (every lane computes the identical read-out)
  const double* meta = a.meta + (size_t)b * CP_META_DOUBLES;
  double* sm = pnp_sm + warp * PNP_SCRATCH;
  for (int e = warp; e < n; e += NW) {
    double mean[16], sd[16], conf_avg, sc[3], su[3];
    entry_readout(a.cfg, &nxt[e], mean, sd, &conf_avg, sc, su);
    pose::PnPOut po;
    po.status = CP_PNP_NOT_RUN;
    po.n_pts = 0;
    int in_boxes = 0;
    if (a.cfg.use_pnp && (a.cfg.kalman || a.cfg.scale_pool)) {
      double V[24];
      if (a.cfg.scale_pool)
        pose::cuboid_vertices_d(sc, V);
      else
        pose::cuboid_vertices(nxt[e].rec + CP_P_OBJ_SCALE, V);
      __syncwarp();
      solve_and_shell_warp_v(mean, 8, V, meta + 5, meta[3], meta[4], a.visible_thresh, a.opencv_return, &po, sm, lane);
      in_boxes = (po.status == CP_PNP_OK && conf_avg > 0.25) ? 1 : 0;
    } else {
      in_boxes = ((int)nxt[e].rec[CP_P_STATUS] == CP_PNP_OK && nxt[e].age == 1) ? 1 : 0;
    }
    __syncwarp();
    if (lane == 0) {
      slot_store_pose(&nxt[e], po);
      slot_store_pnp_kf(&nxt[e], po);
      write_track_record(&nxt[e], mean, sd, conf_avg, sc, su, &po, in_boxes,
                         a.tracks_out + ((size_t)b * a.T + e) * CP_TRACK_RECORD);
    }
    __syncwarp();
  }
  // unused output rows are zeroed so the tensor is deterministic
  float* tail = a.tracks_out + ((size_t)b * a.T + n) * CP_TRACK_RECORD;
  for (int i = tid; i < (a.T - n) * CP_TRACK_RECORD; i += TRK_THREADS) tail[i] = 0.f;
  if (tid == 0) {
    a.new_n[b] = n;
    a.n_out[b] = n;
  }
}

__global__ void tracker_reset_kernel(int* n0, int* n1, int* id_count, int streams, int index) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < streams && (index < 0 || i == index)) {
    n0[i] = 0;
    n1[i] = 0;
    id_count[i] = 0;
  }
}

// ---- previous-frame heat maps ----------------------------------------------------------------------------------------
struct RenderArgs {
  Cfg cfg;
  int T, inp_h, inp_w, render_hm_mode, render_hmhp_mode;
  double pre_thresh;
  const Slot* slots;
  const int* n;
  const double* meta;      // [B,16]: [3] original width, [4] original height
  const double* trans;     // [B,6]
  float* pre_hm;           // [B,1,h,w]
  float* pre_hm_hp;        // [B,8,h,w]
};

struct Patch {
  int x, y, r, live;
  double k;
};

// np.dot(t, [x, y, 1]) with a float32 point and the float64 2 x 3 matrix (utils/image.py:71-74)
__device__ __forceinline__ void affine_pt(const double* t, float x, float y, double* ox, double* oy) {
  *ox = t[0] * (double)x + t[1] * (double)y + t[2] * 1.0;
  *oy = t[3] * (double)x + t[4] * (double)y + t[5] * 1.0;
}

// one CTA per (track, stream): thread 0 derives the nine patches exactly like base_detector.py:213-315, all threads draw
__global__ void __launch_bounds__(256) tracker_render_kernel(const RenderArgs a) {
  const int t = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  if (t >= a.n[b]) return;
  __shared__ Patch pt[9];
  if (tid == 0) {
    for (int i = 0; i < 9; ++i) pt[i].live = 0;
    const Slot& s = a.slots[(size_t)b * a.T + t];
    const float* r = s.rec;
    const double* tr = a.trans + (size_t)b * 6;
    const double ori_w = a.meta[(size_t)b * CP_META_DOUBLES + 3], ori_h = a.meta[(size_t)b * CP_META_DOUBLES + 4];
    if (!((double)r[CP_P_SCORE] < a.pre_thresh)) {
      // _trans_bbox (base_detector.py:79-89): float32 box, transformed corners rounded back to float32, clipped
      double x0, y0, x1, y1;
      affine_pt(tr, r[CP_P_BBOX], r[CP_P_BBOX + 1], &x0, &y0);
      affine_pt(tr, r[CP_P_BBOX + 2], r[CP_P_BBOX + 3], &x1, &y1);
      float bx0 = (float)x0, by0 = (float)y0, bx1 = (float)x1, by1 = (float)y1;
      const float wm = (float)(a.inp_w - 1), hm = (float)(a.inp_h - 1);
      bx0 = fminf(fmaxf(bx0, 0.f), wm);
      bx1 = fminf(fmaxf(bx1, 0.f), wm);
      by0 = fminf(fmaxf(by0, 0.f), hm);
      by1 = fminf(fmaxf(by1, 0.f), hm);
      const float h = __fsub_rn(by1, by0), w = __fsub_rn(bx1, bx0);
      if (h > 0.f && w > 0.f) {
        double rad = gaussian_radius(ceil((double)h), ceil((double)w));
        int radius = (int)rad;
        if (radius < 0) radius = 0;
        const float cx = __fadd_rn(bx0, bx1) / 2.0f, cy = __fadd_rn(by0, by1) / 2.0f;
        pt[0].x = (int)cx;
        pt[0].y = (int)cy;
        pt[0].r = radius;
        pt[0].k = a.render_hm_mode == 1 ? (double)r[CP_P_SCORE] : 1.0;
        pt[0].live = 1;
        // keypoints: normalised 9-point sets, entry 0 is the centre (base_detector.py:238-251)
        double px[8], py[8];
        bool have = true;
        const int mode = a.render_hmhp_mode;
        if (mode == 0 || mode == 1) {       // kps_ori: the detection's own keypoints, normalised
          for (int j = 0; j < 8; ++j) {
            px[j] = ((double)r[CP_P_KPS + 2 * j] / ori_w) * ori_w;
            py[j] = ((double)r[CP_P_KPS + 2 * j + 1] / ori_h) * ori_h;
          }
        } else if (a.cfg.kalman || a.cfg.scale_pool) {
          // kps_pnp_kf when the filtered PnP returned a tuple.  The reference's fall-back (kps_mean_kf[1:]: seven PIXEL
          // coordinates multiplied by the image size) can never land inside the image: nothing is drawn
          have = s.has_pnp_kf != 0;
          for (int j = 0; j < 8 && have; ++j) {
            px[j] = (double)s.kps_pnp_kf[2 * (j + 1)] * ori_w;
            py[j] = (double)s.kps_pnp_kf[2 * (j + 1) + 1] * ori_h;
          }
        } else {
          // 'kps_pnp' of the first PnP, or zeros when that failed (base_detector.py:248-253)
          const bool pose = ((int)r[CP_P_STATUS] == CP_PNP_OK || (int)r[CP_P_STATUS] == CP_PNP_INVISIBLE);
          for (int j = 0; j < 8; ++j) {
            px[j] = pose ? (double)r[CP_P_KPS_PNP + 2 * (j + 1)] * ori_w : 0.0;
            py[j] = pose ? (double)r[CP_P_KPS_PNP + 2 * (j + 1) + 1] * ori_h : 0.0;
          }
        }
        if (have) {
          for (int j = 0; j < 8; ++j) {
            Patch& q = pt[1 + j];
            q.live = 0;
            // COCO-style visibility, int64 truncation, affine of the truncated point, truncation again
            const bool outside = px[j] >= ori_w || px[j] < 0 || py[j] < 0 || py[j] >= ori_h;
            if (outside) continue;
            const long long ix = (long long)px[j], iy = (long long)py[j];
            double ax, ay;
            affine_pt(tr, (float)ix, (float)iy, &ax, &ay);
            const long long jx = (long long)ax, jy = (long long)ay;
            if (!(jx >= 0 && jx < a.inp_w && jy >= 0 && jy < a.inp_h)) continue;
            double k = 1.0;
            if (mode == 0 || mode == 2) {
              const double rd = a.cfg.hps_uncertainty ? s.fus_std[2 * j] : (double)r[CP_P_KPS_HM_STD + 2 * j];
              if (!((int)rd > 0)) continue;                   // radius_detector[j, 0] > 0 (int32 truncation)
              if (a.cfg.kalman && s.has_kf) {
                const double std_c = sqrt(s.f.P[j][0] + s.f.P[j][5]);
                k = 1.0 - pow(exp(log(0.15) / (a.cfg.conf_lo - a.cfg.conf_hi)), std_c - a.cfg.conf_hi);
                if (k < 0.0) k = 0.0;
              } else if (a.cfg.hps_uncertainty) {
                const double std_c = sqrt(s.fus_std[2 * j] + s.fus_std[2 * j + 1]);
                k = 1.0 - pow(exp(log(0.15) / (a.cfg.conf_lo - a.cfg.conf_hi)), std_c - a.cfg.conf_hi);
                if (k < 0.0) k = 0.0;
              } else {
                k = (double)r[CP_P_KPS_HM_HEIGHT + j];
              }
            }
            q.x = (int)jx;
            q.y = (int)jy;
            q.r = radius;
            q.k = k;
            q.live = 1;
          }
        }
      }
    }
  }
  __syncthreads();
  const size_t plane = (size_t)a.inp_h * a.inp_w;
  for (int i = 0; i < 9; ++i) {
    if (!(pt[i].live & 1)) continue;
    const Patch q = pt[i];
    float* map = (i == 0) ? a.pre_hm + (size_t)b * plane : a.pre_hm_hp + ((size_t)b * 8 + (i - 1)) * plane;
    // draw_umich_gaussian (utils/image.py:135-150): the patch clipped to the map, np.maximum compositing
    const int left = min(q.x, q.r), right = min(a.inp_w - q.x, q.r + 1);
    const int top = min(q.y, q.r), bottom = min(a.inp_h - q.y, q.r + 1);
    const int pw = left + right, ph = top + bottom;
    if (pw <= 0 || ph <= 0) continue;
    for (int e = tid; e < pw * ph; e += blockDim.x) {
      const int dy = e / pw - top, dx = e % pw - left;
      const float v = umich_value(dx, dy, q.r, q.k);
      // values are >= 0: the unsigned order of the bit patterns is the float order
      atomicMax(reinterpret_cast<unsigned int*>(map + (size_t)(q.y + dy) * a.inp_w + (q.x + dx)), __float_as_uint(v));
    }
  }
}

Cfg make_cfg(const cp_tracker_config& c) {
  Cfg g;
  g.kalman = c.kalman;
  g.scale_pool = c.scale_pool;
  g.use_pnp = c.use_pnp;
  g.hps_uncertainty = c.hps_uncertainty;
  g.max_age = c.max_age;
  g.new_thresh = (double)c.new_thresh;
  g.R = (double)c.R;
  g.conf_lo = (double)c.conf_lo;
  g.conf_hi = (double)c.conf_hi;
  return g;
}

}  // namespace
}  // namespace cp

extern "C" {

int cp_tracker_create(const cp_tracker_config* cfg, cp_tracker** out) {
  if (!cfg || !out) return fail(CP_ERR_INVALID, "cp_tracker_create: null argument");
  if (cfg->streams <= 0 || cfg->max_tracks <= 0 || cfg->max_tracks > CP_MAX_K)
    return fail(CP_ERR_INVALID, "cp_tracker_create: streams must be > 0 and max_tracks in 1..128");
  if (cfg->conf_lo == cfg->conf_hi) return fail(CP_ERR_INVALID, "cp_tracker_create: conf_border needs two distinct values");
  cp_tracker* t = new (std::nothrow) cp_tracker();
  if (!t) return fail(CP_ERR_INVALID, "cp_tracker_create: out of host memory");
  t->cfg = *cfg;
  struct DeviceGuard {
    int prev = -1;
    ~DeviceGuard() {
      if (prev >= 0) cudaSetDevice(prev);
    }
  } guard;
  cudaError_t e = cudaGetDevice(&guard.prev);
  if (e == cudaSuccess) e = cudaSetDevice(cfg->device);
  const size_t ns = (size_t)cfg->streams * cfg->max_tracks;
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
    e = cudaMalloc(&t->slots[i], ns * sizeof(Slot));
    if (e == cudaSuccess) e = cudaMalloc(&t->n_tracks[i], sizeof(int) * cfg->streams);
    if (e == cudaSuccess) e = cudaMemset(t->n_tracks[i], 0, sizeof(int) * cfg->streams);
  }
  if (e == cudaSuccess) e = cudaMalloc(&t->id_count, sizeof(int) * cfg->streams);
  if (e == cudaSuccess) e = cudaMemset(t->id_count, 0, sizeof(int) * cfg->streams);
  if (e != cudaSuccess) {
    cp_tracker_destroy(t);
    return fail(CP_ERR_CUDA, std::string("cp_tracker_create: ") + cudaGetErrorString(e));
  }
  *out = t;
  return CP_OK;
}

int cp_tracker_destroy(cp_tracker* t) {
  if (!t) return CP_OK;
  for (int i = 0; i < 2; ++i) {
    if (t->slots[i]) cudaFree(t->slots[i]);
    if (t->n_tracks[i]) cudaFree(t->n_tracks[i]);
  }
  if (t->id_count) cudaFree(t->id_count);
  delete t;
  return CP_OK;
}

int cp_tracker_reset(cp_tracker* t, int32_t index, void* stream) {
  if (!t) return fail(CP_ERR_INVALID, "cp_tracker_reset: null tracker");
  if (index >= t->cfg.streams) return fail(CP_ERR_INVALID, "cp_tracker_reset: stream index out of range");
  tracker_reset_kernel<<<(t->cfg.streams + 127) / 128, 128, 0, (cudaStream_t)stream>>>(t->n_tracks[0], t->n_tracks[1],
                                                                                         t->id_count, t->cfg.streams, index);
  CP_LAUNCH_CHECK("tracker_reset_kernel");
  return CP_OK;
}

int cp_tracker_step(cp_tracker* t, int32_t batch, const float* poses, const int32_t* n_valid, int32_t K, const double* meta,
                    float* tracks_out, int32_t* n_tracks, void* stream) {
  if (!t || !poses || !n_valid || !meta || !tracks_out || !n_tracks) return fail(CP_ERR_INVALID, "cp_tracker_step: null argument");
  if (batch <= 0 || batch > t->cfg.streams) return fail(CP_ERR_INVALID, "cp_tracker_step: batch exceeds the tracker's streams");
  if (K <= 0 || K > CP_MAX_K) return fail(CP_ERR_INVALID, "cp_tracker_step: K must be in 1..128");
  StepArgs a;
  a.cfg = make_cfg(t->cfg);
  a.visible_thresh = t->cfg.visible_thresh;
  a.opencv_return = t->cfg.opencv_return;
  a.T = t->cfg.max_tracks;
  a.K = K;
  a.poses = poses;
  a.n_valid = n_valid;
  a.meta = meta;
  a.old_slots = t->slots[t->cur];
  a.old_n = t->n_tracks[t->cur];
  a.new_slots = t->slots[t->cur ^ 1];
  a.new_n = t->n_tracks[t->cur ^ 1];
  a.id_count = t->id_count;
  a.tracks_out = tracks_out;
  a.n_out = n_tracks;
  tracker_step_kernel<<<batch, TRK_THREADS, 0, (cudaStream_t)stream>>>(a);
  CP_LAUNCH_CHECK("tracker_step_kernel");
  t->cur ^= 1;           // calls are issued in frame order on one stream
  return CP_OK;
}

int cp_tracker_render(cp_tracker* t, int32_t batch, const double* meta, const double* trans_input, int32_t inp_h,
                      int32_t inp_w, float* pre_hm, float* pre_hm_hp, void* stream) {
  if (!t || !meta || !trans_input || !pre_hm || !pre_hm_hp) return fail(CP_ERR_INVALID, "cp_tracker_render: null argument");
  if (batch <= 0 || batch > t->cfg.streams || inp_h <= 0 || inp_w <= 0) return fail(CP_ERR_INVALID, "cp_tracker_render: bad shape");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t plane = (size_t)inp_h * inp_w;
  CP_CUDA_CHECK(cudaMemsetAsync(pre_hm, 0, sizeof(float) * plane * batch, s));
  CP_CUDA_CHECK(cudaMemsetAsync(pre_hm_hp, 0, sizeof(float) * plane * 8 * batch, s));
  RenderArgs a;
  a.cfg = make_cfg(t->cfg);
  a.T = t->cfg.max_tracks;
  a.inp_h = inp_h;
  a.inp_w = inp_w;
  a.render_hm_mode = t->cfg.render_hm_mode;
  a.render_hmhp_mode = t->cfg.render_hmhp_mode;
  a.pre_thresh = (double)t->cfg.pre_thresh;
  a.slots = t->slots[t->cur];
  a.n = t->n_tracks[t->cur];
  a.meta = meta;
  a.trans = trans_input;
  a.pre_hm = pre_hm;
  a.pre_hm_hp = pre_hm_hp;
  dim3 grid(t->cfg.max_tracks, batch);
  tracker_render_kernel<<<grid, 256, 0, s>>>(a);
  CP_LAUNCH_CHECK("tracker_render_kernel");
  return CP_OK;
}

}  // extern "C"

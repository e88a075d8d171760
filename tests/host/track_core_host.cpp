// TEST INFRASTRUCTURE ONLY: compiles centerpose_b200/csrc/track_core.h (+ pose_core.h) for the host so that the tracker
// logic the CUDA kernel runs (tracker.cu) can be checked on the CPU against tests/golden/tracker_seq.json.  The
// orchestration below is the serial statement of tracker_step_kernel.  Never loaded by the product.
#include <vector>

#include "../../centerpose_b200/csrc/track_core.h"

using namespace cp;
using namespace cp::track;

struct HostTracker {
  Cfg cfg;
  int visible_thresh, opencv_return, max_tracks;
  int id_count = 0;
  std::vector<Slot> tracks;
};

extern "C" {

void* trk_create(int kalman, int scale_pool, int use_pnp, int hps_uncertainty, int max_age, double new_thresh, double R,
                 double conf_lo, double conf_hi, int visible_thresh, int opencv_return, int max_tracks) {
  HostTracker* t = new HostTracker();
  t->cfg = Cfg{kalman, scale_pool, use_pnp, hps_uncertainty, max_age, new_thresh, R, conf_lo, conf_hi};
  t->visible_thresh = visible_thresh;
  t->opencv_return = opencv_return;
  t->max_tracks = max_tracks;
  return t;
}
void trk_destroy(void* h) { delete (HostTracker*)h; }
void trk_reset(void* h) {
  HostTracker* t = (HostTracker*)h;
  t->id_count = 0;
  t->tracks.clear();
}

// Tracker.step for one video stream; out: [max_tracks][CP_TRACK_RECORD]; returns the number of tracks
int trk_step(void* h, const float* poses, int n_valid, const double* cam, double width, double height, float* out) {
  HostTracker* t = (HostTracker*)h;
  const int M = (int)t->tracks.size(), K = n_valid;
  std::vector<Entry> entries(t->max_tracks);
  std::vector<int> det_idx(K + 1), ibuf(2 * K + 2 * M + 4);
  std::vector<float> fbuf(3 * (K + M) + 4);
  std::vector<unsigned char> taken(M + 1);
  const int n = plan_step(t->cfg, poses, n_valid, t->tracks.data(), M, &t->id_count, entries.data(), t->max_tracks,
                          det_idx.data(), fbuf.data(), ibuf.data(), taken.data());
  std::vector<Slot> next(n);
  for (int e = 0; e < n; ++e) {
    const Entry& en = entries[e];
    if (en.kind == ENTRY_MATCHED)
      entry_matched(t->cfg, &next[e], &t->tracks[en.trk], poses + (size_t)en.det * CP_POSE_RECORD);
    else if (en.kind == ENTRY_NEW)
      entry_new(t->cfg, &next[e], poses + (size_t)en.det * CP_POSE_RECORD, en.id);
    else
      entry_lost(&next[e], &t->tracks[en.trk]);
  }
  for (int e = 0; e < n; ++e) {
    double mean[16], sd[16], conf_avg, sc[3], su[3];
    entry_readout(t->cfg, &next[e], mean, sd, &conf_avg, sc, su);
    pose::PnPOut po;
    po.status = CP_PNP_NOT_RUN;
    po.n_pts = 0;
    int in_boxes = 0;
    if (t->cfg.use_pnp && (t->cfg.kalman || t->cfg.scale_pool)) {
      double V[24];
      if (t->cfg.scale_pool) {
        pose::cuboid_vertices_d(sc, V);
      } else {
        pose::cuboid_vertices(next[e].rec + CP_P_OBJ_SCALE, V);
      }
      pose::solve_and_shell_v(mean, 8, V, cam, width, height, t->visible_thresh, t->opencv_return, &po);
      slot_store_pose(&next[e], po);
      slot_store_pnp_kf(&next[e], po);
      in_boxes = (po.status == CP_PNP_OK && conf_avg > 0.25) ? 1 : 0;
    } else {
      in_boxes = ((int)next[e].rec[CP_P_STATUS] == CP_PNP_OK && next[e].age == 1) ? 1 : 0;   // `boxes` passes through
    }
    write_track_record(&next[e], mean, sd, conf_avg, sc, su, &po, in_boxes, out + (size_t)e * CP_TRACK_RECORD);
  }
  t->tracks.swap(next);
  return n;
}

// previous-frame heat-map primitives (checked against utils/image.py on the host)
double trk_gaussian_radius(double h, double w) { return gaussian_radius(h, w); }
float trk_umich_value(int dx, int dy, int r, double k) { return umich_value(dx, dy, r, k); }
}

// TEST INFRASTRUCTURE ONLY: compiles centerpose_b200/csrc/pose_core.h for the
// host so the per-detection math that runs inside the CUDA decode kernel can
// be checked against the oracle without a GPU.  Built by tests/conftest.py
// into tests/host/_build/libpose_core_host.so; never loaded by the product.
#include "../../centerpose_b200/csrc/pose_core.h"

extern "C" {

void host_solve_and_shell(const double* pts, int n_in, const float* obj_scale, const double* Kc, double width,
                          double height, int visible_thresh, int opencv_return, double* out /*[64+]*/, int* status,
                          int* n_pts) {
  cp::pose::PnPOut o;
  o.status = 0;
  o.n_pts = 0;
  for (int i = 0; i < 3; ++i) o.loc[i] = 0;
  for (int i = 0; i < 4; ++i) o.quat[i] = 0;
  o.reproj = 0;
  for (int i = 0; i < 16; ++i) o.proj[i] = 0;
  for (int i = 0; i < 27; ++i) o.kps3d[i] = 0;
  for (int i = 0; i < 18; ++i) o.kpspnp[i] = 0;
  cp::pose::solve_and_shell(pts, n_in, obj_scale, Kc, width, height, visible_thresh, opencv_return, &o);
  *status = o.status;
  *n_pts = o.n_pts;
  int p = 0;
  for (int i = 0; i < 3; ++i) out[p++] = o.loc[i];
  for (int i = 0; i < 4; ++i) out[p++] = o.quat[i];
  out[p++] = o.reproj;
  for (int i = 0; i < 16; ++i) out[p++] = o.proj[i];
  for (int i = 0; i < 27; ++i) out[p++] = o.kps3d[i];
  for (int i = 0; i < 18; ++i) out[p++] = o.kpspnp[i];
}

int host_soft_nms(double* bbox, double* score, int* perm, int n, double threshold) {
  return cp::pose::soft_nms(bbox, score, perm, n, threshold);
}

int host_moments(const double* w, int nr, int nc, double* out5) {
  double h, x, y, wx, wy;
  bool ok = cp::pose::moments(w, nr, nc, &h, &x, &y, &wx, &wy);
  out5[0] = h; out5[1] = x; out5[2] = y; out5[3] = wx; out5[4] = wy;
  return ok ? 1 : 0;
}

void host_cuboid_vertices(const float* scale, double* V) { cp::pose::cuboid_vertices(scale, V); }
}

// Warp-cooperative PnP (shared by decode.cu and tracker.cu): see the comment below.  Device-only; the serial,
// host-tested statement of the same math is pose_core.h.
#pragma once
#include "../../include/centerpose_b200.h"
#include "pose_core.h"

namespace cp {
namespace {

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
  if (n < 4) return;
  if (n < 6) {          // EPnP (rare path): evaluated redundantly by every lane, identical results
    pose::pnp_few_points(V, X, uv, n, Kc, width, height, visible_thresh, opencv_return, o);
    return;
  }
  double R[9], t[3];
  dlt_init_warp(X, uv, n, Kc[0], Kc[4], Kc[2], Kc[5], R, t, sm, lane);
  const double cost = refine_lm_warp(X, uv, n, Kc[0], Kc[4], Kc[2], Kc[5], R, t, sm, lane);
  pose::pnp_finish(V, R, t, cost, n, Kc, width, height, visible_thresh, opencv_return, o);
}

// same with the cuboid vertices given (tracker: second PnP with the pooled float64 scale)
__device__ void solve_and_shell_warp_v(const double* pts, int n_in, const double* V, const double* Kc, double width,
                                       double height, int visible_thresh, int opencv_return, pose::PnPOut* o, double* sm,
                                       int lane) {
  double X[48], uv[32];
  const int n = pose::pnp_collect_v(pts, n_in, V, X, uv);
  o->n_pts = n;
  o->status = CP_PNP_FEW_POINTS;
  if (n < 4) return;
  if (n < 6) {          // EPnP (rare path): evaluated redundantly by every lane, identical results
    pose::pnp_few_points(V, X, uv, n, Kc, width, height, visible_thresh, opencv_return, o);
    return;
  }
  double R[9], t[3];
  dlt_init_warp(X, uv, n, Kc[0], Kc[4], Kc[2], Kc[5], R, t, sm, lane);
  const double cost = refine_lm_warp(X, uv, n, Kc[0], Kc[4], Kc[2], Kc[5], R, t, sm, lane);
  pose::pnp_finish(V, R, t, cost, n, Kc, width, height, visible_thresh, opencv_return, o);
}

}  // namespace
}  // namespace cp

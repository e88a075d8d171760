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

// Eigen-decomposition of the symmetric 12 x 12 matrix A = sm[0,144) (destroyed: eigenvalues end on the diagonal), eigenvectors
// in the columns of V = sm[144,288) (must hold the identity on entry), by one warp.  sm[288,306) is scratch.
__device__ void jacobi12_warp(double* sm, int lane) {
  double* A = sm;
  double* V = sm + 144;
  // Cyclic Jacobi in the ROUND-ROBIN order: the 66 index pairs of a sweep are 11 rounds of 6 disjoint pairs, and the six
  // rotations of a round commute, so a round is ONE parallel step -- lanes 0-5 compute the six angles, then 144 lane-tasks
  // rotate the columns of A and V (A J, V J) and 72 the rows of A (J^T A).  The row-cyclic order of pose::dlt_init (the
  // host-tested statement, pose_core.h) converges to the same eigenvectors; the device needs ~2 K instead of ~9.2 K
  // dependent instructions per sweep (ncu: the serial sweep was 80 % of the 100 K instructions of a PnP solve, each
  // stalled ~7 cycles on its predecessor).  The lane -> (pair slot, row) assignment of a task is fixed, so it is decoded
  // once; angles use rsqrt / reciprocal (1 ulp; Jacobi is self-correcting) instead of two divisions and two square roots.
  double* rcs = sm + 288;                               // [6][2]: c, s of the round (the image points were consumed above)
  int* rpq = reinterpret_cast<int*>(sm + 300);          // [6][2]: p, q
  int c_off[5], r_pr[3], r_k[3];
  bool c_on[5], r_on[3];
#pragma unroll
  for (int it = 0; it < 5; ++it) {
    const int task = lane + 32 * it;
    c_on[it] = task < 144;
    const int pr = (task % 72) / 12, k = task % 12;
    c_off[it] = (task < 72 ? 0 : 144) + k * 12 + (pr << 16);      // element offset of row k in A | V, pair slot in the high half
  }
#pragma unroll
  for (int it = 0; it < 3; ++it) {
    const int task = lane + 32 * it;
    r_on[it] = task < 72;
    r_pr[it] = (task % 72) / 12;
    r_k[it] = task % 12;
  }
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int e = lane; e < 144; e += 32) {
      const int a = e / 12, b = e - a * 12;
      const double v = A[e] * A[e];
      if (a == b) diag += v;
      else if (a < b) off += v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      off += __shfl_xor_sync(0xffffffffu, off, o);
      diag += __shfl_xor_sync(0xffffffffu, diag, o);
    }
    if (off <= 1e-60 * diag || off == 0.0) break;
    for (int r = 0; r < 11; ++r) {
      if (lane < 6) {
        // round r of the tournament on 12 players: (11, r) and ((r + i) mod 11, (r - i) mod 11), i = 1..5
        int p = lane == 0 ? 11 : (r + lane) % 11, q = lane == 0 ? r : (r - lane + 11) % 11;
        if (p > q) {
          const int t0 = p;
          p = q;
          q = t0;
        }
        const double apq = A[p * 12 + q];
        double c = 1.0, s = 0.0;
        if (apq != 0.0) {
          const double app = A[p * 12 + p], aqq = A[q * 12 + q];
          const double theta = (aqq - app) / (2.0 * apq);
          double tt;
          if (fabs(theta) > 1e100) {
            tt = 0.5 / theta;                          // 1 / (|theta| + sqrt(theta^2 + 1)) without the overflow
          } else {
            const double r1 = theta * theta + 1.0;
            tt = copysign(__drcp_rn(fabs(theta) + r1 * rsqrt(r1)), theta);
          }
          c = rsqrt(tt * tt + 1.0);
          s = tt * c;
        }
        rpq[lane * 2] = p;
        rpq[lane * 2 + 1] = q;
        rcs[lane * 2] = c;
        rcs[lane * 2 + 1] = s;
      }
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 5; ++it) {                  // columns p, q of A (tasks 0-71) and of V (72-143)
        if (c_on[it]) {
          const int pr = c_off[it] >> 16;
          double* row = sm + (c_off[it] & 0xffff);
          const int p = rpq[pr * 2], q = rpq[pr * 2 + 1];
          const double c = rcs[pr * 2], s = rcs[pr * 2 + 1];
          const double xp = row[p], xq = row[q];
          row[p] = c * xp - s * xq;
          row[q] = s * xp + c * xq;
        }
      }
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 3; ++it) {                  // rows p, q of A
        if (r_on[it]) {
          const int pr = r_pr[it], k = r_k[it];
          const int p = rpq[pr * 2], q = rpq[pr * 2 + 1];
          const double c = rcs[pr * 2], s = rcs[pr * 2 + 1];
          const double xp = A[p * 12 + k], xq = A[q * 12 + k];
          A[p * 12 + k] = c * xp - s * xq;
          A[q * 12 + k] = s * xp + c * xq;
        }
      }
      __syncwarp();
    }
  }
}

__device__ void dlt_init_warp(const double* X, const double* uv, int n, double fx, double fy, double cx, double cy,
                              double* R, double* t, double* sm, int lane) {
  double* A = sm;
  double* V = sm + 144;
  // M^T M of the 2n x 12 DLT matrix (rows [h 0 -x h], [0 h -y h], h = (X, Y, Z, 1), (x, y) the normalised image point)
  // is a 3 x 3 arrangement of 4 x 4 weighted moment matrices of h:  [H 0 -Hx; 0 H -Hy; -Hx -Hy Hxx+yy].  The points are
  // staged once in shared memory and the 4 x 16 moments summed by 2 entries per lane -- the first version evaluated all
  // 144 entries with per-entry conditionals on local arrays (6.9 K of the 57 K instructions of a solve).
  double* P = V;                        // [n][6]: X, Y, Z, 1, x, y (consumed before V is initialised)
  double* S = sm + 240;                 // [4][4][4]: weights 1, x, y, x^2 + y^2
  if (lane < n) {
    P[6 * lane + 0] = X[3 * lane];
    P[6 * lane + 1] = X[3 * lane + 1];
    P[6 * lane + 2] = X[3 * lane + 2];
    P[6 * lane + 3] = 1.0;
    P[6 * lane + 4] = (uv[2 * lane] - cx) / fx;
    P[6 * lane + 5] = (uv[2 * lane + 1] - cy) / fy;
  }
  __syncwarp();
  for (int e = lane; e < 64; e += 32) {
    const int w = e >> 4, i = (e >> 2) & 3, j = e & 3;
    double acc = 0.0;
    for (int k = 0; k < n; ++k) {
      const double* pk = P + 6 * k;
      const double x = pk[4], y = pk[5];
      const double wt = w == 0 ? 1.0 : (w == 1 ? x : (w == 2 ? y : x * x + y * y));
      acc += wt * (pk[i] * pk[j]);
    }
    S[e] = acc;
  }
  __syncwarp();
  for (int e = lane; e < 144; e += 32) {
    const int a = e / 12, b = e - a * 12;
    const int ba = a >> 2, bb = b >> 2, ij = (a & 3) * 4 + (b & 3);
    double v = 0.0;
    if (ba == bb)
      v = S[(ba == 2 ? 48 : 0) + ij];
    else if (ba == 2 || bb == 2)
      v = -S[((ba == 2 ? bb : ba) == 0 ? 16 : 32) + ij];
    A[e] = v;
  }
  __syncwarp();
  for (int e = lane; e < 144; e += 32) V[e] = (e / 12 == e % 12) ? 1.0 : 0.0;
  __syncwarp();
  jacobi12_warp(sm, lane);
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

// pose::pnp_few_points (EPnP, 4 - 5 valid points) by a whole warp.  The serial version spent 2.4 ms per solve in the
// 12 x 12 Jacobi on per-lane local arrays (measured: decode 0.18 -> 2.5 ms per frame when every object had 5 key points;
// in tracking the filter confidence hides key points regularly).  Here M^T M is built 4.5 entries per lane in shared
// memory and decomposed by jacobi12_warp; the small dependent rest (betas, Gauss-Newton, absolute orientation) is
// evaluated by every lane from identical inputs.  The null space of M is degenerate below 6 points, so -- exactly as
// between pose_core.h and LAPACK -- the basis, and with noisy points the pose, depends on the
// rotation order; consistent points give the unique answer.
__device__ void pnp_few_points_warp(const double* V, const double* X, const double* uv, int n, const double* Kc, double width,
                                    double height, int visible_thresh, int opencv_return, pose::PnPOut* o, double* sm,
                                    int lane) {
  pose::EpnpPre P;
  if (!pose::epnp_prepare(X, n, &P)) {
    o->status = CP_PNP_SOLVER_FAIL;
    return;
  }
  __syncwarp();
  for (int e = lane; e < 144; e += 32) {
    const int a = e / 12, b = e - a * 12;
    sm[e] = pose::epnp_mtm_entry(P, uv, n, Kc[0], Kc[4], Kc[2], Kc[5], a, b);
    sm[144 + e] = (a == b) ? 1.0 : 0.0;
  }
  __syncwarp();
  jacobi12_warp(sm, lane);
  __syncwarp();
  double R[9], t[3];
  const double err = pose::epnp_finish(X, uv, n, Kc[0], Kc[4], Kc[2], Kc[5], P, sm, sm + 144, R, t);
  __syncwarp();
  if (!(err < 1e299)) {
    o->status = CP_PNP_SOLVER_FAIL;
    return;
  }
  pose::pnp_finish(V, R, t, pose::reproj_cost(X, uv, n, R, t, Kc[0], Kc[4], Kc[2], Kc[5]), n, Kc, width, height, visible_thresh,
                   opencv_return, o);
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
  if (n < 6) {          // EPnP: 4 - 5 valid points (frequent in tracking, where the filter confidence gates key points)
    pnp_few_points_warp(V, X, uv, n, Kc, width, height, visible_thresh, opencv_return, o, sm, lane);
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
  if (n < 6) {
    pnp_few_points_warp(V, X, uv, n, Kc, width, height, visible_thresh, opencv_return, o, sm, lane);
    return;
  }
  double R[9], t[3];
  dlt_init_warp(X, uv, n, Kc[0], Kc[4], Kc[2], Kc[5], R, t, sm, lane);
  const double cost = refine_lm_warp(X, uv, n, Kc[0], Kc[4], Kc[2], Kc[5], R, t, sm, lane);
  pose::pnp_finish(V, R, t, cost, n, Kc, width, height, visible_thresh, opencv_return, o);
}

}  // namespace
}  // namespace cp

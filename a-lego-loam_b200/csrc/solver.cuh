// solver.cuh — the four analytic cost functions of include/alego/utility.h:122-349 and a block-wide
// Levenberg-Marquardt that follows what ceres::Solve does for the reference's settings (trust region,
// LEVENBERG_MARQUARDT, HuberLoss(0.1), jacobi scaling, one 6-parameter block; call sites
// laserOdometry.cpp:413-418,487-492 and laserMapping.cpp:468-475).
//
// One CTA solves one sequence's problem end to end: every pass over the residuals is a strided loop with a
// warp-shuffle + shared-memory reduction of the 6x6 normal equations (21 + 6 + 1 doubles); thread 0 runs the
// trust-region logic.  The reference's DENSE_QR factorises [J; D]; here the same minimiser is obtained from
// the normal equations (J^T J + D^2) s = J^T r by Cholesky — identical in exact arithmetic, ~1e-12 apart in
// double, far inside the 1e-4 pose tolerance.
#pragma once
#include "common.cuh"

// residual r and 6-vector Jacobian exactly as the reference's Evaluate() bodies fill them.
// kind: 0 CornerCostFunction, 1 SurfCostFunction, 2 LidarEdgeCostFunction, 3 LidarPlaneCostFunction
__device__ __forceinline__ void eval_resid_dev(int kind, const double *cpt, const double *a3, const double *b3, const double *c3,
                                               double dd, const double *x, const PoseTrig &T, double *r, double *J) {
  const double px = cpt[0], py = cpt[1], pz = cpt[2];
  const double lx = T.R[0] * px + T.R[1] * py + T.R[2] * pz + x[0];
  const double ly = T.R[3] * px + T.R[4] * py + T.R[5] * pz + x[1];
  const double lz = T.R[6] * px + T.R[7] * py + T.R[8] * pz + x[2];
  const double sr = T.sr, cr = T.cr, sp = T.sp, cp = T.cp, sy = T.sy, cy = T.cy;
  // d(lp)/d(roll,pitch,yaw), D[xyz][rpy] (utility.h:148-158).  D[1][1] keeps the reference's cr*sr*cp term
  // (utility.h:153,217,273,325; the true derivative has sy*cp*cr) — replicated on purpose.
  double D00 = 0, D10 = 0, D20 = 0, D01 = 0, D11 = 0, D21 = 0, D02 = 0, D12 = 0;
  if (J) {
    D00 = (cy * sp * cr + sr * sy) * py + (sy * cr - cy * sr * sp) * pz;
    D10 = (-cy * sr + sy * sp * cr) * py + (-sr * sy * sp - cy * cr) * pz;
    D20 = cp * cr * py - cp * sr * pz;
    D01 = -cy * sp * px + cy * cp * sr * py + cy * cr * cp * pz;
    D11 = -sp * sy * px + sy * cp * sr * py + cr * sr * cp * pz;
    D21 = -cp * px - sp * sr * py - sp * cr * pz;
    D02 = -sy * cp * px - (sy * sp * sr + cr * cy) * py + (cy * sr - sy * cr * sp) * pz;
    D12 = cp * cy * px + (-sy * cr + cy * sp * sr) * py + (cy * cr * sp + sy * sr) * pz;
  }
  if (kind == 0 || kind == 2) {
    const double *j = a3, *l = b3;
    const double e0 = j[0] - l[0], e1 = j[1] - l[1], e2 = j[2] - l[2];
    const double k = sqrt(e0 * e0 + e1 * e1 + e2 * e2);
    const double a = (ly - j[1]) * (lz - l[2]) - (lz - j[2]) * (ly - l[1]);
    const double b = (lz - j[2]) * (lx - l[0]) - (lx - j[0]) * (lz - l[2]);
    const double c = (lx - j[0]) * (ly - l[1]) - (ly - j[1]) * (lx - l[0]);
    const double m = sqrt(a * a + b * b + c * c);
    *r = m / k;
    if (J) {
      const double gx = (b * (l[2] - j[2]) + c * (j[1] - l[1])) / m;
      const double gy = (a * (j[2] - l[2]) - c * (j[0] - l[0])) / m;
      const double gz = (-a * (j[1] - l[1]) + b * (j[0] - l[0])) / m;
      if (kind == 0) {
        J[0] = gx / k; J[1] = gy / k; J[2] = 0.; J[3] = 0.; J[4] = 0.;
        J[5] = (gx * D02 + gy * D12 + gz * 0.) / k;
      } else {
        J[0] = gx / k; J[1] = gy / k; J[2] = gz / k;
        J[3] = (gx * D00 + gy * D10 + gz * D20) / k;
        J[4] = (gx * D01 + gy * D11 + gz * D21) / k;
        J[5] = (gx * D02 + gy * D12 + gz * 0.) / k;
      }
    }
  } else if (kind == 1) {
    const double *j = a3, *l = b3, *mm = c3;
    double a = (j[1] - l[1]) * (j[2] - mm[2]) - (j[2] - l[2]) * (j[1] - mm[1]);
    double b = (j[2] - l[2]) * (j[0] - mm[0]) - (j[0] - l[0]) * (j[2] - mm[2]);
    double c = (j[0] - l[0]) * (j[1] - mm[1]) - (j[1] - l[1]) * (j[0] - mm[0]);
    a *= a; b *= b; c *= c;
    const double ux = lx - j[0], uy = ly - j[1], uz = lz - j[2];
    const double m = sqrt(ux * ux * a + uy * uy * b + uz * uz * c);
    const double k = sqrt(a + b + c);
    *r = m / k;
    if (J) {
      const double gz = (uz * c) / (m * k);
      J[0] = 0.; J[1] = 0.; J[2] = gz / k; J[3] = 0.; J[4] = 0.; J[5] = 0.;
    }
  } else {
    *r = a3[0] * lx + a3[1] * ly + a3[2] * lz + dd;
    if (J) {
      J[0] = a3[0]; J[1] = a3[1]; J[2] = a3[2];
      J[3] = a3[0] * D00 + a3[1] * D10 + a3[2] * D20;
      J[4] = a3[0] * D01 + a3[1] * D11 + a3[2] * D21;
      J[5] = a3[0] * D02 + a3[1] * D12 + a3[2] * 0.;
    }
  }
}

// HuberLoss::Evaluate + Corrector (rho'' <= 0 branch): returns rho(s), *w = sqrt(rho'(s))
__device__ __forceinline__ double huber(double r, double a, double *w) {
  const double s = r * r, b = a * a;
  if (s > b) {
    const double rr = sqrt(s);
    *w = sqrt(fmax(2.2250738585072014e-308, a / rr));
    return 2.0 * a * rr - b;
  }
  *w = 1.0;
  return s;
}

#define LM_NACC 28  // 21 upper-triangular JtJ + 6 Jtr + cost

struct LmShared {
  double red[32][LM_NACC];
  double acc[LM_NACC];
  double x[6], xc[6];
  double cand_cost;
  int flag, flag2;  // flag2: the gradient test after an accepted step (its own word: flag is still being read then)
};

// block reduction of acc[LM_NACC] per thread into sh->acc
__device__ __forceinline__ void lm_block_reduce(double *v, int nacc, LmShared *sh) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int q = 0; q < nacc; ++q) v[q] = warp_sum(v[q]);
  __syncthreads();
  if (lane == 0)
    for (int q = 0; q < nacc; ++q) sh->red[wid][q] = v[q];
  __syncthreads();
  if (threadIdx.x < nacc) {
    double s = 0;
    for (int w = 0; w < nw; ++w) s += sh->red[w][threadIdx.x];
    sh->acc[threadIdx.x] = s;
  }
  __syncthreads();
}

// RS must provide:  __device__ int slots() const;  __device__ bool load(int i, int &kind, double cp[3], double a[3],
// double b[3], double c[3], double &d) const  (false = empty slot)
template <class RS>
__device__ void lm_eval_full(const RS &rs, const double *x, double huber_a, LmShared *sh) {
  const PoseTrig T(x);
  double acc[LM_NACC];
#pragma unroll
  for (int q = 0; q < LM_NACC; ++q) acc[q] = 0;
  const int n = rs.slots();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    int kind;
    double cpt[3], a[3], b[3], c[3], d;
    if (!rs.load(i, kind, cpt, a, b, c, d)) continue;
    double r, J[6], w;
    eval_resid_dev(kind, cpt, a, b, c, d, x, T, &r, J);
    const double rho = huber(r, huber_a, &w);
    r *= w;
#pragma unroll
    for (int q = 0; q < 6; ++q) J[q] *= w;
    int t = 0;
#pragma unroll
    for (int p = 0; p < 6; ++p)
#pragma unroll
      for (int q = p; q < 6; ++q) acc[t++] += J[p] * J[q];
#pragma unroll
    for (int q = 0; q < 6; ++q) acc[21 + q] += J[q] * r;
    acc[27] += 0.5 * rho;
  }
  lm_block_reduce(acc, LM_NACC, sh);
}

// in-place Cholesky solve of the symmetric positive definite 6x6 system A s = g; false if not SPD
__device__ __forceinline__ bool chol6_solve(double A[6][6], const double *g, double *s) {
  for (int i = 0; i < 6; ++i) {
    for (int j = 0; j <= i; ++j) {
      double sum = A[i][j];
      for (int k = 0; k < j; ++k) sum -= A[i][k] * A[j][k];
      if (i == j) {
        if (!(sum > 0.0)) return false;
        A[i][i] = sqrt(sum);
      } else {
        A[i][j] = sum / A[j][j];
      }
    }
  }
  double y[6];
  for (int i = 0; i < 6; ++i) {
    double sum = g[i];
    for (int k = 0; k < i; ++k) sum -= A[i][k] * y[k];
    y[i] = sum / A[i][i];
  }
  for (int i = 5; i >= 0; --i) {
    double sum = y[i];
    for (int k = i + 1; k < 6; ++k) sum -= A[k][i] * s[k];
    s[i] = sum / A[i][i];
  }
  for (int i = 0; i < 6; ++i)
    if (!isfinite(s[i])) return false;
  return true;
}

// Ordered compaction of the valid residual blocks of one sequence into shared memory: the trust-region loop sweeps the
// residuals once per iteration, and with one CTA per sequence each global-memory round trip of that sweep is exposed
// latency — staged blocks are read at shared-memory latency and the threads no longer idle on empty slots.
// valid(i) / copy(i, dst): slot i of n_slots; W words of type T per block.  smem: >= 33 ints.  Returns the number staged.
template <typename T, int W, class Valid, class Copy>
__device__ __forceinline__ int stage_blocks(int n_slots, T *dst, Valid valid, Copy copy, int *smem) {
  int run = 0;
  for (int c0 = 0; c0 < n_slots; c0 += blockDim.x) {
    const int i = c0 + threadIdx.x;
    const bool v = i < n_slots && valid(i);
    int total;
    const int ex = block_excl_scan(v ? 1 : 0, smem, &total);
    if (v) copy(i, dst + (size_t)(run + ex) * W);
    run += total;
  }
  __syncthreads();
  return run;
}

struct LmResult {
  int iterations;
  double initial_cost, final_cost;
};

#define LM_FLAG_STOP 0
#define LM_FLAG_CANDIDATE 1
#define LM_FLAG_INVALID 2
#define LM_FLAG_ACCEPT 3
#define LM_FLAG_REJECT 4

// x (6 doubles, shared or global) is updated in place.  All threads of the block must call.
template <class RS>
__device__ LmResult block_lm_solve(const RS &rs, double *x_io, int max_iters, double huber_a, LmShared *sh, double *trace,
                                   int *trace_n, int trace_cap) {
  LmResult res{0, 0.0, 0.0};
  if (threadIdx.x < 6) sh->x[threadIdx.x] = x_io[threadIdx.x];
  __syncthreads();
  lm_eval_full(rs, sh->x, huber_a, sh);
  // thread-0 private trust-region state
  double cost = sh->acc[27];
  double scale[6], diag[6], H[6][6], g[6];
  double radius = 1e4, decrease = 2.0, x_norm = 0.0, model_change = 0.0;
  bool reuse_diag = false;
  int iter = 0, invalid_run = 0;
  res.initial_cost = cost;
  auto unpack = [&]() {  // scaled normal equations from the block reduction
    int t = 0;
    for (int p = 0; p < 6; ++p)
      for (int q = p; q < 6; ++q) {
        const double v = sh->acc[t++] * scale[p] * scale[q];
        H[p][q] = v;
        H[q][p] = v;
      }
    for (int q = 0; q < 6; ++q) g[q] = sh->acc[21 + q] * scale[q];
  };
  auto push_trace = [&](double c) {
    if (trace && *trace_n < trace_cap) {
      double *dst = trace + (size_t)(*trace_n) * 7;
      dst[0] = c;
      for (int q = 0; q < 6; ++q) dst[1 + q] = sh->x[q];
      ++*trace_n;
    }
  };
  if (threadIdx.x == 0) {
    // jacobi scaling from the initial Jacobian: 1 / (1 + ||column||)
    const int dpos[6] = {0, 6, 11, 15, 18, 20};
    for (int q = 0; q < 6; ++q) scale[q] = 1.0 / (1.0 + sqrt(sh->acc[dpos[q]]));
    unpack();
    for (int q = 0; q < 6; ++q) x_norm += sh->x[q] * sh->x[q];
    x_norm = sqrt(x_norm);
    push_trace(cost);
  }
  while (true) {
    if (threadIdx.x == 0) {
      int flag = LM_FLAG_CANDIDATE;
      if (iter >= max_iters || radius <= 1e-32) {
        flag = LM_FLAG_STOP;
      } else {
        ++iter;
        if (!reuse_diag)
          for (int q = 0; q < 6; ++q) diag[q] = fmin(fmax(H[q][q], 1e-6), 1e32);
        double A[6][6], s[6];
        for (int p = 0; p < 6; ++p)
          for (int q = 0; q < 6; ++q) A[p][q] = H[p][q];
        for (int q = 0; q < 6; ++q) A[q][q] += diag[q] / radius;  // D^2 = diag / radius
        const bool ok = chol6_solve(A, g, s);
        reuse_diag = true;
        model_change = 0.0;
        if (ok) {
          double step[6];
          for (int q = 0; q < 6; ++q) step[q] = -s[q];
          // -(J step).(r + J step / 2) = -step.g - step^T H step / 2
          double sg = 0, shs = 0;
          for (int p = 0; p < 6; ++p) {
            sg += step[p] * g[p];
            double row = 0;
            for (int q = 0; q < 6; ++q) row += H[p][q] * step[q];
            shs += step[p] * row;
          }
          model_change = -sg - 0.5 * shs;
          for (int q = 0; q < 6; ++q) sh->xc[q] = sh->x[q] + step[q] * scale[q];
        }
        if (!ok || !(model_change > 0.0)) {
          if (++invalid_run >= 5) flag = LM_FLAG_STOP;
          else { radius *= 0.5; flag = LM_FLAG_INVALID; push_trace(cost); }
        } else {
          invalid_run = 0;
        }
      }
      sh->flag = flag;
    }
    __syncthreads();
    int flag = sh->flag;
    if (flag == LM_FLAG_STOP) break;
    if (flag == LM_FLAG_INVALID) { __syncthreads(); continue; }
    // Residuals AND Jacobian at the candidate in one pass: ceres evaluates the cost at the candidate and, when the step
    // is accepted, the Jacobian at the same point — accepted steps are the common case, so the second pass over the
    // residuals is saved (the cost of this pass is bit-identical to a cost-only pass: same sums in the same order).
    lm_eval_full(rs, sh->xc, huber_a, sh);
    const double cand = sh->acc[27];
    if (threadIdx.x == 0) {
      double step_norm = 0;
      for (int q = 0; q < 6; ++q) step_norm += (sh->x[q] - sh->xc[q]) * (sh->x[q] - sh->xc[q]);
      step_norm = sqrt(step_norm);
      const double cost_change = cost - cand;
      int f;
      if (step_norm <= 1e-8 * (x_norm + 1e-8)) f = LM_FLAG_STOP;            // parameter tolerance
      else if (fabs(cost_change) <= 1e-6 * cost) f = LM_FLAG_STOP;           // function tolerance
      else {
        const double rho = cost_change / model_change;
        if (rho > 1e-3) {
          for (int q = 0; q < 6; ++q) sh->x[q] = sh->xc[q];
          x_norm = 0;
          for (int q = 0; q < 6; ++q) x_norm += sh->x[q] * sh->x[q];
          x_norm = sqrt(x_norm);
          const double t = 2.0 * rho - 1.0;
          radius = radius / fmax(1.0 / 3.0, 1.0 - t * t * t);
          radius = fmin(1e16, radius);
          decrease = 2.0;
          reuse_diag = false;
          f = LM_FLAG_ACCEPT;
        } else {
          radius = radius / decrease;
          decrease *= 2.0;
          reuse_diag = true;
          f = LM_FLAG_REJECT;
          push_trace(cost);
        }
      }
      sh->flag = f;
    }
    __syncthreads();
    flag = sh->flag;
    if (flag == LM_FLAG_STOP) break;
    if (flag == LM_FLAG_ACCEPT) {
      if (threadIdx.x == 0) {  // sh->acc still holds the normal equations of the accepted point
        cost = sh->acc[27];
        unpack();
        push_trace(cost);
        double gmax = 0;
        for (int q = 0; q < 6; ++q) gmax = fmax(gmax, fabs(sh->acc[21 + q]));
        sh->flag2 = gmax <= 1e-10 ? LM_FLAG_STOP : LM_FLAG_CANDIDATE;
      }
      __syncthreads();
      if (sh->flag2 == LM_FLAG_STOP) break;
    }
    __syncthreads();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    sh->red[0][0] = cost;
    sh->red[0][1] = (double)iter;
    sh->red[0][2] = res.initial_cost;
  }
  __syncthreads();
  res.final_cost = sh->red[0][0];
  res.iterations = (int)sh->red[0][1];
  res.initial_cost = sh->red[0][2];
  if (threadIdx.x < 6) x_io[threadIdx.x] = sh->x[threadIdx.x];
  __syncthreads();
  return res;
}

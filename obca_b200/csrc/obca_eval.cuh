// obca_eval.cuh -- K1 stand-alone: fused evaluation of the parking NLP in the REFERENCE's formulation
// (per-stage timeScale, start/end and chain rows present), one thread per (problem, stage).
//
// For every problem it reads the stacked primal point z = (x, timeScale, u, l, n[, sl]) (n values), the row
// multipliers y (m values) and the tracking parameters rx, ry, ryaw (n_par), and writes
//     c      the m constraint rows of ParkingSignedDist.jl:122-208 (bodies of the inequality rows)
//     gradL  grad_z [ f(z) + y' c(z) ]     (n values, reference variable order)
//     fk     the objective split per stage (N+1 values)
// i.e. exactly the fused K1 of SURVEY.md section 8(d): 8 * (2n + 2m + n_par) bytes per problem per evaluation,
// Jacobian and Hessian never reach HBM (the solver consumes them in registers / shared memory).
//
// Row order of c and y (m = 8 + 6N + 4 nOb (N+1) for both variants):
//   start 4 | end 4 | dyn 4xN (stage-major) | chain N | rate N | norm nOb x(N+1) | rot 2nOb x(N+1) | dist nOb x(N+1)
// Variable order of gradL = JuMP declaration order (ParkingSignedDist.jl:49-59):
//   x 4x(N+1) | timeScale (N+1) | u 2xN | l Vx(N+1) | n 4nOb x(N+1) | sl nOb x(N+1)   (sl block only for SD)
#pragma once
#include "obca_common.cuh"
#include "obca_local.cuh"
#include "obca_stage.cuh"

namespace obca {

struct EvalIn {
  const double *x0, *xF, *rx, *ry, *ryaw;       // per problem
  const double *xp, *up, *ts, *lp, *np, *sl;    // reference output shapes (column-major)
  const double* y;                                // m row multipliers (may be null = zeros)
};
struct EvalOut {
  double *c, *gradL, *fk;
};

OBCA_HD size_t eval_m(const ParkProblem& P) { return 8 + (size_t)6 * P.N + (size_t)4 * P.nOb * (P.N + 1); }
OBCA_HD size_t eval_n(const ParkProblem& P) {
  const size_t NS = P.N + 1;
  return 4 * NS + NS + 2 * (size_t)P.N + (size_t)P.V * NS + (size_t)4 * P.nOb * NS + (P.signed_dist ? (size_t)P.nOb * NS : 0);
}

// all pointers already offset to problem b
template <int VM, bool SDV>
OBCA_HD void eval_stage(const ParkProblem& P, int k, const EvalIn& in, const EvalOut& out) {
  const int N = P.N, NS = N + 1, nOb = P.nOb, V = P.V;
  const bool fix = P.fix_time != 0;
  // row offsets
  const size_t oDyn = 8, oChain = oDyn + (size_t)4 * N, oRate = oChain + N, oNorm = oRate + N,
               oRot = oNorm + (size_t)nOb * NS, oDist = oRot + (size_t)2 * nOb * NS;
  // variable offsets
  const size_t vT = (size_t)4 * NS, vU = vT + NS, vL = vU + (size_t)2 * N, vN = vL + (size_t)V * NS, vS = vN + (size_t)4 * nOb * NS;
  const double* y = in.y;
#define YV(i) (y ? y[i] : 0.0)
  const double X = in.xp[4 * k], Y = in.xp[4 * k + 1], ps = in.xp[4 * k + 2], v = in.xp[4 * k + 3];
  const double tsk = fix ? 1.0 : in.ts[k];
  double sn_, cs_;
  sincos(ps, &sn_, &cs_);
  double fobj = 0.0;
  // ---- objective on the state + end-point rows ----
  const double ex = X - in.rx[k], ey = Y - in.ry[k], ep = ps - in.ryaw[k];
  fobj += 1e-4 * v * v + 1e-3 * ex * ex + 1e-3 * ey * ey + P.w_yaw * ep * ep;
  double gX = 2e-3 * ex, gY = 2e-3 * ey, gP = 2.0 * P.w_yaw * ep, gV = 2e-4 * v;
  if (k == 0) {
    const double xs[4] = {X, Y, ps, v};
#pragma unroll
    for (int i = 0; i < 4; ++i) out.c[i] = xs[i] - in.x0[i];
    gX += YV(0); gY += YV(1); gP += YV(2); gV += YV(3);
  }
  if (k == N) {
    const double xs[4] = {X, Y, ps, v};
#pragma unroll
    for (int i = 0; i < 4; ++i) out.c[4 + i] = xs[i] - in.xF[i];
    gX += YV(4); gY += YV(5); gP += YV(6); gV += YV(7);
  }
  if (k >= 1) {   // multiplier of the dynamics row that produced x_k
    gX += YV(oDyn + 4 * (k - 1) + 0); gY += YV(oDyn + 4 * (k - 1) + 1); gP += YV(oDyn + 4 * (k - 1) + 2); gV += YV(oDyn + 4 * (k - 1) + 3);
  }
  double gT = 0.0;
  if (!fix) {
    fobj += 0.5 * tsk + tsk * tsk;
    gT += 0.5 + 2.0 * tsk;
    if (k < N) gT += YV(oChain + k);
    if (k >= 1) gT -= YV(oChain + k - 1);
    if (k < N) out.c[oChain + k] = tsk - in.ts[k + 1];
  } else if (k < N) {
    out.c[oChain + k] = 0.0;
  }
  // ---- controls, input-rate terms, steering-rate row, dynamics ----
  if (k < N) {
    const double de = in.up[2 * k], ac = in.up[2 * k + 1];
    const double h = tsk * P.Ts, ih = 1.0 / h, ih2 = ih * ih;
    double gD = 0.02 * de, gA = 2.0 * P.w_a * ac;
    fobj += 0.01 * de * de + P.w_a * ac * ac;
    // (u[k+1]-u[k])/(ts[k] Ts) term, k <= N-2 (:87), and the u0 term for k == 0 (:88)
    if (k + 1 < N) {
      const double ed = in.up[2 * (k + 1)] - de, ea = in.up[2 * (k + 1) + 1] - ac;
      const double T = 0.1 * (ed * ed + ea * ea) * ih2;
      fobj += T;
      gD += -0.2 * ed * ih2; gA += -0.2 * ea * ih2;
      if (!fix) gT += -2.0 * T / tsk;
    }
    if (k == 0) {
      const double T = 0.1 * (de * de + ac * ac) * ih2;
      fobj += T;
      gD += 0.2 * de * ih2; gA += 0.2 * ac * ih2;
      if (!fix) gT += -2.0 * T / tsk;
    }
    if (k >= 1) {   // the term of stage k-1 in which u[k] is the later control (uses ts[k-1])
      const double hp = (fix ? 1.0 : in.ts[k - 1]) * P.Ts, ihp2 = 1.0 / (hp * hp);
      gD += 0.2 * (de - in.up[2 * (k - 1)]) * ihp2; gA += 0.2 * (ac - in.up[2 * (k - 1) + 1]) * ihp2;
    }
    // rate row k: (u[1,k-1] - u[1,k]) / (ts[k] Ts)   (:167-173)
    const double wd = k > 0 ? in.up[2 * (k - 1)] : 0.0;
    const double r = (wd - de) * ih;
    out.c[oRate + k] = r;
    const double yr = YV(oRate + k);
    gD += -yr * ih;
    if (!fix) gT += -yr * r / tsk;
    if (k + 1 < N) {   // rate row k+1 sees de_k as the previous control
      const double hn = (fix ? 1.0 : in.ts[k + 1]) * P.Ts;
      gD += YV(oRate + k + 1) / hn;
    }
    // dynamics row k
    DynOut dyn;
    dyn_eval(P, X, Y, ps, v, de, ac, tsk, dyn, nullptr, nullptr);
    double pi[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      pi[i] = YV(oDyn + 4 * k + i);
      out.c[oDyn + 4 * k + i] = in.xp[4 * (k + 1) + i] - dyn.f[i];
    }
    gX -= pi[0]; gY -= pi[1];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      gP -= pi[i] * dyn.fx[i][0]; gV -= pi[i] * dyn.fx[i][1];
      gD -= pi[i] * dyn.fu[i][0]; gA -= pi[i] * dyn.fu[i][1];
      if (!fix) gT -= pi[i] * dyn.ft[i];
    }
    out.gradL[vU + 2 * k] = gD; out.gradL[vU + 2 * k + 1] = gA;
  }
  // ---- OBCA blocks ----
  for (int j = 0; j < nOb; ++j) {
    ObsRows<VM> R; ObsVars<VM> Q; ObsGeom<VM> G;
    R.v = P.vOb[j];
#pragma unroll
    for (int i = 0; i < VM; ++i) {
      const bool on = i < R.v;
      const int rr = P.voff[j] + (on ? i : 0);
      R.a1[i] = on ? P.A[rr][0] : 0.0; R.a2[i] = on ? P.A[rr][1] : 0.0; R.bb[i] = on ? P.b[rr] : 0.0;
      Q.lam[i] = on ? in.lp[(size_t)V * k + P.voff[j] + i] : 0.0;
      Q.zlam[i] = 0.0;
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) { Q.mu[m] = in.np[(size_t)4 * nOb * k + 4 * j + m]; Q.zmu[m] = 0.0; }
    Q.sl = SDV ? in.sl[(size_t)nOb * k + j] : 0.0;
    Q.sd = 0.0; Q.sn = 0.0;
    Q.yr1 = YV(oRot + (size_t)2 * nOb * k + 2 * j); Q.yr2 = YV(oRot + (size_t)2 * nOb * k + 2 * j + 1);
    const double ynorm = YV(oNorm + (size_t)nOb * k + j), ydist = YV(oDist + (size_t)nOb * k + j);
    Q.yn = ynorm; Q.vn = ynorm; Q.vd = -ydist;      // obs_lagr_grad uses yI_dist = -vd and y_norm = yn (SD) / vn (Dist)
    obs_geom<VM, SDV>(P, X, Y, cs_, sn_, R, Q, G);
    out.c[oNorm + (size_t)nOb * k + j] = SDV ? G.cn : G.pp;       // norm row body (== 1 residual for SD, <= 1 body for Dist)
    out.c[oRot + (size_t)2 * nOb * k + 2 * j] = G.cr1;
    out.c[oRot + (size_t)2 * nOb * k + 2 * j + 1] = G.cr2;
    out.c[oDist + (size_t)nOb * k + j] = G.gd;                    // >= dmin
    double rl[VM], rm[4], rs, gx[3];
    obs_lagr_grad<VM, SDV>(P, R, Q, G, rl, rm, rs, gx);
    gX += gx[0]; gY += gx[1]; gP += gx[2];
#pragma unroll
    for (int i = 0; i < VM; ++i)
      if (i < R.v) out.gradL[vL + (size_t)V * k + P.voff[j] + i] = rl[i];
#pragma unroll
    for (int m = 0; m < 4; ++m) out.gradL[vN + (size_t)4 * nOb * k + 4 * j + m] = rm[m];
    if (SDV) {
      out.gradL[vS + (size_t)nOb * k + j] = rs;
      fobj += 1e2 * Q.sl + 1e4 * Q.sl * Q.sl;
    }
  }
  out.gradL[4 * k] = gX; out.gradL[4 * k + 1] = gY; out.gradL[4 * k + 2] = gP; out.gradL[4 * k + 3] = gV;
  out.gradL[vT + k] = gT;
  out.fk[k] = fobj;
#undef YV
}

}  // namespace obca

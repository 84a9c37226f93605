// obca_stage.cuh -- per-stage pieces of the parking NLP that are not obstacle blocks:
//   * bicycle dynamics of ParkingSignedDist.jl:142-150 with first and second derivatives
//   * input-rate objective terms (:79-80, :87-88) and the steering-rate rows (:157-174)
//   * the Riccati stage step of the stage-banded KKT solve.
//
// Stage vector used by the KKT solve (DESIGN.md "KKT solve"):
//     y = [ X, Y, psi, v, wd, wa, t | de, a ]      (state s: 7, control u: 2)
// wd, wa = copy of the previous control (so that the input-rate terms are stage-local),
// t      = the time-scale variable (timeScale[i] == timeScale[i+1], :153, carried as a constant state).
#pragma once
#include "obca_common.cuh"

namespace obca {

constexpr int NSV = 7;                 // state entries of the stage vector
constexpr int NYV = 9;                 // state + control
constexpr int NQ = NYV * (NYV + 1) / 2;  // packed symmetric 9x9
constexpr int IX = 0, IY = 1, IP = 2, IV = 3, IWD = 4, IWA = 5, IT = 6, IDE = 7, IAC = 8;

struct DynOut {
  double f[4];        // f(x,u,t)
  double fx[4][2];    // d f_i / d(psi, v)      (d/dX, d/dY are identity)
  double ft[4];       // d f_i / dt
  double fu[4][2];    // d f_i / d(de, a)
};

// Dynamics (ParkingSignedDist.jl:147-150; fixed-time :142-145 when fix != 0) and, if pi != nullptr, the packed
// 5x5 Hessian H5 (order psi, v, t, de, a; upper triangle, 15 entries) of  -sum_i pi_i f_i.
OBCA_HD void dyn_eval(const ParkProblem& P, double X, double Y, double psi, double v, double de, double a, double t,
                      DynOut& o, const double* pi, double* H5) {
  const double Ts = P.Ts;
  const bool fix = P.fix_time != 0;
  const double h = fix ? Ts : t * Ts;
  const double tde = tan(de);
  const double k0 = tde / P.L;
  const double k1 = (1.0 + tde * tde) / P.L;
  const double k2 = 2.0 * tde * k1;
  const double vm = v + 0.5 * h * a;
  const double Dd = h * vm;
  const double th = psi + 0.5 * h * v * k0;
  double S, C;
  sincos(th, &S, &C);
  o.f[0] = X + Dd * C;
  o.f[1] = Y + Dd * S;
  o.f[2] = psi + Dd * k0;
  o.f[3] = v + h * a;
  // first derivatives, q = (psi, v, t, de, a)
  const double Dt = fix ? 0.0 : Ts * (v + h * a);
  const double Dq[5] = {0.0, h, Dt, 0.0, 0.5 * h * h};
  const double Tq[5] = {1.0, 0.5 * h * k0, fix ? 0.0 : 0.5 * Ts * v * k0, 0.5 * h * v * k1, 0.0};
  double g0[5], g1[5], g2[5], g3[5];
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    g0[q] = Dq[q] * C - Dd * S * Tq[q];
    g1[q] = Dq[q] * S + Dd * C * Tq[q];
    g2[q] = Dq[q] * k0;
    g3[q] = 0.0;
  }
  g2[0] += 1.0; g2[3] += Dd * k1;
  g3[1] = 1.0; g3[4] = h; g3[2] = fix ? 0.0 : Ts * a;
  o.fx[0][0] = g0[0]; o.fx[0][1] = g0[1]; o.ft[0] = g0[2]; o.fu[0][0] = g0[3]; o.fu[0][1] = g0[4];
  o.fx[1][0] = g1[0]; o.fx[1][1] = g1[1]; o.ft[1] = g1[2]; o.fu[1][0] = g1[3]; o.fu[1][1] = g1[4];
  o.fx[2][0] = g2[0]; o.fx[2][1] = g2[1]; o.ft[2] = g2[2]; o.fu[2][0] = g2[3]; o.fu[2][1] = g2[4];
  o.fx[3][0] = g3[0]; o.fx[3][1] = g3[1]; o.ft[3] = g3[2]; o.fu[3][0] = g3[3]; o.fu[3][1] = g3[4];
  if (pi == nullptr) return;
  // second derivatives
  double Dqq[5][5], Tqq[5][5];
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int j = 0; j < 5; ++j) { Dqq[i][j] = 0.0; Tqq[i][j] = 0.0; }
  if (!fix) {
    Dqq[1][2] = Dqq[2][1] = Ts;
    Dqq[2][4] = Dqq[4][2] = Ts * h;
    Dqq[2][2] = Ts * Ts * a;
    Tqq[1][2] = Tqq[2][1] = 0.5 * Ts * k0;
    Tqq[2][3] = Tqq[3][2] = 0.5 * Ts * v * k1;
  }
  Tqq[1][3] = Tqq[3][1] = 0.5 * h * k1;
  Tqq[3][3] = 0.5 * h * v * k2;
  const double Kq[5] = {0.0, 0.0, 0.0, k1, 0.0};
  int e = 0;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
#pragma unroll
    for (int j = i; j < 5; ++j) {
      const double cross = Dq[i] * Tq[j] + Dq[j] * Tq[i];
      const double tt = Tq[i] * Tq[j];
      const double h0 = Dqq[i][j] * C - S * cross - Dd * C * tt - Dd * S * Tqq[i][j];
      const double h1 = Dqq[i][j] * S + C * cross - Dd * S * tt + Dd * C * Tqq[i][j];
      double h2 = Dqq[i][j] * k0 + Dq[i] * Kq[j] + Dq[j] * Kq[i];
      if (i == 3 && j == 3) h2 += Dd * k2;
      double h3 = 0.0;
      if (!fix && i == 2 && j == 4) h3 = Ts;
      H5[e++] = -(pi[0] * h0 + pi[1] * h1 + pi[2] * h2 + pi[3] * h3);
    }
  }
}

// map (psi, v, t, de, a) -> stage-vector index
OBCA_HD constexpr int q5_to_y(int q) { return q == 0 ? IP : q == 1 ? IV : q == 2 ? IT : q == 3 ? IDE : IAC; }

// ------------------------------------------------------------------------------------------------------------
// Riccati stage step.   Value function of stage k+1:  V+(s) = 1/2 s' Pn s + pn' s   (s = 7-vector).
// Stage model:  1/2 y' Q y + q' y  with y = (s, u),  s+ = Phi y + rt,  rt = (r0..r3, 0, 0, 0).
// Outputs: P, p of stage k; gain K (2x7), feed-forward kf (2).  Returns 0 if Huu is not positive definite.
// ------------------------------------------------------------------------------------------------------------
struct RicStage {
  double K[2][NSV];
  double kf[2];
};

OBCA_HD int riccati_step(const DynOut& d, const double* r4, const double* Q, const double* q, const double* Pn,
                         const double* pn, double* Pk, double* pk, RicStage& G) {
  // Phi (7x9) sparse rows
  double Phi[NSV][NYV];
#pragma unroll
  for (int i = 0; i < NSV; ++i)
#pragma unroll
    for (int j = 0; j < NYV; ++j) Phi[i][j] = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    Phi[i][IP] = d.fx[i][0]; Phi[i][IV] = d.fx[i][1]; Phi[i][IT] = d.ft[i];
    Phi[i][IDE] = d.fu[i][0]; Phi[i][IAC] = d.fu[i][1];
  }
  Phi[0][IX] = 1.0; Phi[1][IY] = 1.0;
  Phi[4][IDE] = 1.0; Phi[5][IAC] = 1.0; Phi[6][IT] = 1.0;
  // T = Pn * Phi (7x9), g = pn + Pn * rt
  double T[NSV][NYV];
  double gv[NSV];
#pragma unroll
  for (int i = 0; i < NSV; ++i) {
    double acc = pn[i];
#pragma unroll
    for (int l = 0; l < 4; ++l) acc += Pn[sym_idx_any<NSV>(i, l)] * r4[l];
    gv[i] = acc;
#pragma unroll
    for (int j = 0; j < NYV; ++j) {
      double t = 0.0;
#pragma unroll
      for (int l = 0; l < NSV; ++l) t += Pn[sym_idx_any<NSV>(i, l)] * Phi[l][j];
      T[i][j] = t;
    }
  }
  // H = Q + Phi' T ; hv = q + Phi' g
  double H[NQ], hv[NYV];
#pragma unroll
  for (int i = 0; i < NYV; ++i) {
    double acc = q[i];
#pragma unroll
    for (int l = 0; l < NSV; ++l) acc += Phi[l][i] * gv[l];
    hv[i] = acc;
#pragma unroll
    for (int j = i; j < NYV; ++j) {
      double t = Q[sym_idx<NYV>(i, j)];
#pragma unroll
      for (int l = 0; l < NSV; ++l) t += Phi[l][i] * T[l][j];
      H[sym_idx<NYV>(i, j)] = t;
    }
  }
  // eliminate u = (de, a): 2x2 Cholesky-type
  int ok = 1;
  double h77 = H[sym_idx<NYV>(IDE, IDE)], h78 = H[sym_idx<NYV>(IDE, IAC)], h88 = H[sym_idx<NYV>(IAC, IAC)];
  if (!(h77 > 0.0)) { ok = 0; h77 = 1e300; }
  const double i77 = 1.0 / h77;
  double s88 = h88 - h78 * h78 * i77;
  if (!(s88 > 0.0)) { ok = 0; s88 = 1e300; }
  const double i88 = 1.0 / s88;
  // inverse of Huu
  const double n00 = i77 + h78 * h78 * i77 * i77 * i88, n01 = -h78 * i77 * i88, n11 = i88;
#pragma unroll
  for (int j = 0; j < NSV; ++j) {
    const double a = H[sym_idx<NYV>(j, IDE)], b = H[sym_idx<NYV>(j, IAC)];
    G.K[0][j] = -(n00 * a + n01 * b);
    G.K[1][j] = -(n01 * a + n11 * b);
  }
  G.kf[0] = -(n00 * hv[IDE] + n01 * hv[IAC]);
  G.kf[1] = -(n01 * hv[IDE] + n11 * hv[IAC]);
#pragma unroll
  for (int i = 0; i < NSV; ++i) {
    const double a = H[sym_idx<NYV>(i, IDE)], b = H[sym_idx<NYV>(i, IAC)];
    pk[i] = hv[i] + a * G.kf[0] + b * G.kf[1];
#pragma unroll
    for (int j = i; j < NSV; ++j)
      Pk[sym_idx<NSV>(i, j)] = H[sym_idx<NYV>(i, j)] + a * G.K[0][j] + b * G.K[1][j];
  }
  return ok;
}

}  // namespace obca

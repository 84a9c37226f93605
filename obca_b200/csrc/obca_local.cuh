// obca_local.cuh -- the OBCA constraint block of one (stage k, obstacle j) pair.
//
// Reference rows (AutonomousParking/ParkingSignedDist.jl:190-207, ParkingDist.jl:192-208), p = A_j' * lambda:
//   norm : p1^2 + p2^2 == 1 (SD, :198)            /  <= 1 (Dist, ParkingDist.jl:200)
//   rot1 : mu1 - mu3 + cos(psi) p1 + sin(psi) p2 == 0                                   (:201)
//   rot2 : mu2 - mu4 - sin(psi) p1 + cos(psi) p2 == 0                                   (:202)
//   dist : -g'mu + (X + cos(psi) off) p1 + (Y + sin(psi) off) p2 - b'lambda [+ sl] >= dmin   (:205-206)
// plus lambda >= 0, mu >= 0 (:114-115) and the slack cost 1e2 sl + 1e4 sl^2 (:92).
//
// This file provides, for one block:
//   * residuals / Lagrangian-gradient pieces (KKT error, merit function)
//   * the local Newton block: every local unknown (lambda, mu, sl and the multipliers of norm/rot) is eliminated
//     in registers, leaving a 3x3 Schur complement + 3-vector on the pose (X, Y, psi) of the stage
//     ("condensation"), and the triangular factor needed to recover the local step afterwards.
// Elimination order (static, inertia-revealing -- see DESIGN.md "KKT solve"):
//   1. dist slack (and, Dist variant, norm slack):   quasi-definite pivots  -> +Sigma * G G'
//   2. (mu3, y_rot1), (mu4, y_rot2):                 2x2 pivots with off-diagonal -1, inertia (1,1) each
//   3. SD only: (lambda_piv, y_norm) 2x2 pivot, lambda_piv = argmax |d norm / d lambda_i|, inertia (1,1)
//   4. remaining lambda, mu1, mu2, sl:               1x1 pivots that must be > 0 for the correct KKT inertia
#pragma once
#include "obca_common.cuh"

namespace obca {

template <int VM, bool SDV>
struct LocalDims {
  static constexpr int YN = SDV ? 1 : 0;
  static constexpr int SLN = SDV ? 1 : 0;
  static constexpr int I_MU1 = VM + YN;
  static constexpr int I_MU2 = VM + YN + 1;
  static constexpr int I_SL = VM + YN + 2;
  static constexpr int NLT = VM + YN + 2 + SLN;   // eliminated unknowns
  static constexpr int I_X = NLT, I_Y = NLT + 1, I_P = NLT + 2;
  static constexpr int ND = NLT + 3;
  static constexpr int NM = ND * (ND + 1) / 2;
  static constexpr int NFAC = 4 * NLT;            // local step as an affine map of the pose step: a + B (dX, dY, dpsi)
  OBCA_HD static constexpr int il(int i) { return i == 0 ? 0 : i + YN; }   // position of lambda_i
};

template <int VM>
struct ObsRows {
  double a1[VM], a2[VM], bb[VM];
  int v;
};

template <int VM>
struct ObsVars {
  double lam[VM], zlam[VM];
  double mu[4], zmu[4];
  double sl;
  double yn, yr1, yr2;
  double sd, vd;   // dist slack (>= dmin) and its bound multiplier
  double sn, vn;   // Dist variant: norm slack (<= 1) and its bound multiplier
};

template <int VM>
struct ObsGeom {
  double p1, p2, e1, e2;
  double ah1[VM], ah2[VM], rho[VM], gn[VM];
  double gd;           // value of the dist expression (incl. +sl for SD)
  double pp;           // p1^2 + p2^2
  double cn, cr1, cr2, cd;   // residuals: norm (eq or g - s), rot1, rot2, dist (g - s)
  int piv;
};

template <int VM>
OBCA_HD void swap_rows(ObsRows<VM>& R, ObsVars<VM>& Q, int piv) {
#pragma unroll
  for (int i = 1; i < VM; ++i) {
    if (i == piv) {
      double t;
      t = R.a1[0]; R.a1[0] = R.a1[i]; R.a1[i] = t;
      t = R.a2[0]; R.a2[0] = R.a2[i]; R.a2[i] = t;
      t = R.bb[0]; R.bb[0] = R.bb[i]; R.bb[i] = t;
      t = Q.lam[0]; Q.lam[0] = Q.lam[i]; Q.lam[i] = t;
      t = Q.zlam[0]; Q.zlam[0] = Q.zlam[i]; Q.zlam[i] = t;
    }
  }
}

// geometry + residuals.  (c, s) = (cos psi, sin psi).
template <int VM, bool SDV>
OBCA_HD void obs_geom(const ParkProblem& P, double X, double Y, double c, double s, const ObsRows<VM>& R,
                      const ObsVars<VM>& Q, ObsGeom<VM>& G) {
  double p1 = 0.0, p2 = 0.0, bl = 0.0;
#pragma unroll
  for (int i = 0; i < VM; ++i) {
    if (i < R.v) { p1 += R.a1[i] * Q.lam[i]; p2 += R.a2[i] * Q.lam[i]; bl += R.bb[i] * Q.lam[i]; }
  }
  const double tcx = X + c * P.off, tcy = Y + s * P.off;
  G.p1 = p1; G.p2 = p2;
  G.e1 = c * p1 + s * p2;
  G.e2 = -s * p1 + c * p2;
#pragma unroll
  for (int i = 0; i < VM; ++i) {
    G.ah1[i] = c * R.a1[i] + s * R.a2[i];
    G.ah2[i] = -s * R.a1[i] + c * R.a2[i];
    G.rho[i] = R.a1[i] * tcx + R.a2[i] * tcy - R.bb[i];
    G.gn[i] = 2.0 * (R.a1[i] * p1 + R.a2[i] * p2);
  }
  G.gd = -(P.g[0] * Q.mu[0] + P.g[1] * Q.mu[1] + P.g[2] * Q.mu[2] + P.g[3] * Q.mu[3]) + tcx * p1 + tcy * p2 - bl +
         (SDV ? Q.sl : 0.0);
  G.pp = p1 * p1 + p2 * p2;
  G.cn = SDV ? (G.pp - 1.0) : (G.pp - Q.sn);
  G.cr1 = Q.mu[0] - Q.mu[2] + G.e1;
  G.cr2 = Q.mu[1] - Q.mu[3] + G.e2;
  G.cd = G.gd - Q.sd;
  G.piv = 0;
}

template <int VM, bool SDV>
OBCA_HD int choose_pivot(const ObsRows<VM>& R, const ObsGeom<VM>& G) {
  int piv = 0;
  if (SDV) {
    double best = dabs(G.gn[0]);
#pragma unroll
    for (int i = 1; i < VM; ++i) {
      if (i < R.v && dabs(G.gn[i]) > best) { best = dabs(G.gn[i]); piv = i; }
    }
  }
  return piv;
}

// ------------------------------------------------------------------------------------------------------------
// Local Newton block: build, eliminate, emit Schur complement on (X, Y, psi).
//   Sxx[6] = {XX, XY, XP, YY, YP, PP},  rx[3]: to be ADDED to the stage Hessian / gradient ("M d = -r" convention)
//   fac[NFAC]: factor for obs_recover.   returns 1 if all pivots have the sign required for inertia (n, m, 0).
// R and Q must already be permuted (swap_rows) with the pivot returned by choose_pivot.
// ------------------------------------------------------------------------------------------------------------
template <int VM, bool SDV>
OBCA_HD int obs_condense(const ParkProblem& P, const ObsRows<VM>& R, const ObsVars<VM>& Q, const ObsGeom<VM>& G,
                         double mu_b, double dw, double dc, double* Sxx, double* rx, double* fac, int fs) {
  typedef LocalDims<VM, SDV> D;
  constexpr int ND = D::ND;
  double M[D::NM];
  double r[ND];
#pragma unroll
  for (int i = 0; i < D::NM; ++i) M[i] = 0.0;
#pragma unroll
  for (int i = 0; i < ND; ++i) r[i] = 0.0;

  const double yId = -Q.vd;
  const double g1 = P.g[0], g2 = P.g[1], g3 = P.g[2], g4 = P.g[3];
  const double im3 = rcp(Q.mu[2]), im4 = rcp(Q.mu[3]);
  const double s3 = Q.zmu[2] * im3 + dw, s4 = Q.zmu[3] * im4 + dw;
  const double igd = rcp(Q.sd - P.dmin);
  const double Sd = Q.vd * igd;
  const double yd0 = -mu_b * igd + Sd * (G.cd - g3 * G.cr1 - g4 * G.cr2);
  const double c3 = s3 * G.cr1 - mu_b * im3;
  const double c4 = s4 * G.cr2 - mu_b * im4;

  // Rank-one terms  s3 t3 t3' + s4 t4 t4' + Sd Gt Gt'  and the gradient  t3 c3 + t4 c4 + Gt yd0,  written out by
  // sparsity pattern (t3, t4: rot rows after eliminating mu3, mu4; Gt: dist row after eliminating its slack):
  //   t3 = { lambda_i: ah1_i, mu1: 1,        psi: e2 }
  //   t4 = { lambda_i: ah2_i, mu2: 1,        psi: -e1 }
  //   Gt = { lambda_i: gt_i,  mu1: -g1-g3,   mu2: -g2-g4,  sl: 1 (SD),  X: p1,  Y: p2,  psi: gP }
  // (structural zeros are not multiplied: fewer FP64 operations and no registers for them)
  const double gM1 = -g1 - g3, gM2 = -g2 - g4, gP = P.off * G.e2 - g3 * G.e2 + g4 * G.e1;
  const double sM1 = Sd * gM1, sM2 = Sd * gM2, sP = Sd * gP, sX = Sd * G.p1, sY = Sd * G.p2;
  const double s3e = s3 * G.e2, s4e = -s4 * G.e1;
#pragma unroll
  for (int i = 0; i < VM; ++i) {
    if (i < R.v) {
      const int li = D::il(i);
      const double gti = G.rho[i] - g3 * G.ah1[i] - g4 * G.ah2[i];
      const double a3 = s3 * G.ah1[i], a4 = s4 * G.ah2[i], sg = Sd * gti;
#pragma unroll
      for (int l = i; l < VM; ++l) {
        if (l < R.v) {
          const double gtl = G.rho[l] - g3 * G.ah1[l] - g4 * G.ah2[l];
          M[sym_idx<ND>(li, D::il(l))] = a3 * G.ah1[l] + a4 * G.ah2[l] + sg * gtl;
        }
      }
      M[sym_idx<ND>(li, D::I_MU1)] = a3 + sg * gM1;
      M[sym_idx<ND>(li, D::I_MU2)] = a4 + sg * gM2;
      if (SDV) M[sym_idx<ND>(li, D::I_SL)] = sg;
      M[sym_idx<ND>(li, D::I_X)] = sg * G.p1;
      M[sym_idx<ND>(li, D::I_Y)] = sg * G.p2;
      M[sym_idx<ND>(li, D::I_P)] = a3 * G.e2 - a4 * G.e1 + sg * gP;
      r[li] = G.ah1[i] * c3 + G.ah2[i] * c4 + gti * yd0;
    }
  }
  M[sym_idx<ND>(D::I_MU1, D::I_MU1)] = s3 + sM1 * gM1;
  M[sym_idx<ND>(D::I_MU1, D::I_MU2)] = sM1 * gM2;
  M[sym_idx<ND>(D::I_MU2, D::I_MU2)] = s4 + sM2 * gM2;
  if (SDV) {
    M[sym_idx<ND>(D::I_MU1, D::I_SL)] = sM1;
    M[sym_idx<ND>(D::I_MU2, D::I_SL)] = sM2;
    M[sym_idx<ND>(D::I_SL, D::I_SL)] = Sd;
    M[sym_idx<ND>(D::I_SL, D::I_X)] = sX; M[sym_idx<ND>(D::I_SL, D::I_Y)] = sY; M[sym_idx<ND>(D::I_SL, D::I_P)] = sP;
    r[D::I_SL] = yd0;
  }
  M[sym_idx<ND>(D::I_MU1, D::I_X)] = sM1 * G.p1; M[sym_idx<ND>(D::I_MU1, D::I_Y)] = sM1 * G.p2;
  M[sym_idx<ND>(D::I_MU1, D::I_P)] = s3e + sM1 * gP;
  M[sym_idx<ND>(D::I_MU2, D::I_X)] = sM2 * G.p1; M[sym_idx<ND>(D::I_MU2, D::I_Y)] = sM2 * G.p2;
  M[sym_idx<ND>(D::I_MU2, D::I_P)] = s4e + sM2 * gP;
  M[sym_idx<ND>(D::I_X, D::I_X)] = sX * G.p1; M[sym_idx<ND>(D::I_X, D::I_Y)] = sX * G.p2; M[sym_idx<ND>(D::I_X, D::I_P)] = sX * gP;
  M[sym_idx<ND>(D::I_Y, D::I_Y)] = sY * G.p2; M[sym_idx<ND>(D::I_Y, D::I_P)] = sY * gP;
  M[sym_idx<ND>(D::I_P, D::I_P)] = s3e * G.e2 - s4e * G.e1 + sP * gP;
  r[D::I_MU1] = c3 + gM1 * yd0;
  r[D::I_MU2] = c4 + gM2 * yd0;
  r[D::I_X] = G.p1 * yd0; r[D::I_Y] = G.p2 * yd0;
  r[D::I_P] = G.e2 * c3 - G.e1 * c4 + gP * yd0;
  // Lagrangian Hessian + barrier diagonal of lambda
  const double ynorm = SDV ? Q.yn : Q.vn;   // multiplier of the norm row (row multiplier yI = vU for the Dist variant)
  double Sn = 0.0, yn0 = 0.0;
  if (!SDV) {
    const double ign = rcp(1.0 - Q.sn);
    Sn = Q.vn * ign;
    yn0 = mu_b * ign + Sn * G.cn;
  }
#pragma unroll
  for (int i = 0; i < VM; ++i) {
    const int li = D::il(i);
    if (i < R.v) {
#pragma unroll
      for (int l = i; l < VM; ++l) {
        if (l < R.v) {
          double w = 2.0 * ynorm * (R.a1[i] * R.a1[l] + R.a2[i] * R.a2[l]);
          if (!SDV) w += Sn * G.gn[i] * G.gn[l];
          M[sym_idx_any<ND>(li, D::il(l))] += w;
        }
      }
      const double il_ = rcp(Q.lam[i]);
      M[sym_idx<ND>(li, li)] += Q.zlam[i] * il_ + dw;
      M[sym_idx_any<ND>(li, D::I_X)] += yId * R.a1[i];
      M[sym_idx_any<ND>(li, D::I_Y)] += yId * R.a2[i];
      M[sym_idx_any<ND>(li, D::I_P)] += Q.yr1 * G.ah2[i] - Q.yr2 * G.ah1[i] + yId * P.off * G.ah2[i];
      r[li] += -mu_b * il_ + (SDV ? 0.0 : G.gn[i] * yn0);
      if (SDV) M[sym_idx_any<ND>(li, 1)] = G.gn[i];
    } else {
      // padded half-space: decoupled unit pivot
#pragma unroll
      for (int j = 0; j < ND; ++j) M[sym_idx_any<ND>(li, j)] = 0.0;
      M[sym_idx<ND>(li, li)] = 1.0;
      r[li] = 0.0;
    }
  }
  M[sym_idx<ND>(D::I_P, D::I_P)] += -Q.yr1 * G.e1 - Q.yr2 * G.e2 - yId * P.off * G.e1;
  const double im1 = rcp(Q.mu[0]), im2 = rcp(Q.mu[1]);
  M[sym_idx<ND>(D::I_MU1, D::I_MU1)] += Q.zmu[0] * im1 + dw;
  M[sym_idx<ND>(D::I_MU2, D::I_MU2)] += Q.zmu[1] * im2 + dw;
  r[D::I_MU1] += -mu_b * im1;
  r[D::I_MU2] += -mu_b * im2;
  if (SDV) {
    M[sym_idx<ND>(D::I_SL, D::I_SL)] += 2.0e4 + dw;
    r[D::I_SL] += 1.0e2 + 2.0e4 * Q.sl;
    M[sym_idx<ND>(1, 1)] = -dc;
    r[1] = G.cn + dc * Q.yn;
  }

  int ok = 1;
  // ---- elimination ----
  int first = 0;
  if (SDV) {
    // 2x2 pivot on (lambda_0, y_norm)
    const double m00 = M[sym_idx<ND>(0, 0)], m01 = M[sym_idx<ND>(0, 1)], m11 = M[sym_idx<ND>(1, 1)];
    double det = m00 * m11 - m01 * m01;
    if (!(det < 0.0)) { ok = 0; det = -1e-300; }
    const double id = rcp(det);
    const double i00 = m11 * id, i01 = -m01 * id, i11 = m00 * id;
    double u0[ND], u1[ND];
#pragma unroll
    for (int c_ = 2; c_ < ND; ++c_) {
      const double a = M[sym_idx<ND>(0, c_)], b = M[sym_idx<ND>(1, c_)];
      u0[c_] = i00 * a + i01 * b;
      u1[c_] = i01 * a + i11 * b;
    }
#pragma unroll
    for (int rr = 2; rr < ND; ++rr) {
#pragma unroll
      for (int c_ = rr; c_ < ND; ++c_)
        M[sym_idx<ND>(rr, c_)] -= u0[rr] * M[sym_idx<ND>(0, c_)] + u1[rr] * M[sym_idx<ND>(1, c_)];
      r[rr] -= u0[rr] * r[0] + u1[rr] * r[1];
    }
    // store inverse of the 2x2 block in place of it
    M[sym_idx<ND>(0, 0)] = i00; M[sym_idx<ND>(0, 1)] = i01; M[sym_idx<ND>(1, 1)] = i11;
    first = 2;
  }
#pragma unroll
  for (int i = first; i < D::NLT; ++i) {
    double piv = M[sym_idx<ND>(i, i)];
    if (!(piv > 0.0)) { ok = 0; piv = 1e300; }
    const double ip = rcp(piv);
    M[sym_idx<ND>(i, i)] = ip;   // store inverse pivot
#pragma unroll
    for (int rr = i + 1; rr < ND; ++rr) {
      const double f = M[sym_idx<ND>(i, rr)] * ip;
#pragma unroll
      for (int c_ = rr; c_ < ND; ++c_) M[sym_idx<ND>(rr, c_)] -= f * M[sym_idx<ND>(i, c_)];
      r[rr] -= f * r[i];
    }
  }
  Sxx[0] = M[sym_idx<ND>(D::I_X, D::I_X)]; Sxx[1] = M[sym_idx<ND>(D::I_X, D::I_Y)]; Sxx[2] = M[sym_idx<ND>(D::I_X, D::I_P)];
  Sxx[3] = M[sym_idx<ND>(D::I_Y, D::I_Y)]; Sxx[4] = M[sym_idx<ND>(D::I_Y, D::I_P)]; Sxx[5] = M[sym_idx<ND>(D::I_P, D::I_P)];
  rx[0] = r[D::I_X]; rx[1] = r[D::I_Y]; rx[2] = r[D::I_P];
  // Local step as an affine map of the pose step:  x_loc = a + B (dX, dY, dpsi).  Four back-substitutions through the
  // eliminated rows (rhs: -r, and minus the three coupling columns); 4 NLT doubles go to the workspace instead of the
  // NM - 6 + NLT of the triangular factor, and the recovery becomes 3 multiply-adds per unknown.
  //   fac[(4 i + 0) fs] = a_i,   fac[(4 i + 1 + m) fs] = B_i,m   (m = X, Y, psi)
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    double x[D::NLT];
#pragma unroll
    for (int i = D::NLT - 1; i >= first; --i) {
      double acc = (m == 0) ? -r[i] : -M[sym_idx<ND>(i, D::NLT + m - 1)];
#pragma unroll
      for (int c_ = i + 1; c_ < D::NLT; ++c_) acc -= M[sym_idx<ND>(i, c_)] * x[c_];
      x[i] = acc * M[sym_idx<ND>(i, i)];      // stored inverse pivot
    }
    if (SDV) {
      double b0 = (m == 0) ? -r[0] : -M[sym_idx<ND>(0, D::NLT + m - 1)];
      double b1 = (m == 0) ? -r[1] : -M[sym_idx<ND>(1, D::NLT + m - 1)];
#pragma unroll
      for (int c_ = 2; c_ < D::NLT; ++c_) { b0 -= M[sym_idx<ND>(0, c_)] * x[c_]; b1 -= M[sym_idx<ND>(1, c_)] * x[c_]; }
      const double i00 = M[sym_idx<ND>(0, 0)], i01 = M[sym_idx<ND>(0, 1)], i11 = M[sym_idx<ND>(1, 1)];
      x[0] = i00 * b0 + i01 * b1;
      x[1] = i01 * b0 + i11 * b1;
    }
#pragma unroll
    for (int i = 0; i < D::NLT; ++i) fac[(size_t)(4 * i + m) * fs] = x[i];
  }
  return ok;
}

// Step of the local unknowns of one block after the pose step (dX, dY, dpsi) is known.
template <int VM>
struct ObsStep {
  double dlam[VM], dmu[4], dsl;
  double yn_new, yr1_new, yr2_new;
  double dsd, dsn;
};

template <int VM, bool SDV>
OBCA_HD void obs_recover(const ParkProblem& P, const ObsRows<VM>& R, const ObsVars<VM>& Q, const ObsGeom<VM>& G,
                         double mu_b, double dw, const double* fac, int fs, double dX, double dY, double dP,
                         ObsStep<VM>& S) {
  typedef LocalDims<VM, SDV> D;
  constexpr int ND = D::ND;
  double x[ND];
  x[D::I_X] = dX; x[D::I_Y] = dY; x[D::I_P] = dP;
#pragma unroll
  for (int i = 0; i < D::NLT; ++i)
    x[i] = fac[(size_t)(4 * i) * fs] + fac[(size_t)(4 * i + 1) * fs] * dX + fac[(size_t)(4 * i + 2) * fs] * dY +
           fac[(size_t)(4 * i + 3) * fs] * dP;
  const double g3 = P.g[2], g4 = P.g[3];
  double t3d = x[D::I_MU1] + G.e2 * dP, t4d = x[D::I_MU2] - G.e1 * dP;
  double Gd = (-P.g[0] - g3) * x[D::I_MU1] + (-P.g[1] - g4) * x[D::I_MU2] + G.p1 * dX + G.p2 * dY +
              (P.off * G.e2 - g3 * G.e2 + g4 * G.e1) * dP + (SDV ? x[D::I_SL] : 0.0);
  double gnd = 0.0;
#pragma unroll
  for (int i = 0; i < VM; ++i) {
    const double dl = (i < R.v) ? x[D::il(i)] : 0.0;
    S.dlam[i] = dl;
    if (i < R.v) {
      t3d += G.ah1[i] * dl; t4d += G.ah2[i] * dl;
      Gd += (G.rho[i] - g3 * G.ah1[i] - g4 * G.ah2[i]) * dl;
      gnd += G.gn[i] * dl;
    }
  }
  S.dmu[0] = x[D::I_MU1]; S.dmu[1] = x[D::I_MU2];
  S.dmu[2] = t3d + G.cr1;
  S.dmu[3] = t4d + G.cr2;
  S.dsl = SDV ? x[D::I_SL] : 0.0;
  const double igd = rcp(Q.sd - P.dmin);
  const double Sd = Q.vd * igd;
  const double ds = Gd + G.cd - g3 * G.cr1 - g4 * G.cr2;    // = grad(dist).d + (g - s)
  const double yd_new = -mu_b * igd + Sd * ds;               // new row multiplier of the dist row (= -vd_new)
  S.dsd = ds;
  const double im3 = rcp(Q.mu[2]), im4 = rcp(Q.mu[3]);
  const double s3 = Q.zmu[2] * im3 + dw, s4 = Q.zmu[3] * im4 + dw;
  S.yr1_new = s3 * S.dmu[2] - g3 * yd_new - mu_b * im3;
  S.yr2_new = s4 * S.dmu[3] - g4 * yd_new - mu_b * im4;
  S.yn_new = SDV ? x[1] : 0.0;
  S.dsn = SDV ? 0.0 : (gnd + G.cn);
}

// Lagrangian-gradient pieces of one block with the CURRENT multipliers (for the KKT error).
//   rl[VM], rm[4], rs: stationarity residuals of lambda, mu, sl;   gx[3]: contribution to the rows of (X, Y, psi)
template <int VM, bool SDV>
OBCA_HD void obs_lagr_grad(const ParkProblem& P, const ObsRows<VM>& R, const ObsVars<VM>& Q, const ObsGeom<VM>& G,
                           double* rl, double* rm, double& rs, double* gx) {
  const double yId = -Q.vd;
  const double ynorm = SDV ? Q.yn : Q.vn;
#pragma unroll
  for (int i = 0; i < VM; ++i)
    rl[i] = (i < R.v) ? (ynorm * G.gn[i] + Q.yr1 * G.ah1[i] + Q.yr2 * G.ah2[i] + yId * G.rho[i] - Q.zlam[i]) : 0.0;
  rm[0] = Q.yr1 - yId * P.g[0] - Q.zmu[0];
  rm[1] = Q.yr2 - yId * P.g[1] - Q.zmu[1];
  rm[2] = -Q.yr1 - yId * P.g[2] - Q.zmu[2];
  rm[3] = -Q.yr2 - yId * P.g[3] - Q.zmu[3];
  rs = SDV ? (1.0e2 + 2.0e4 * Q.sl + yId) : 0.0;
  gx[0] = yId * G.p1;
  gx[1] = yId * G.p2;
  gx[2] = Q.yr1 * G.e2 - Q.yr2 * G.e1 + yId * P.off * G.e2;
}

}  // namespace obca

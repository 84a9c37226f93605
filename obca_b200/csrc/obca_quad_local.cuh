// obca_quad_local.cuh -- the OBCA block of one (stage k, box obstacle o) pair of the quadcopter NLPs.
//
// Reference rows (QuadcopterNavigation/QuadcopterSignedDist.jl:165-197, QuadcopterDist.jl:159-191), A = [I; -I] (:162-163)
// so p = A' lam = lam[0:3] - lam[3:6]; ball ego of radius R (no mu multipliers):
//   norm : p1^2 + p2^2 + p3^2 == 1                     (== in BOTH variants)
//   dist : -b' lam + X p1 + Y p2 + Z p3 [+ 0.01 slack] >= R
// plus lam >= 0, slack >= 0 (:98-105), the costs 1e2 slack + 1e3 slack^2 and reg2 * lam^2 (:69-71).
//
// Same elimination scheme as obca_local.cuh: dist slack (quasi-definite) -> (lam_piv, y_norm) 2x2 pivot ->
// remaining lam, slack by 1x1 pivots that must be positive -> 3x3 Schur complement on the position (X, Y, Z).
// Unknown order: [lam_piv, y_norm, lam (5 others), slack (SD) | X, Y, Z].
#pragma once
#include "obca_common.cuh"

namespace obca {

constexpr double QUAD_REG2 = 1e-4;   // QuadcopterSignedDist.jl:54
constexpr double QUAD_SLK_COEF = 0.01;

template <bool SDV>
struct QLocalDims {
  static constexpr int SLN = SDV ? 1 : 0;
  static constexpr int NLT = 6 + 1 + SLN;       // eliminated unknowns (6 lam, y_norm, slack)
  static constexpr int I_SL = 7;
  static constexpr int I_X = NLT, I_Y = NLT + 1, I_Z = NLT + 2;
  static constexpr int ND = NLT + 3;
  static constexpr int NM = ND * (ND + 1) / 2;
  static constexpr int NFAC = NM - 6 + NLT;
  OBCA_HD static constexpr int il(int i) { return i == 0 ? 0 : i + 1; }   // position of lam_i (y_norm sits at 1)
};

struct QObsVars {
  double lam[6], zlam[6];
  double bb[6];        // box vector b of the obstacle (permuted together with lam)
  double sg[6];        // +1 for lam[0:3], -1 for lam[3:6]
  int cp[6];           // position component 0..2 the row acts on
  double sl, zsl;      // slack >= 0 (SD)
  double yn;           // multiplier of the norm row
  double sd, vd;       // dist slack (>= R) and its bound multiplier
};

struct QObsGeom {
  double p[3];
  double gn[6], rho[6];
  double gd;           // dist expression value (incl. 0.01 slack for SD)
  double cn, cd;       // residuals: p.p - 1 and gd - sd
};

OBCA_HD void qobs_load_const(QObsVars& Q, const double* b6) {
#pragma unroll
  for (int i = 0; i < 6; ++i) { Q.bb[i] = b6[i]; Q.sg[i] = i < 3 ? 1.0 : -1.0; Q.cp[i] = i % 3; }
}

template <bool SDV>
OBCA_HD void qobs_geom(const double* pos, const QObsVars& Q, QObsGeom& G) {
  G.p[0] = G.p[1] = G.p[2] = 0.0;
  double bl = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    bl += Q.bb[i] * Q.lam[i];
#pragma unroll
    for (int c = 0; c < 3; ++c) G.p[c] += (Q.cp[i] == c) ? Q.sg[i] * Q.lam[i] : 0.0;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double pc = 0.0, xc = 0.0;
#pragma unroll
    for (int c = 0; c < 3; ++c) { pc = (Q.cp[i] == c) ? G.p[c] : pc; xc = (Q.cp[i] == c) ? pos[c] : xc; }
    G.gn[i] = 2.0 * Q.sg[i] * pc;
    G.rho[i] = -Q.bb[i] + Q.sg[i] * xc;
  }
  G.gd = -bl + pos[0] * G.p[0] + pos[1] * G.p[1] + pos[2] * G.p[2] + (SDV ? QUAD_SLK_COEF * Q.sl : 0.0);
  G.cn = G.p[0] * G.p[0] + G.p[1] * G.p[1] + G.p[2] * G.p[2] - 1.0;
  G.cd = G.gd - Q.sd;
}

OBCA_HD int qobs_choose_pivot(const QObsGeom& G) {
  int piv = 0;
  double best = dabs(G.gn[0]);
#pragma unroll
  for (int i = 1; i < 6; ++i)
    if (dabs(G.gn[i]) > best) { best = dabs(G.gn[i]); piv = i; }
  return piv;
}

OBCA_HD void qobs_swap(QObsVars& Q, int piv) {
#pragma unroll
  for (int i = 1; i < 6; ++i) {
    if (i == piv) {
      double t;
      t = Q.lam[0]; Q.lam[0] = Q.lam[i]; Q.lam[i] = t;
      t = Q.zlam[0]; Q.zlam[0] = Q.zlam[i]; Q.zlam[i] = t;
      t = Q.bb[0]; Q.bb[0] = Q.bb[i]; Q.bb[i] = t;
      t = Q.sg[0]; Q.sg[0] = Q.sg[i]; Q.sg[i] = t;
      int c = Q.cp[0]; Q.cp[0] = Q.cp[i]; Q.cp[i] = c;
    }
  }
}

// Dual warm start of one (position, box) pair -- the closed-form analogue of DualMultWS.jl for a ball ego and a box
// {lo <= y <= hi}, b = [hi; -lo]: lam = positive / negative parts of the unit vector from the closest box point to the
// position, so that A'lam has norm 1 and -b'lam + pos.A'lam equals the distance.  It replaces the reference's
// l = 0.05 (QuadcopterSignedDist.jl:204-208), for which A'lam = 0 and the norm row has an empty linearisation; the NLP
// and its KKT points are unchanged, only the starting multipliers differ (documented deviation).
OBCA_HD void quad_dual_ws(const double* pos, const double* b6, double* lam) {
  double d[3], n2 = 0.0, pen = 1e300;
  int face = 0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double hi = b6[c], lo = -b6[3 + c];
    const double cl = pos[c] < lo ? lo : (pos[c] > hi ? hi : pos[c]);
    d[c] = pos[c] - cl;
    n2 += d[c] * d[c];
    // inside the box along c: distances to the two faces (used only if the point is inside the box)
    if (hi - pos[c] < pen) { pen = hi - pos[c]; face = c; }
    if (pos[c] - lo < pen) { pen = pos[c] - lo; face = 3 + c; }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) lam[i] = 0.0;
  if (n2 > 1e-24) {
    const double inv = 1.0 / sqrt(n2);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double pc = d[c] * inv;
      lam[c] = pc > 0.0 ? pc : 0.0;
      lam[3 + c] = pc < 0.0 ? -pc : 0.0;
    }
  } else {
    lam[face] = 1.0;     // inside the box: outward normal of the nearest face
  }
}

// Norm rows whose linearisation is (numerically) empty -- p = 0, the reference's own starting point l = 0.05
// (QuadcopterSignedDist.jl:204-208) -- keep their multiplier for this iteration.
constexpr double QUAD_FREEZE = 1e-6;

template <bool SDV>
OBCA_HD int qobs_condense(double R_ego, const QObsVars& Q, const QObsGeom& G, double mu_b, double dw, double dc,
                          double* Sxx, double* rx, double* fac, int fs) {
  typedef QLocalDims<SDV> D;
  constexpr int ND = D::ND;
  double M[D::NM], r[ND], Gt[ND];
#pragma unroll
  for (int i = 0; i < D::NM; ++i) M[i] = 0.0;
#pragma unroll
  for (int i = 0; i < ND; ++i) { r[i] = 0.0; Gt[i] = 0.0; }
  const double yId = -Q.vd;
  const double igd = rcp(Q.sd - R_ego);
  const double Sd = Q.vd * igd;
  const double yd0 = -mu_b * igd + Sd * G.cd;
#pragma unroll
  for (int i = 0; i < 6; ++i) Gt[D::il(i)] = G.rho[i];
  if (SDV) Gt[D::I_SL] = QUAD_SLK_COEF;
  Gt[D::I_X] = G.p[0]; Gt[D::I_Y] = G.p[1]; Gt[D::I_Z] = G.p[2];
#pragma unroll
  for (int i = 0; i < ND; ++i) {
    if (i == 1) continue;
#pragma unroll
    for (int j = i; j < ND; ++j) {
      if (j == 1) continue;
      M[sym_idx<ND>(i, j)] = Sd * Gt[i] * Gt[j];
    }
    r[i] = Gt[i] * yd0;
  }
  double gmax = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int li = D::il(i);
#pragma unroll
    for (int l = i; l < 6; ++l)
      if (Q.cp[i] == Q.cp[l]) M[sym_idx_any<ND>(li, D::il(l))] += 2.0 * Q.yn * Q.sg[i] * Q.sg[l];
    const double il_ = rcp(Q.lam[i]);
    M[sym_idx<ND>(li, li)] += Q.zlam[i] * il_ + dw + 2.0 * QUAD_REG2;
    r[li] += 2.0 * QUAD_REG2 * Q.lam[i] - mu_b * il_;
    M[sym_idx_any<ND>(li, 1)] = G.gn[i];
    gmax = dmax(gmax, dabs(G.gn[i]));
    // d2 dist / d pos_c d lam_i = sg_i (c == cp_i), weighted by the row multiplier of dist
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (Q.cp[i] == c) M[sym_idx_any<ND>(li, D::I_X + c)] += yId * Q.sg[i];
  }
  if (SDV) {
    const double is_ = rcp(Q.sl);
    M[sym_idx<ND>(D::I_SL, D::I_SL)] += Q.zsl * is_ + dw + 2.0e3;
    r[D::I_SL] += 1.0e2 + 2.0e3 * Q.sl - mu_b * is_;
  }
  M[sym_idx<ND>(1, 1)] = -dc;
  r[1] = (gmax < QUAD_FREEZE ? 0.0 : G.cn) + dc * Q.yn;

  int ok = 1;
  {
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
    M[sym_idx<ND>(0, 0)] = i00; M[sym_idx<ND>(0, 1)] = i01; M[sym_idx<ND>(1, 1)] = i11;
  }
#pragma unroll
  for (int i = 2; i < D::NLT; ++i) {
    double piv = M[sym_idx<ND>(i, i)];
    if (!(piv > 0.0)) { ok = 0; piv = 1e300; }
    const double ip = rcp(piv);
    M[sym_idx<ND>(i, i)] = ip;
#pragma unroll
    for (int rr = i + 1; rr < ND; ++rr) {
      const double f = M[sym_idx<ND>(i, rr)] * ip;
#pragma unroll
      for (int c_ = rr; c_ < ND; ++c_) M[sym_idx<ND>(rr, c_)] -= f * M[sym_idx<ND>(i, c_)];
      r[rr] -= f * r[i];
    }
  }
  Sxx[0] = M[sym_idx<ND>(D::I_X, D::I_X)]; Sxx[1] = M[sym_idx<ND>(D::I_X, D::I_Y)]; Sxx[2] = M[sym_idx<ND>(D::I_X, D::I_Z)];
  Sxx[3] = M[sym_idx<ND>(D::I_Y, D::I_Y)]; Sxx[4] = M[sym_idx<ND>(D::I_Y, D::I_Z)]; Sxx[5] = M[sym_idx<ND>(D::I_Z, D::I_Z)];
  rx[0] = r[D::I_X]; rx[1] = r[D::I_Y]; rx[2] = r[D::I_Z];
  {
    int q = 0;
#pragma unroll
    for (int i = 0; i < D::NLT; ++i) {
#pragma unroll
      for (int c_ = i; c_ < ND; ++c_) { fac[(size_t)q * fs] = M[sym_idx<ND>(i, c_)]; ++q; }
    }
#pragma unroll
    for (int i = 0; i < D::NLT; ++i) { fac[(size_t)q * fs] = r[i]; ++q; }
  }
  return ok;
}

struct QObsStep {
  double dlam[6], dsl, yn_new, dsd;
};

template <bool SDV>
OBCA_HD void qobs_recover(double R_ego, const QObsVars& Q, const QObsGeom& G, double mu_b, const double* fac, int fs,
                          const double* dpos, QObsStep& S) {
  typedef QLocalDims<SDV> D;
  constexpr int ND = D::ND;
  double x[ND];
  x[D::I_X] = dpos[0]; x[D::I_Y] = dpos[1]; x[D::I_Z] = dpos[2];
#define OBCA_ROFF(i) ((i) * ND - ((i) * ((i) - 1)) / 2)
#define OBCA_FAC(e) fac[(size_t)(e) * fs]
  constexpr int RH = D::NM - 6;
#pragma unroll
  for (int i = D::NLT - 1; i >= 2; --i) {
    double acc = -OBCA_FAC(RH + i);
#pragma unroll
    for (int c_ = i + 1; c_ < ND; ++c_) acc -= OBCA_FAC(OBCA_ROFF(i) + (c_ - i)) * x[c_];
    x[i] = acc * OBCA_FAC(OBCA_ROFF(i));
  }
  {
    double b0 = -OBCA_FAC(RH + 0), b1 = -OBCA_FAC(RH + 1);
#pragma unroll
    for (int c_ = 2; c_ < ND; ++c_) { b0 -= OBCA_FAC(OBCA_ROFF(0) + c_) * x[c_]; b1 -= OBCA_FAC(OBCA_ROFF(1) + (c_ - 1)) * x[c_]; }
    const double i00 = OBCA_FAC(OBCA_ROFF(0)), i01 = OBCA_FAC(OBCA_ROFF(0) + 1), i11 = OBCA_FAC(OBCA_ROFF(1));
    x[0] = i00 * b0 + i01 * b1;
    x[1] = i01 * b0 + i11 * b1;
  }
#undef OBCA_ROFF
#undef OBCA_FAC
  double Gd = G.p[0] * dpos[0] + G.p[1] * dpos[1] + G.p[2] * dpos[2] + (SDV ? QUAD_SLK_COEF * x[D::I_SL] : 0.0);
#pragma unroll
  for (int i = 0; i < 6; ++i) { S.dlam[i] = x[D::il(i)]; Gd += G.rho[i] * S.dlam[i]; }
  S.dsl = SDV ? x[D::I_SL] : 0.0;
  S.yn_new = x[1];
  S.dsd = Gd + G.cd;
}

// Lagrangian-gradient pieces with the current multipliers: rl[6] (lam rows), rs (slack row), gx[3] (position rows)
template <bool SDV>
OBCA_HD void qobs_lagr_grad(const QObsVars& Q, const QObsGeom& G, double* rl, double& rs, double* gx) {
  const double yId = -Q.vd;
#pragma unroll
  for (int i = 0; i < 6; ++i) rl[i] = 2.0 * QUAD_REG2 * Q.lam[i] + Q.yn * G.gn[i] + yId * G.rho[i] - Q.zlam[i];
  rs = SDV ? (1.0e2 + 2.0e3 * Q.sl + QUAD_SLK_COEF * yId - Q.zsl) : 0.0;
  gx[0] = yId * G.p[0]; gx[1] = yId * G.p[1]; gx[2] = yId * G.p[2];
}

}  // namespace obca

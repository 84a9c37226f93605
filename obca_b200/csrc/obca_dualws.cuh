// obca_dualws.cuh -- K2: dual-multiplier warm start, one tiny convex program per (stage i, obstacle j).
//
// Reference: AutonomousParking/DualMultWS.jl:29-86.  The JuMP model there is separable -- no row couples two
// different (i, j) pairs (:57-75) -- so the (N+1)*nOb programs
//     max d = -g'mu + (A_j t_i - b_j)'lam   s.t.  |A_j'lam|^2 <= 1 (:65),  G'mu + R(psi_i)'A_j'lam = 0 (:68-69),
//                                                  lam >= 0, mu >= 0 (:54-55)
// are solved independently, each by a register-resident primal-dual interior-point iteration that follows the
// same Ipopt scheme the reference uses (tol 1e-5, max_iter 100, start at 0 pushed into the bounds, :36-37).
#pragma once
#include "obca_common.cuh"
#include "obca_local.cuh"

namespace obca {

template <int VM>
OBCA_HD int dualws_solve(const ParkProblem& P, const ObsRows<VM>& R, double X, double Y, double psi, double tol,
                         int max_iter, double* lam_out, double* mu_out, double* d_out) {
  constexpr int NU = VM + 2;   // unknowns after eliminating mu3, mu4 through the rot rows
  double sn_, cs_;
  sincos(psi, &sn_, &cs_);
  const double tcx = X + cs_ * P.off, tcy = Y + sn_ * P.off;
  double ah1[VM], ah2[VM], rho[VM];
#pragma unroll
  for (int i = 0; i < VM; ++i) {
    ah1[i] = cs_ * R.a1[i] + sn_ * R.a2[i];
    ah2[i] = -sn_ * R.a1[i] + cs_ * R.a2[i];
    rho[i] = R.a1[i] * tcx + R.a2[i] * tcy - R.bb[i];
  }
  const double g[4] = {P.g[0], P.g[1], P.g[2], P.g[3]};
  // iterate
  double lam[VM], zlam[VM], mu[4], zmu[4];
  const double push = 1e-2;
#pragma unroll
  for (int i = 0; i < VM; ++i) { lam[i] = push; zlam[i] = 1.0; }
#pragma unroll
  for (int m = 0; m < 4; ++m) { mu[m] = push; zmu[m] = 1.0; }
  double p1 = 0, p2 = 0;
#pragma unroll
  for (int i = 0; i < VM; ++i) if (i < R.v) { p1 += R.a1[i] * lam[i]; p2 += R.a2[i] * lam[i]; }
  double sn = dmin_(p1 * p1 + p2 * p2, 1.0 - push), vn = 1.0;   // norm slack <= 1 and its multiplier
  double yr1 = 0.0, yr2 = 0.0;
  double mub = 0.1;
  const double mu_min = tol / 10.0;
  int it = 0, status = 0;
  for (;; ++it) {
    p1 = 0; p2 = 0;
#pragma unroll
    for (int i = 0; i < VM; ++i) if (i < R.v) { p1 += R.a1[i] * lam[i]; p2 += R.a2[i] * lam[i]; }
    const double e1 = cs_ * p1 + sn_ * p2, e2 = -sn_ * p1 + cs_ * p2;
    const double cn = p1 * p1 + p2 * p2 - sn;
    const double cr1 = mu[0] - mu[2] + e1, cr2 = mu[1] - mu[3] + e2;
    double gn[VM];
#pragma unroll
    for (int i = 0; i < VM; ++i) gn[i] = 2.0 * (R.a1[i] * p1 + R.a2[i] * p2);
    // KKT error (objective = minimise  g'mu - rho'lam)
    double e_dual = 0.0, cmax = 0.0, cmin = 1e300, sy = dabs(yr1) + dabs(yr2) + vn, sz = vn;
#pragma unroll
    for (int i = 0; i < VM; ++i)
      if (i < R.v) {
        e_dual = dmax(e_dual, dabs(-rho[i] + vn * gn[i] + yr1 * ah1[i] + yr2 * ah2[i] - zlam[i]));
        cmax = dmax(cmax, lam[i] * zlam[i]); cmin = dmin_(cmin, lam[i] * zlam[i]); sz += zlam[i];
      }
    {
      const double rm[4] = {g[0] + yr1 - zmu[0], g[1] + yr2 - zmu[1], g[2] - yr1 - zmu[2], g[3] - yr2 - zmu[3]};
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        e_dual = dmax(e_dual, dabs(rm[m]));
        cmax = dmax(cmax, mu[m] * zmu[m]); cmin = dmin_(cmin, mu[m] * zmu[m]); sz += zmu[m];
      }
    }
    cmax = dmax(cmax, (1.0 - sn) * vn); cmin = dmin_(cmin, (1.0 - sn) * vn);
    const double e_pr = dmax(dabs(cn), dmax(dabs(cr1), dabs(cr2)));
    const double nm = (double)(R.v + 4 + 1 + 2 + 1), nb = (double)(R.v + 4 + 1);
    const double sd = dmax(100.0, (sy + sz) / nm) / 100.0, sc = dmax(100.0, sz / nb) / 100.0;
    const double e0 = dmax(dmax(e_dual / sd, e_pr), cmax / sc);
    if (e0 <= tol) { status = 1; break; }
    if (it >= max_iter) break;
    while (mub > mu_min && dmax(dmax(e_dual / sd, e_pr), dmax(cmax - mub, mub - cmin) / sc) <= 10.0 * mub)
      mub = dmax(mu_min, dmin_(0.2 * mub, pow(mub, 1.5)));
    const double tau = dmax(0.99, 1.0 - mub);
    // ---- Newton system on (lam, mu1, mu2) ----
    const double s3 = zmu[2] / mu[2], s4 = zmu[3] / mu[3];
    const double gapn = 1.0 - sn, Sn = vn / gapn, yn0 = mub / gapn + Sn * cn;
    const double c3 = s3 * cr1 + g[2] - mub / mu[2], c4 = s4 * cr2 + g[3] - mub / mu[3];
    double M[NU * (NU + 1) / 2], r[NU], t3[NU], t4[NU];
#pragma unroll
    for (int i = 0; i < NU; ++i) { t3[i] = 0.0; t4[i] = 0.0; }
#pragma unroll
    for (int i = 0; i < VM; ++i) if (i < R.v) { t3[i] = ah1[i]; t4[i] = ah2[i]; }
    t3[VM] = 1.0; t4[VM + 1] = 1.0;
#pragma unroll
    for (int i = 0; i < NU; ++i) {
#pragma unroll
      for (int j = i; j < NU; ++j) M[sym_idx<NU>(i, j)] = s3 * t3[i] * t3[j] + s4 * t4[i] * t4[j];
      r[i] = t3[i] * c3 + t4[i] * c4;
    }
#pragma unroll
    for (int i = 0; i < VM; ++i) {
      if (i < R.v) {
#pragma unroll
        for (int l = i; l < VM; ++l)
          if (l < R.v) M[sym_idx<NU>(i, l)] += 2.0 * vn * (R.a1[i] * R.a1[l] + R.a2[i] * R.a2[l]) + Sn * gn[i] * gn[l];
        M[sym_idx<NU>(i, i)] += zlam[i] / lam[i];
        r[i] += -rho[i] - mub / lam[i] + gn[i] * yn0;
      } else {
#pragma unroll
        for (int j = 0; j < NU; ++j) M[sym_idx_any<NU>(i, j)] = 0.0;
        M[sym_idx<NU>(i, i)] = 1.0; r[i] = 0.0;
      }
    }
    M[sym_idx<NU>(VM, VM)] += zmu[0] / mu[0];
    M[sym_idx<NU>(VM + 1, VM + 1)] += zmu[1] / mu[1];
    r[VM] += g[0] - mub / mu[0];
    r[VM + 1] += g[1] - mub / mu[1];
    // SPD elimination
    double x[NU];
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      const double ip = 1.0 / M[sym_idx<NU>(i, i)];
      M[sym_idx<NU>(i, i)] = ip;
#pragma unroll
      for (int rr = i + 1; rr < NU; ++rr) {
        const double f = M[sym_idx<NU>(i, rr)] * ip;
#pragma unroll
        for (int c_ = rr; c_ < NU; ++c_) M[sym_idx<NU>(rr, c_)] -= f * M[sym_idx<NU>(i, c_)];
        r[rr] -= f * r[i];
      }
    }
#pragma unroll
    for (int i = NU - 1; i >= 0; --i) {
      double acc = -r[i];
#pragma unroll
      for (int c_ = i + 1; c_ < NU; ++c_) acc -= M[sym_idx<NU>(i, c_)] * x[c_];
      x[i] = acc * M[sym_idx<NU>(i, i)];
    }
    double dlam[VM], dmu[4], t3d = x[VM], t4d = x[VM + 1], gnd = 0.0;
#pragma unroll
    for (int i = 0; i < VM; ++i) {
      dlam[i] = (i < R.v) ? x[i] : 0.0;
      t3d += ah1[i] * dlam[i]; t4d += ah2[i] * dlam[i]; gnd += gn[i] * dlam[i];
    }
    dmu[0] = x[VM]; dmu[1] = x[VM + 1]; dmu[2] = t3d + cr1; dmu[3] = t4d + cr2;
    const double dsn = gnd + cn;
    const double yr1n = s3 * dmu[2] + g[2] - mub / mu[2], yr2n = s4 * dmu[3] + g[3] - mub / mu[3];
    // step lengths
    double apr = 1.0, adu = 1.0;
    double dzl[VM], dzm[4];
#pragma unroll
    for (int i = 0; i < VM; ++i) {
      dzl[i] = 0.0;
      if (i < R.v) {
        if (dlam[i] < 0) apr = dmin_(apr, -tau * lam[i] / dlam[i]);
        dzl[i] = mub / lam[i] - zlam[i] - zlam[i] / lam[i] * dlam[i];
        if (dzl[i] < 0) adu = dmin_(adu, -tau * zlam[i] / dzl[i]);
      }
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      if (dmu[m] < 0) apr = dmin_(apr, -tau * mu[m] / dmu[m]);
      dzm[m] = mub / mu[m] - zmu[m] - zmu[m] / mu[m] * dmu[m];
      if (dzm[m] < 0) adu = dmin_(adu, -tau * zmu[m] / dzm[m]);
    }
    if (dsn > 0) apr = dmin_(apr, tau * gapn / dsn);
    const double dvn = mub / gapn - vn + vn / gapn * dsn;
    if (dvn < 0) adu = dmin_(adu, -tau * vn / dvn);
    const double ay = dmin_(apr, adu);
#pragma unroll
    for (int i = 0; i < VM; ++i) if (i < R.v) { lam[i] += apr * dlam[i]; zlam[i] += adu * dzl[i]; }
#pragma unroll
    for (int m = 0; m < 4; ++m) { mu[m] += apr * dmu[m]; zmu[m] += adu * dzm[m]; }
    sn += apr * dsn; vn += adu * dvn;
    yr1 += ay * (yr1n - yr1); yr2 += ay * (yr2n - yr2);
  }
  double d = 0.0;
#pragma unroll
  for (int i = 0; i < VM; ++i) { lam_out[i] = (i < R.v) ? lam[i] : 0.0; if (i < R.v) d += rho[i] * lam[i]; }
#pragma unroll
  for (int m = 0; m < 4; ++m) { mu_out[m] = mu[m]; d -= g[m] * mu[m]; }
  *d_out = d;
  return status ? it : -it - 1;
}

}  // namespace obca

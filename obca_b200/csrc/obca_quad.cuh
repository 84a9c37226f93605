// obca_quad.cuh -- model policy of the quadcopter NLPs for the generic interior-point driver (obca_solver.cuh).
//
// Replaces the JuMP + Ipopt solve inside QuadcopterNavigation/QuadcopterSignedDist.jl:25-300 (SDV = true) and
// QuadcopterDist.jl:25-282 (SDV = false): 12-state quadrotor, 4 rotor speeds, ball ego, five box obstacles, time scaling.
// Same exact reformulations as the parking model (pinned end states are parameters, one time-scale variable with
// multiplicity N+1, bounds are bounds) plus quirk Q5: the body-rate products of :153-155 are the
// stage-1 values x0[9..11] (constants).
//
// Stage vector of the KKT sweep:  y = [ x (12) | w = previous control (4) | t | u (4) ],  17 states + 4 controls.
// First version of this path: the stage models live in the per-CTA global workspace and the Riccati sweep is a plain
// dense recursion run by one thread (the warp-cooperative shared-memory version of the parking model is the template
// for the next round).
#pragma once
#include "obca_quad_dyn_gen.cuh"
#include "obca_quad_local.cuh"
#include "obca_solver.cuh"

namespace obca {

constexpr int QNX = 12, QNU = 4, QNSV = 17, QNYV = 21;
constexpr int QIW = 12, QIT = 16, QIU = 17;
constexpr int QNQ = QNYV * (QNYV + 1) / 2;     // 231
constexpr int QNP = QNSV * (QNSV + 1) / 2;     // 153
constexpr double QUAD_REG3 = 1e-4;              // :55
constexpr double QUAD_WH = 4.479906037125444;   // sqrt(mass*g/(4 k_F)) = sqrt(0.5*9.81/(0.0611*4)), :62
constexpr int QNOB = 5;

struct QuadProblem {
  int N;
  double Ts, R;
  double obs[QNOB][6];
  double xlo[QNX], xhi[QNX];   // :78-93 (x10 bounds differ between the variants, QuadcopterDist.jl:88)
  int signed_dist;
};

struct QLay {
  int NSP;
  int X, U, LAM, SLK;
  int ZXL, ZXU, ZUL, ZUU, ZLAM, ZSLK;
  int PI, YN, SD, VD;
  int dX, dU, dLAM, dSLK, PIn, YNn, dSD;
  int LF, nfac;
  int QS, qs, JV, R12, RK, RP;
  int total;
};

inline QLay make_qlayout(const QuadProblem& P) {
  QLay L;
  const int NS = P.N + 1;
  L.NSP = ((NS + 31) / 32) * 32;
  int c = 0;
  auto take = [&](int n) { int o = c; c += n; return o; };
  L.X = take(QNX); L.U = take(QNU); L.LAM = take(6 * QNOB); L.SLK = take(QNOB);
  L.ZXL = take(QNX); L.ZXU = take(QNX); L.ZUL = take(QNU); L.ZUU = take(QNU); L.ZLAM = take(6 * QNOB); L.ZSLK = take(QNOB);
  L.PI = take(QNX); L.YN = take(QNOB); L.SD = take(QNOB); L.VD = take(QNOB);
  L.dX = take(QNX); L.dU = take(QNU); L.dLAM = take(6 * QNOB); L.dSLK = take(QNOB); L.PIn = take(QNX); L.YNn = take(QNOB);
  L.dSD = take(QNOB);
  L.nfac = P.signed_dist ? QLocalDims<true>::NFAC : QLocalDims<false>::NFAC;
  L.LF = take(QNOB * L.nfac);
  L.QS = take(QNQ); L.qs = take(QNYV); L.JV = take(QD_NJ); L.R12 = take(QNX);
  L.RK = take(QNU * QNSV + QNU); L.RP = take(QNX * QNSV + QNX);
  L.total = c;
  return L;
}

struct QInputs {
  const double* x0;    // 12
  const double* xF;    // 12
  const double* xWS;   // 12 x (N+1) column-major (QuadcopterSignedDist.jl:201: setvalue(x, xWS))
  double timeWS;       // :199
};
struct QOutputs {
  double *xp, *up, *ts, *lp, *slack;   // 12x(N+1), 4xN, (N+1), 30x(N+1), 5x(N+1) (:277-298)
};

struct QCtx {
  const QuadProblem* P;
  const IpmOpts* O;
  QLay L;
  double* W;
  void* red_scratch;
  double* tile;
  ProbState* S;
  QInputs in;
};

// Problem, options and layout: __constant__ memory on the device (as for the parking model, obca_solver.cuh), the context on the host.
#if defined(__CUDACC__)
__constant__ QuadProblem c_qP;
__constant__ QLay c_qL;
#endif
#if defined(__CUDA_ARCH__)
#define QCTX_P(C) c_qP
#define QCTX_O(C) c_pkO
#define QCTX_L(C) c_qL
// the workspace is global memory: lets the compiler emit LDG / STG instead of generic accesses
#define QLOCALS(C) double* const W_ = (C).W; __builtin_assume(__isGlobal(W_)); (void)W_
#else
#define QCTX_P(C) (*(C).P)
#define QCTX_O(C) (*(C).O)
#define QCTX_L(C) ((C).L)
#define QLOCALS(C) double* const W_ = (C).W; (void)W_
#endif
#define QA(name, i, k) (W_[(size_t)(QCTX_L(C).name + (i)) * QCTX_L(C).NSP + (k)])

OBCA_HD void quad_jac_tables(const int*& jr, const int*& jc, const int*& hi, const int*& hj) {
  static constexpr int JR[QD_NJ] = OBCA_QD_J_ROW;
  static constexpr int JC[QD_NJ] = OBCA_QD_J_COL;
  static constexpr int HI[QD_NH] = OBCA_QD_H_I;
  static constexpr int HJ[QD_NH] = OBCA_QD_H_J;
  jr = JR; jc = JC; hi = HI; hj = HJ;
}

#ifdef OBCA_QPROF
__device__ unsigned long long g_qprof[8];   // development: cycles of the sweep's phases (thread 0 of every CTA)
#endif
template <bool SDV>
struct QuadSolver {
  typedef QCtx Ctx;
  typedef QLocalDims<SDV> LD;

  OBCA_HD static int n_stages(const QCtx& C) { return QCTX_P(C).N + 1; }
  OBCA_HD static bool fixed_time(const QCtx&) { return false; }
  OBCA_HD static void mult_counts(const QCtx& C, double& n_mult, double& n_bmult) {
    const double N = QCTX_P(C).N, NS = N + 1;
    double nb = 24.0 * (N - 1) + 8.0 * N + 30.0 * NS + 2.0 * NS + (SDV ? 5.0 * NS : 0.0);   // variable bounds
    nb += 5.0 * NS;                                                                          // dist slack bounds
    n_bmult = nb; n_mult = nb + 12.0 * N + 5.0 * NS + 5.0 * NS;
  }
  OBCA_HD static void init_scalars(const QCtx& C, int restart) {
    ProbState& S = *C.S;
    S.t = push_lo(restart ? S.t : C.in.timeWS, 0.5, 2.0, QCTX_O(C).kappa1, QCTX_O(C).kappa2);   // :96, :199
    S.zTL = 1.0; S.zTU = 1.0; S.dt = 0.0;
  }
  OBCA_HD static void update_scalars(const QCtx& C) {
    ProbState& S = *C.S;
    double q = S.t, zl = S.zTL, zu = S.zTU;
    upd_pair(q, S.dt, zl, zu, 0.5, 2.0, S.alpha, S.a_du, S.mu, QCTX_O(C).kappa_sigma);
    S.t = q; S.zTL = zl; S.zTU = zu;
  }
  OBCA_HD static void ftb(double gap, double dgap, double tau, double& amax) {
    if (dgap < 0.0) amax = dmin_(amax, -tau * gap * rcp(dgap));
  }
  OBCA_HD static double dzb(double z, double gap, double dgap, double mu_b) { return (mu_b - z * dgap) * rcp(gap) - z; }
  OBCA_HD static double clipz(double z, double gap, double mu_b, double ks) {
    const double mg = mu_b * rcp(gap);
    return dmax(dmin_(z, ks * mg), mg * rcp(ks));
  }
  OBCA_HD static void upd_pair(double& q, double dq, double& zl, double& zu, double lo, double hi, double alpha,
                               double adu, double mu_b, double ks) {
    const double gl = q - lo, gu = hi - q;
    const double dzl = dzb(zl, gl, dq, mu_b), dzu = dzb(zu, gu, -dq, mu_b);
    q += alpha * dq; zl += adu * dzl; zu += adu * dzu;
    zl = clipz(zl, q - lo, mu_b, ks); zu = clipz(zu, hi - q, mu_b, ks);
  }

  OBCA_HD static void load_obs(const QCtx& C, int k, int o, QObsVars& Q, double alpha) {
    QLOCALS(C);
    qobs_load_const(Q, QCTX_P(C).obs[o]);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      Q.lam[i] = QA(LAM, 6 * o + i, k) + (alpha != 0.0 ? alpha * QA(dLAM, 6 * o + i, k) : 0.0);
      Q.zlam[i] = QA(ZLAM, 6 * o + i, k);
    }
    Q.sl = SDV ? QA(SLK, o, k) + (alpha != 0.0 ? alpha * QA(dSLK, o, k) : 0.0) : 0.0;
    Q.zsl = SDV ? QA(ZSLK, o, k) : 0.0;
    Q.yn = QA(YN, o, k);
    Q.sd = QA(SD, o, k) + (alpha != 0.0 ? alpha * QA(dSD, o, k) : 0.0);
    Q.vd = QA(VD, o, k);
  }

  // ---- P0 ----
  OBCA_HD_NI static void init_stage(const QCtx& C, int k, int restart) {
    QLOCALS(C);
    const QuadProblem& P = QCTX_P(C);
    const IpmOpts& O = QCTX_O(C);
    const int N = P.N;
    const bool free_x = (k >= 1 && k <= N - 1);
    for (int i = 0; i < QNX; ++i) {
      double v = restart ? QA(X, i, k) : C.in.xWS[(size_t)QNX * k + i];
      if (k == 0) v = C.in.x0[i];
      if (k == N) v = C.in.xF[i];
      if (free_x) v = push_lo(v, P.xlo[i], P.xhi[i], O.kappa1, O.kappa2);
      QA(X, i, k) = v; QA(ZXL, i, k) = 1.0; QA(ZXU, i, k) = 1.0; QA(PI, i, k) = 0.0;
    }
    for (int j = 0; j < QNU; ++j) {
      double v = restart ? QA(U, j, k) : QUAD_WH;                          // setvalue(u, w_H) (:202)
      QA(U, j, k) = k < N ? push_lo(v, 1.2, 7.8, O.kappa1, O.kappa2) : 0.0;
      QA(ZUL, j, k) = 1.0; QA(ZUU, j, k) = 1.0;
    }
    for (int o = 0; o < QNOB; ++o) {
      double lw[6];
      if (!restart) {
        if (QCTX_O(C).quad_dual_ws) {
          const double pos[3] = {QA(X, 0, k), QA(X, 1, k), QA(X, 2, k)};
          quad_dual_ws(pos, P.obs[o], lw);                                  // closed-form dual warm start
        } else {
          for (int r = 0; r < 6; ++r) lw[r] = 0.05;                         // setvalue(l, 0.05) (:204-208)
        }
      }
      for (int r = 0; r < 6; ++r) {
        QA(LAM, 6 * o + r, k) = dmax(restart ? QA(LAM, 6 * o + r, k) : lw[r], O.kappa1);
        QA(ZLAM, 6 * o + r, k) = 1.0;
      }
    }
    for (int o = 0; o < QNOB; ++o) {
      if (SDV) { QA(SLK, o, k) = dmax(restart ? QA(SLK, o, k) : 1.0, O.kappa1); QA(ZSLK, o, k) = 1.0; }   // slack = 1 (:210)
      QA(YN, o, k) = 0.0;
    }
  }
  OBCA_HD_NI static void init_slacks(const QCtx& C, int k) {
    QLOCALS(C);
    const QuadProblem& P = QCTX_P(C);
    double pos[3] = {QA(X, 0, k), QA(X, 1, k), QA(X, 2, k)};
    for (int o = 0; o < QNOB; ++o) {
      QObsVars Q; QObsGeom G;
      load_obs(C, k, o, Q, 0.0);
      Q.sd = 0.0;
      qobs_geom<SDV>(pos, Q, G);
      QA(SD, o, k) = dmax(G.gd, P.R + QCTX_O(C).kappa1 * dmax(1.0, dabs(P.R)));
      QA(VD, o, k) = 1.0;
    }
  }

  // ---- K1 ----
  OBCA_HD_NI static void stage_eval(const QCtx& C, int k, bool do_err, bool do_asm, EvalPart& out) {
    QLOCALS(C);
    const QuadProblem& P = QCTX_P(C);
    const ProbState& S = *C.S;
    const int N = P.N;
    const double mu_b = S.mu, dw = S.dw, t = S.t;
    const bool free_x = (k >= 1 && k <= N - 1);
    const bool has_u = k < N;
    double e_dual = 0.0, e_pr = 0.0, cmax = 0.0, cmin = 1e300, sum_y = 0.0, sum_z = 0.0, th = 0.0, phi = 0.0, fobj = 0.0, rz_t = 0.0;
    int ok = 1;
    LogAcc lacc;
    double x[QNX], rzx[QNX];
    for (int i = 0; i < QNX; ++i) { x[i] = QA(X, i, k); rzx[i] = 0.0; }
    if (do_asm && has_u) {
      for (int e = 0; e < QNQ; ++e) QA(QS, e, k) = 0.0;
      for (int e = 0; e < QNYV; ++e) QA(qs, e, k) = 0.0;
    }
#define QQ(i, j) QA(QS, sym_idx_any<QNYV>((i), (j)), k)
#define Qq(i) QA(qs, (i), k)
    // ---- state objective + bounds ----
    fobj += QUAD_REG3 * (x[9] * x[9] + x[10] * x[10] + x[11] * x[11]);
    if (free_x) {
      for (int i = 0; i < QNX; ++i) {
        const double g = i >= 9 ? 2.0 * QUAD_REG3 * x[i] : 0.0;
        const double al = x[i] - P.xlo[i], au = P.xhi[i] - x[i];
        const double ial = rcp(al), iau = rcp(au);
        const double zl = QA(ZXL, i, k), zu = QA(ZXU, i, k);
        if (do_asm) {
          QQ(i, i) += (i >= 9 ? 2.0 * QUAD_REG3 : 0.0) + zl * ial + zu * iau + dw;
          Qq(i) += g - mu_b * ial + mu_b * iau;
        }
        rzx[i] += g - zl + zu + QA(PI, i, k - 1);
        if (do_err) {
          cmax = dmax(cmax, dmax(al * zl, au * zu)); cmin = dmin_(cmin, dmin_(al * zl, au * zu));
          sum_z += zl + zu;
          lacc.add(al); lacc.add(au);
        }
      }
    }
    // ---- controls + dynamics ----
    if (has_u) {
      double u[QNU], pi[QNX], c0[3] = {C.in.x0[9], C.in.x0[10], C.in.x0[11]};
      for (int i = 0; i < QNX; ++i) pi[i] = QA(PI, i, k);
      double rzu[QNU];
      for (int j = 0; j < QNU; ++j) {
        u[j] = QA(U, j, k);
        const double w = k > 0 ? QA(U, j, k - 1) : 0.0;
        const double eh = QUAD_WH - u[j], ed = k > 0 ? (w - u[j]) : 0.0;
        fobj += 1e-3 * eh * eh + 1e-2 * ed * ed;
        const double gu = -2e-3 * eh - 2e-2 * ed, gw = 2e-2 * ed;
        const double al = u[j] - 1.2, au = 7.8 - u[j], ial = rcp(al), iau = rcp(au);
        const double zl = QA(ZUL, j, k), zu = QA(ZUU, j, k);
        if (do_asm) {
          QQ(QIU + j, QIU + j) += 2e-3 + (k > 0 ? 2e-2 : 0.0) + zl * ial + zu * iau + dw;
          if (k > 0) { QQ(QIW + j, QIW + j) += 2e-2; QQ(QIW + j, QIU + j) += -2e-2; }
          Qq(QIU + j) += gu - mu_b * ial + mu_b * iau;
          Qq(QIW + j) += gw;
        }
        rzu[j] = gu - zl + zu;
        if (k + 1 < N) rzu[j] += 2e-2 * (u[j] - QA(U, j, k + 1));      // role as "previous control" of stage k+1
        if (do_err) {
          cmax = dmax(cmax, dmax(al * zl, au * zu)); cmin = dmin_(cmin, dmin_(al * zl, au * zu));
          sum_z += zl + zu;
          lacc.add(al); lacc.add(au);
        }
      }
      double f[QNX], Jv[QD_NJ], Hv[QD_NH];
      quad_dyn_full(x, u, t, P.Ts, c0, pi, f, Jv, Hv);
      const int *jr, *jc, *hi, *hj;
      quad_jac_tables(jr, jc, hi, hj);
      if (do_asm) {
        for (int e = 0; e < QD_NH; ++e) QQ(hi[e], hj[e]) += Hv[e];
        for (int e = 0; e < QD_NJ; ++e) QA(JV, e, k) = Jv[e];
      }
      for (int i = 0; i < QNX; ++i) {
        const double xn = (k + 1 == N) ? C.in.xF[i] : QA(X, i, k + 1);
        const double r = f[i] - xn;
        if (do_asm) QA(R12, i, k) = r;
        if (do_err) { e_pr = dmax(e_pr, dabs(r)); th += dabs(r); sum_y += dabs(pi[i]); }
      }
      // Lagrangian gradient: - J' pi
      for (int e = 0; e < QD_NJ; ++e) {
        const double v = pi[jr[e]] * Jv[e];
        const int c = jc[e];
        if (c < QNX) rzx[c] -= v;
        else if (c == QIT) rz_t -= v;
        else rzu[c - QIU] -= v;
      }
      if (do_err)
        for (int j = 0; j < QNU; ++j) e_dual = dmax(e_dual, dabs(rzu[j]));
    }
    // ---- obstacle blocks ----
    double pos[3] = {x[0], x[1], x[2]};
    for (int o = 0; o < QNOB; ++o) {
      QObsVars Q; QObsGeom G;
      load_obs(C, k, o, Q, 0.0);
      qobs_geom<SDV>(pos, Q, G);
      if (do_err) {
        double rl[6], rs, gx[3];
        qobs_lagr_grad<SDV>(Q, G, rl, rs, gx);
        rzx[0] += gx[0]; rzx[1] += gx[1]; rzx[2] += gx[2];
        for (int i = 0; i < 6; ++i) {
          e_dual = dmax(e_dual, dabs(rl[i]));
          const double cp = Q.lam[i] * Q.zlam[i];
          cmax = dmax(cmax, cp); cmin = dmin_(cmin, cp);
          sum_z += Q.zlam[i];
          lacc.add(Q.lam[i]);
          fobj += QUAD_REG2 * Q.lam[i] * Q.lam[i];
        }
        if (SDV) {
          e_dual = dmax(e_dual, dabs(rs));
          const double cp = Q.sl * Q.zsl;
          cmax = dmax(cmax, cp); cmin = dmin_(cmin, cp);
          sum_z += Q.zsl;
          lacc.add(Q.sl);
          fobj += 1e2 * Q.sl + 1e3 * Q.sl * Q.sl;
        }
        e_pr = dmax(e_pr, dmax(dabs(G.cn), dabs(G.cd)));
        th += dabs(G.cn) + dabs(G.cd);
        sum_y += dabs(Q.yn) + Q.vd; sum_z += Q.vd;
        const double cp = (Q.sd - P.R) * Q.vd;
        cmax = dmax(cmax, cp); cmin = dmin_(cmin, cp);
        lacc.add(Q.sd - P.R);
      }
      if (do_asm) {
        const int piv = qobs_choose_pivot(G);
        if (piv != 0) { qobs_swap(Q, piv); qobs_geom<SDV>(pos, Q, G); }
        double Sxx[6], rx3[3];
        ok &= qobs_condense<SDV>(P.R, Q, G, mu_b, dw, QCTX_O(C).dc, Sxx, rx3, &QA(LF, o * QCTX_L(C).nfac, k), QCTX_L(C).NSP);
        if (free_x && has_u) {
          QQ(0, 0) += Sxx[0]; QQ(0, 1) += Sxx[1]; QQ(0, 2) += Sxx[2]; QQ(1, 1) += Sxx[3]; QQ(1, 2) += Sxx[4]; QQ(2, 2) += Sxx[5];
          Qq(0) += rx3[0]; Qq(1) += rx3[1]; Qq(2) += rx3[2];
        }
      }
    }
    // ---- time scale: objective (N+1)(0.25 t + 5 t^2) (:68) and its N+1 bound pairs ----
    if (k == 0) {
      const double m = (double)(N + 1);
      const double gl = t - 0.5, gu = 2.0 - t, igl = rcp(gl), igu = rcp(gu);
      fobj += m * (0.25 * t + 5.0 * t * t);
      if (do_asm) {
        QQ(QIT, QIT) += 10.0 * m + m * (S.zTL * igl + S.zTU * igu) + m * dw;   // N+1 copies of timeScale, each regularised
        Qq(QIT) += m * (0.25 + 10.0 * t) + m * (-mu_b * igl + mu_b * igu);
      }
      if (do_err) {
        rz_t += m * (0.25 + 10.0 * t) - m * (S.zTL - S.zTU);
        cmax = dmax(cmax, dmax(gl * S.zTL, gu * S.zTU)); cmin = dmin_(cmin, dmin_(gl * S.zTL, gu * S.zTU));
        sum_z += m * (S.zTL + S.zTU);
        phi -= m * mu_b * (log(gl) + log(gu));
      }
    }
    if (do_err) phi -= mu_b * lacc.total();
    if (do_err && free_x)
      for (int i = 0; i < QNX; ++i) e_dual = dmax(e_dual, dabs(rzx[i]));
    out.e_dual = e_dual; out.e_pr = e_pr; out.cmax = cmax; out.cmin = cmin; out.sy = sum_y; out.sz = sum_z;
    out.th = th; out.phi = phi + fobj; out.rt = rz_t; out.f = fobj; out.ok = ok;
#undef QQ
#undef Qq
  }

  // ---- K3: dense Riccati sweep over [x | w | t] (17) with 4 controls, run by one thread ----
  OBCA_HD static int kkt_dense(const QCtx& C) {
    QLOCALS(C);
    const QuadProblem& Pp = QCTX_P(C);
    ProbState& S = *C.S;
    const int N = Pp.N;
    const int *jr, *jc, *hi, *hj;
    quad_jac_tables(jr, jc, hi, hj);
    double P[QNSV][QNSV], p[QNSV];
    for (int a = 0; a < QNSV; ++a) { p[a] = 0.0; for (int b = 0; b < QNSV; ++b) P[a][b] = 0.0; }
    const double rho = 1.0 / QCTX_O(C).dc;
    for (int i = 0; i < QNX; ++i) { P[i][i] = rho; p[i] = -QA(PI, i, N - 1); }
    int ok = 1;
    for (int k = N - 1; k >= 0 && ok; --k) {
      // rows 0..11 of P_{k+1}, p_{k+1} for the multiplier recovery
      for (int i = 0; i < QNX; ++i) {
        for (int l = 0; l < QNSV; ++l) QA(RP, i * QNSV + l, k + 1) = P[i][l];
        QA(RP, QNX * QNSV + i, k + 1) = p[i];
      }
      double Jv[QD_NJ], r12[QNX];
      for (int e = 0; e < QD_NJ; ++e) Jv[e] = QA(JV, e, k);
      for (int i = 0; i < QNX; ++i) r12[i] = QA(R12, i, k);
      // g = p + P r~ ;  T = P Phi
      double g[QNSV], T[QNSV][QNYV];
      for (int a = 0; a < QNSV; ++a) {
        double acc = p[a];
        for (int l = 0; l < QNX; ++l) acc += P[a][l] * r12[l];
        g[a] = acc;
        for (int c = 0; c < QNYV; ++c) T[a][c] = 0.0;
        for (int e = 0; e < QD_NJ; ++e) T[a][jc[e]] += P[a][jr[e]] * Jv[e];
        for (int j = 0; j < QNU; ++j) T[a][QIU + j] += P[a][QIW + j];
        T[a][QIT] += P[a][QIT];
      }
      // H = Q + Phi' T (full), hv = q + Phi' g
      double H[QNYV][QNYV], hv[QNYV];
      for (int a = 0; a < QNYV; ++a) {
        hv[a] = QA(qs, a, k);
        for (int b = 0; b < QNYV; ++b) H[a][b] = QA(QS, sym_idx_any<QNYV>(a, b), k);
      }
      for (int e = 0; e < QD_NJ; ++e) {
        const int c = jc[e], r = jr[e];
        const double v = Jv[e];
        hv[c] += v * g[r];
        for (int b = 0; b < QNYV; ++b) H[c][b] += v * T[r][b];
      }
      for (int j = 0; j < QNU; ++j) {
        hv[QIU + j] += g[QIW + j];
        for (int b = 0; b < QNYV; ++b) H[QIU + j][b] += T[QIW + j][b];
      }
      hv[QIT] += g[QIT];
      for (int b = 0; b < QNYV; ++b) H[QIT][b] += T[QIT][b];
      // Cholesky of Huu (4x4), must be positive definite
      double Lc[QNU][QNU];
      for (int a = 0; a < QNU; ++a) {
        for (int b = 0; b <= a; ++b) {
          double acc = H[QIU + a][QIU + b];
          for (int l = 0; l < b; ++l) acc -= Lc[a][l] * Lc[b][l];
          if (a == b) {
            if (!(acc > 0.0)) { ok = 0; acc = 1e300; }
            Lc[a][a] = sqrt(acc);
          } else {
            Lc[a][b] = acc / Lc[b][b];
          }
        }
      }
      if (!ok) break;
      // K = -Huu^{-1} Hus (4x17), kf = -Huu^{-1} hu
      double K[QNU][QNSV + 1];
      for (int c = 0; c <= QNSV; ++c) {
        double y4[QNU];
        for (int a = 0; a < QNU; ++a) {
          double acc = c < QNSV ? H[QIU + a][c] : hv[QIU + a];
          for (int l = 0; l < a; ++l) acc -= Lc[a][l] * y4[l];
          y4[a] = acc / Lc[a][a];
        }
        for (int a = QNU - 1; a >= 0; --a) {
          double acc = y4[a];
          for (int l = a + 1; l < QNU; ++l) acc -= Lc[l][a] * K[l][c];
          K[a][c] = acc / Lc[a][a];
        }
        for (int a = 0; a < QNU; ++a) K[a][c] = -K[a][c];
      }
      for (int a = 0; a < QNU; ++a) {
        for (int c = 0; c < QNSV; ++c) QA(RK, a * QNSV + c, k) = K[a][c];
        QA(RK, QNU * QNSV + a, k) = K[a][QNSV];
      }
      for (int a = 0; a < QNSV; ++a) {
        double acc = hv[a];
        for (int l = 0; l < QNU; ++l) acc += H[a][QIU + l] * K[l][QNSV];
        p[a] = acc;
        for (int b = a; b < QNSV; ++b) {
          double v = H[a][b];
          for (int l = 0; l < QNU; ++l) v += H[a][QIU + l] * K[l][b];
          P[a][b] = v; P[b][a] = v;
        }
      }
    }
    if (!ok) return 0;
    double ptt = P[QIT][QIT];
    if (!(ptt > 0.0)) return 0;
    const double dt = -p[QIT] / ptt;
    S.dt = dt;
    // forward roll-out
    double s[QNSV];
    for (int a = 0; a < QNSV; ++a) s[a] = 0.0;
    s[QIT] = dt;
    for (int k = 0; k < N; ++k) {
      double u[QNU];
      for (int a = 0; a < QNU; ++a) {
        double acc = QA(RK, QNU * QNSV + a, k);
        for (int c = 0; c < QNSV; ++c) acc += QA(RK, a * QNSV + c, k) * s[c];
        u[a] = acc;
        QA(dU, a, k) = acc;
      }
      double sn[QNX];
      for (int i = 0; i < QNX; ++i) sn[i] = QA(R12, i, k);
      for (int e = 0; e < QD_NJ; ++e) {
        const int c = jc[e];
        const double yv = c < QNSV ? s[c] : u[c - QIU];
        sn[jr[e]] += QA(JV, e, k) * yv;
      }
      for (int i = 0; i < QNX; ++i) {
        if (k + 1 < N) QA(dX, i, k + 1) = sn[i];
        else S.eNq[i] = sn[i];
        s[i] = sn[i];
      }
      for (int a = 0; a < QNU; ++a) s[QIW + a] = u[a];
    }
    for (int i = 0; i < QNX; ++i) { QA(dX, i, 0) = 0.0; QA(dX, i, N) = 0.0; }
    for (int a = 0; a < QNU; ++a) QA(dU, a, N) = 0.0;
    return 1;
  }
  OBCA_HD static int kkt_host(const QCtx& C) { return kkt_dense(C); }
  static constexpr int STG_N = QNQ + QNYV + QD_NJ + QNX;     // per-stage data of the sweep: Q 231 | q 21 | Jv 66 | r 12
  static constexpr int SM_P = 0, SM_p = SM_P + QNSV * QNSV, SM_g = SM_p + QNSV, SM_T = SM_g + QNSV, SM_H = SM_T + QNSV * QNYV,
                       SM_hv = SM_H + QNYV * QNYV, SM_K = SM_hv + QNYV, SM_s = SM_K + QNU * (QNSV + 1), SM_u = SM_s + QNSV,
                       SM_flag = SM_u + QNU, SM_STG = SM_flag + 1 + ((SM_flag + 1) & 1), SM_Kall = SM_STG + 2 * STG_N;
  static constexpr int KROW = QNU * (QNSV + 1);     // 72 gain entries per stage
  OBCA_HD static int smem_doubles(int N) { return SM_Kall + N * KROW; }
  static constexpr bool KKT_BLOCK = true;

#if defined(__CUDA_ARCH__)
  // -------------------------------------------------------------------------------------------------
  // Device version: the whole CTA cooperates on every stage of the backward sweep through shared memory (entry-parallel
  // T = P Phi, H = Q + Phi' T, gains, value-function update), four barriers per stage; the per-stage data (Q, q, the
  // dynamics Jacobian values and residual: 330 doubles scattered over the [array][stage] workspace) is fetched one
  // stage ahead with 8-byte cp.async into a double buffer, so no global-memory latency sits inside a stage.  The 4x4
  // pivot block Huu is factored (LDL', no square roots) redundantly by each of the 18 threads that then solve one
  // column of the gain.  The forward roll-out is run by warp 0 alone (shuffle-free, two __syncwarp per stage).
  // Same elimination as kkt_dense(); P is computed for a <= b and mirrored.  Shared layout (doubles):
  //   P 17x17 | p 17 | g 17 | T 17x21 | H 21x21 | hv 21 | K 4x18 | s 17 | u 4 | flag | stage buffers 2 x 330 | Kall N x 72
  // -------------------------------------------------------------------------------------------------
  // ---- statically unrolled pieces of the sweep: Phi's sparsity (generated tables) is folded at compile time ----
  // T(a, c) for the columns c = G, G+4, ... of one row a of P (held in registers);  G = warp index
  template <int G>
  __device__ static __forceinline__ void sweep_T_cols(const double (&Pr)[QNSV], const double* Jv, double* Trow) {
    constexpr int JR[QD_NJ] = OBCA_QD_J_ROW;
    constexpr int CPT[QNYV + 1] = OBCA_QD_CSC_PTR;
    constexpr int CIX[QD_NJ] = OBCA_QD_CSC_IDX;
#pragma unroll
    for (int c = G; c < QNYV; c += 4) {
      double acc = 0.0;
#pragma unroll
      for (int q = CPT[c]; q < CPT[c + 1]; ++q) acc += Pr[JR[CIX[q]]] * Jv[CIX[q]];
      if (c >= QIU) acc += Pr[QIW + (c - QIU)];
      if (c == QIT) acc += Pr[QIT];
      Trow[c] = acc;
    }
  }
  // H(c1, c2) = Q(c1, c2) + Phi(:, c1)' T(:, c2) for the rows c1 = G, G+4, ... and one column c2 (T(:, c2) in registers);
  // c2 == QNYV stands for the gradient column: hv(c1) = q(c1) + Phi(:, c1)' g
  template <int G>
  __device__ static __forceinline__ void sweep_H_rows(const double (&Tc)[QNSV], int c2, const double* Qs, const double* qv,
                                                      const double* Jv, double* H, double* hv) {
    constexpr int JR[QD_NJ] = OBCA_QD_J_ROW;
    constexpr int CPT[QNYV + 1] = OBCA_QD_CSC_PTR;
    constexpr int CIX[QD_NJ] = OBCA_QD_CSC_IDX;
#pragma unroll
    for (int c1 = G; c1 < QNYV; c1 += 4) {
      double acc = (c2 < QNYV) ? Qs[c1 <= c2 ? sym_idx<QNYV>(c1, c2) : sym_idx<QNYV>(c2, c1)] : qv[c1];
#pragma unroll
      for (int q = CPT[c1]; q < CPT[c1 + 1]; ++q) acc += Jv[CIX[q]] * Tc[JR[CIX[q]]];
      if (c1 >= QIU) acc += Tc[QIW + (c1 - QIU)];
      if (c1 == QIT) acc += Tc[QIT];
      if (c2 < QNYV) H[c1 * QNYV + c2] = acc; else hv[c1] = acc;
    }
  }
#ifdef OBCA_QPROF
#define QPROF(i) do { if (threadIdx.x == 0) { long long t_ = clock64(); qacc[i] += t_ - qt; qt = t_; } } while (0)
#else
#define QPROF(i) do { } while (0)
#endif
  __device__ __noinline__ static int kkt_solve_block(const QCtx& C) {
    QLOCALS(C);
    extern __shared__ __align__(16) double obca_dyn_smem[];      // == C.tile; named here so that the accesses are LDS / STS
#ifdef OBCA_QPROF
    long long qacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, qt = clock64();
#endif
    const QuadProblem& Pp = QCTX_P(C);
    ProbState& S = *C.S;
    const int N = Pp.N;
    const int tid = threadIdx.x, nt = blockDim.x;
    double* const sm = obca_dyn_smem;
    double *P = sm + SM_P, *p = sm + SM_p, *g = sm + SM_g, *T = sm + SM_T, *H = sm + SM_H, *hv = sm + SM_hv, *K = sm + SM_K,
           *sv = sm + SM_s, *uv = sm + SM_u, *stg = sm + SM_STG, *Kall = sm + SM_Kall;
    static constexpr int JR[QD_NJ] = OBCA_QD_J_ROW;
    static constexpr int JC[QD_NJ] = OBCA_QD_J_COL;
    static constexpr int RPT[QNX + 1] = OBCA_QD_ROW_PTR;
    static constexpr int CPT[QNYV + 1] = OBCA_QD_CSC_PTR;
    static constexpr int CIX[QD_NJ] = OBCA_QD_CSC_IDX;
    const double rho = 1.0 / QCTX_O(C).dc;
    const size_t nsp = (size_t)QCTX_L(C).NSP;
    const double* const gQ = W_ + (size_t)QCTX_L(C).QS * nsp;     // QS, qs, JV, R12 are consecutive arrays of the workspace
    auto prefetch = [&](int k, int buf, int first, int count) {   // elements [first, first+count) of stage k
      if (k >= 0 && k < N)
        for (int e = tid; e < count; e += nt) cp_async8(stg + buf * STG_N + first + e, gQ + (size_t)(first + e) * nsp + k);
      cp_async_commit();
    };
    const int lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
    // the (at most two) entries a <= b of the value function this thread updates in every stage
    int pa[2] = {-1, -1}, pb[2] = {-1, -1};
    for (int i = 0; i < 2; ++i) {
      const int e = tid + i * nt;
      if (e < QNP) {
        int a = 0, off = e;
        while (off >= QNSV - a) { off -= QNSV - a; ++a; }
        pa[i] = a; pb[i] = a + off;
      }
    }
    for (int e = tid; e < QNSV * QNSV; e += nt) P[e] = (e / QNSV == e % QNSV && e / QNSV < QNX) ? rho : 0.0;
    for (int a = tid; a < QNSV; a += nt) p[a] = a < QNX ? -QA(PI, a, N - 1) : 0.0;
    if (tid == 0) sm[SM_flag] = 1.0;
    prefetch(N - 1, (N - 1) & 1, 0, STG_N);
    for (int k = N - 1; k >= 0; --k) {
      const double* const Qs = stg + (k & 1) * STG_N;
      const double* const qv = Qs + QNQ;
      const double* const Jv = qv + QNYV;
      const double* const r12 = Jv + QD_NJ;
      prefetch(k - 1, (k - 1) & 1, 0, STG_N);
      QPROF(0);
      cp_async_wait<1>();
      __syncthreads();                     // stage data landed; P, p of stage k+1 final
      QPROF(1);
      // rows 0..11 of P_{k+1}, p_{k+1} for the multiplier recovery (fire-and-forget stores)
      for (int e = tid; e < QNX * QNSV; e += nt) QA(RP, e, k + 1) = P[e];
      for (int i = tid; i < QNX; i += nt) QA(RP, QNX * QNSV + i, k + 1) = p[i];
      // g = p + P r~ ;  T = P Phi:  lane a < 17 of every warp holds row a of P; warp w computes the columns w, w+4, ...
      if (lane < QNSV) {
        double Pr[QNSV];
#pragma unroll
        for (int l = 0; l < QNSV; ++l) Pr[l] = P[lane * QNSV + l];
        for (int gi = warp; gi < 4; gi += nwarp) {
          if (gi == 0) {
            double acc = p[lane];
#pragma unroll
            for (int l = 0; l < QNX; ++l) acc += Pr[l] * r12[l];
            g[lane] = acc;
            sweep_T_cols<0>(Pr, Jv, T + lane * QNYV);
          } else if (gi == 1) sweep_T_cols<1>(Pr, Jv, T + lane * QNYV);
          else if (gi == 2) sweep_T_cols<2>(Pr, Jv, T + lane * QNYV);
          else sweep_T_cols<3>(Pr, Jv, T + lane * QNYV);
        }
      }
      __syncthreads();
      QPROF(2);
      // H = Q + Phi' T ; hv = q + Phi' g:  lane c2 <= 21 holds column c2 of T (21: g); warp w computes the rows w, w+4, ...
      if (lane <= QNYV) {
        double Tc[QNSV];
#pragma unroll
        for (int l = 0; l < QNSV; ++l) Tc[l] = (lane < QNYV) ? T[l * QNYV + lane] : g[l];
        for (int gi = warp; gi < 4; gi += nwarp) {
          if (gi == 0) sweep_H_rows<0>(Tc, lane, Qs, qv, Jv, H, hv);
          else if (gi == 1) sweep_H_rows<1>(Tc, lane, Qs, qv, Jv, H, hv);
          else if (gi == 2) sweep_H_rows<2>(Tc, lane, Qs, qv, Jv, H, hv);
          else sweep_H_rows<3>(Tc, lane, Qs, qv, Jv, H, hv);
        }
      }
      __syncthreads();
      QPROF(3);
      // K = -Huu^{-1} [Hus | hu]: column c = 0..17, one thread per column, each with its own LDL' of Huu
      if (tid <= QNSV) {
        const int c = tid;
        double Lm[QNU][QNU], dd[QNU], di[QNU];      // unit lower factor, pivots d and 1/d
        int ok = 1;
#pragma unroll
        for (int a = 0; a < QNU; ++a) {
#pragma unroll
          for (int b = 0; b < a; ++b) {
            double acc = H[(QIU + a) * QNYV + QIU + b];
#pragma unroll
            for (int l = 0; l < b; ++l) acc -= Lm[a][l] * Lm[b][l] * dd[l];
            Lm[a][b] = acc * di[b];
          }
          double d = H[(QIU + a) * QNYV + QIU + a];
#pragma unroll
          for (int l = 0; l < a; ++l) d -= Lm[a][l] * Lm[a][l] * dd[l];
          if (!(d > 0.0)) { ok = 0; d = 1e300; }
          dd[a] = d; di[a] = __drcp_rn(d);
        }
        if (!ok && tid == 0) sm[SM_flag] = 0.0;
        double y4[QNU], k4[QNU];
#pragma unroll
        for (int a = 0; a < QNU; ++a) {
          double acc = c < QNSV ? H[(QIU + a) * QNYV + c] : hv[QIU + a];
#pragma unroll
          for (int l = 0; l < a; ++l) acc -= Lm[a][l] * y4[l];
          y4[a] = acc;
        }
#pragma unroll
        for (int a = QNU - 1; a >= 0; --a) {
          double acc = y4[a] * di[a];
#pragma unroll
          for (int l = a + 1; l < QNU; ++l) acc -= Lm[l][a] * k4[l];
          k4[a] = acc;
        }
#pragma unroll
        for (int a = 0; a < QNU; ++a) { K[a * (QNSV + 1) + c] = -k4[a]; Kall[k * KROW + a * (QNSV + 1) + c] = -k4[a]; }
      }
      __syncthreads();
      QPROF(4);
      if (sm[SM_flag] == 0.0) { cp_async_wait<0>(); return 0; }
      // value function of stage k (a <= b, mirrored); the barrier at the top of the next stage publishes it
      for (int i = 0; i < 2; ++i) {
        const int a = pa[i], b = pb[i];
        if (a < 0) continue;
        double v = H[a * QNYV + b];
#pragma unroll
        for (int l = 0; l < QNU; ++l) v += H[a * QNYV + QIU + l] * K[l * (QNSV + 1) + b];
        P[a * QNSV + b] = v; P[b * QNSV + a] = v;
      }
      for (int a = tid; a < QNSV; a += nt) {
        double acc = hv[a];
#pragma unroll
        for (int l = 0; l < QNU; ++l) acc += H[a * QNYV + QIU + l] * K[l * (QNSV + 1) + QNSV];
        p[a] = acc;
      }
    }
    cp_async_wait<0>();
    __syncthreads();
    QPROF(5);
    // root
    if (tid == 0) {
      const double ptt = P[QIT * QNSV + QIT];
      if (!(ptt > 0.0)) sm[SM_flag] = 0.0;
      else S.dt = -p[QIT] / ptt;
    }
    __syncthreads();
    if (sm[SM_flag] == 0.0) return 0;
    // forward roll-out: warp 0 alone; Jv and r of the next stage are prefetched while the current one is applied
    if (tid < 32) {
      const int lane = tid;
      constexpr int FW0 = QNQ + QNYV, FWN = QD_NJ + QNX;       // Jv | r
      auto pf = [&](int k, int buf) {
        if (k < N)
          for (int e = lane; e < FWN; e += 32) cp_async8(stg + buf * STG_N + FW0 + e, gQ + (size_t)(FW0 + e) * nsp + k);
        cp_async_commit();
      };
      if (lane < QNSV) sv[lane] = (lane == QIT) ? S.dt : 0.0;
      pf(0, 0);
      __syncwarp();
      for (int k = 0; k < N; ++k) {
        pf(k + 1, (k + 1) & 1);
        if (lane < QNU) {
          const double* kr = Kall + k * KROW + lane * (QNSV + 1);
          double acc = kr[QNSV];
#pragma unroll
          for (int c = 0; c < QNSV; ++c) acc += kr[c] * sv[c];
          uv[lane] = acc;
          QA(dU, lane, k) = acc;
        }
        cp_async_wait<1>();
        __syncwarp();
        const double* const Jv = stg + (k & 1) * STG_N + FW0;
        const double* const r12 = Jv + QD_NJ;
        double snv = 0.0;
        if (lane < QNX) {
          snv = r12[lane];
          for (int q = RPT[lane]; q < RPT[lane + 1]; ++q) { const int c = JC[q]; snv += Jv[q] * (c < QNSV ? sv[c] : uv[c - QIU]); }
          if (k + 1 < N) QA(dX, lane, k + 1) = snv; else S.eNq[lane] = snv;
        }
        __syncwarp();
        if (lane < QNX) sv[lane] = snv;
        else if (lane < QNX + QNU) sv[lane] = uv[lane - QNX];
        __syncwarp();
      }
      cp_async_wait<0>();
      for (int i = lane; i < QNX; i += 32) { QA(dX, i, 0) = 0.0; QA(dX, i, N) = 0.0; }
      if (lane < QNU) QA(dU, lane, N) = 0.0;
    }
    __syncthreads();
#ifdef OBCA_QPROF
    QPROF(6);
    if (threadIdx.x == 0) { for (int i = 0; i < 7; ++i) atomicAdd(&g_qprof[i], (unsigned long long)qacc[i]); atomicAdd(&g_qprof[7], 1ull); }
#endif
    return 1;
  }
  __device__ static int kkt_solve_warp(const QCtx&, double*) { return 0; }   // unused (KKT_BLOCK)
#endif

  // ---- K4a ----
  OBCA_HD_NI static void recover_stage(const QCtx& C, int k, StepPart& out) {
    QLOCALS(C);
    const QuadProblem& P = QCTX_P(C);
    const ProbState& S = *C.S;
    const int N = P.N;
    const double mu_b = S.mu, tau = S.tau;
    const bool free_x = (k >= 1 && k <= N - 1);
    double apr = 1.0, adu = 1.0, dphi = 0.0;
    if (k < N) {
      double sn[QNSV];
      for (int i = 0; i < QNX; ++i) sn[i] = (k + 1 < N) ? QA(dX, i, k + 1) : S.eNq[i];
      for (int a = 0; a < QNU; ++a) sn[QIW + a] = QA(dU, a, k);
      sn[QIT] = S.dt;
      for (int i = 0; i < QNX; ++i) {
        double acc = QA(RP, QNX * QNSV + i, k + 1);
        for (int l = 0; l < QNSV; ++l) acc += QA(RP, i * QNSV + l, k + 1) * sn[l];
        QA(PIn, i, k) = -acc;
      }
      for (int j = 0; j < QNU; ++j) {
        const double u = QA(U, j, k), du = QA(dU, j, k);
        const double w = k > 0 ? QA(U, j, k - 1) : 0.0, dwv = k > 0 ? QA(dU, j, k - 1) : 0.0;
        const double al = u - 1.2, au = 7.8 - u;
        ftb(al, du, tau, apr); ftb(au, -du, tau, apr);
        double z = QA(ZUL, j, k); ftb(z, dzb(z, al, du, mu_b), tau, adu);
        z = QA(ZUU, j, k); ftb(z, dzb(z, au, -du, mu_b), tau, adu);
        const double eh = QUAD_WH - u, ed = k > 0 ? (w - u) : 0.0;
        dphi += (-2e-3 * eh - 2e-2 * ed - mu_b * rcp(al) + mu_b * rcp(au)) * du + 2e-2 * ed * dwv;
      }
    }
    double x[QNX], dx[QNX];
    for (int i = 0; i < QNX; ++i) { x[i] = QA(X, i, k); dx[i] = QA(dX, i, k); }
    if (free_x) {
      for (int i = 0; i < QNX; ++i) {
        const double al = x[i] - P.xlo[i], au = P.xhi[i] - x[i];
        ftb(al, dx[i], tau, apr); ftb(au, -dx[i], tau, apr);
        double z = QA(ZXL, i, k); ftb(z, dzb(z, al, dx[i], mu_b), tau, adu);
        z = QA(ZXU, i, k); ftb(z, dzb(z, au, -dx[i], mu_b), tau, adu);
        dphi += ((i >= 9 ? 2.0 * QUAD_REG3 * x[i] : 0.0) - mu_b * rcp(al) + mu_b * rcp(au)) * dx[i];
      }
    }
    double pos[3] = {x[0], x[1], x[2]}, dpos[3] = {dx[0], dx[1], dx[2]};
    for (int o = 0; o < QNOB; ++o) {
      QObsVars Q; QObsGeom G;
      load_obs(C, k, o, Q, 0.0);
      qobs_geom<SDV>(pos, Q, G);
      const int piv = qobs_choose_pivot(G);
      if (piv != 0) { qobs_swap(Q, piv); qobs_geom<SDV>(pos, Q, G); }
      QObsStep St;
      qobs_recover<SDV>(P.R, Q, G, mu_b, &QA(LF, o * QCTX_L(C).nfac, k), QCTX_L(C).NSP, dpos, St);
      if (piv != 0) {
        for (int i = 1; i < 6; ++i)
          if (i == piv) {
            double tmp = St.dlam[0]; St.dlam[0] = St.dlam[i]; St.dlam[i] = tmp;
            tmp = Q.lam[0]; Q.lam[0] = Q.lam[i]; Q.lam[i] = tmp;
            tmp = Q.zlam[0]; Q.zlam[0] = Q.zlam[i]; Q.zlam[i] = tmp;
          }
      }
      for (int i = 0; i < 6; ++i) {
        QA(dLAM, 6 * o + i, k) = St.dlam[i];
        ftb(Q.lam[i], St.dlam[i], tau, apr);
        ftb(Q.zlam[i], dzb(Q.zlam[i], Q.lam[i], St.dlam[i], mu_b), tau, adu);
        dphi += (2.0 * QUAD_REG2 * Q.lam[i] - mu_b * rcp(Q.lam[i])) * St.dlam[i];
      }
      if (SDV) {
        QA(dSLK, o, k) = St.dsl;
        ftb(Q.sl, St.dsl, tau, apr);
        ftb(Q.zsl, dzb(Q.zsl, Q.sl, St.dsl, mu_b), tau, adu);
        dphi += (1e2 + 2e3 * Q.sl - mu_b * rcp(Q.sl)) * St.dsl;
      }
      QA(YNn, o, k) = St.yn_new;
      QA(dSD, o, k) = St.dsd;
      const double gap = Q.sd - P.R;
      ftb(gap, St.dsd, tau, apr);
      ftb(Q.vd, dzb(Q.vd, gap, St.dsd, mu_b), tau, adu);
      dphi += -mu_b * rcp(gap) * St.dsd;
    }
    if (k == 0) {
      const double m = (double)(N + 1), t = S.t;
      const double gl = t - 0.5, gu = 2.0 - t;
      ftb(gl, S.dt, tau, apr); ftb(gu, -S.dt, tau, apr);
      ftb(S.zTL, dzb(S.zTL, gl, S.dt, mu_b), tau, adu);
      ftb(S.zTU, dzb(S.zTU, gu, -S.dt, mu_b), tau, adu);
      dphi += m * (0.25 + 10.0 * t - mu_b * rcp(gl) + mu_b * rcp(gu)) * S.dt;
    }
    out.apr = apr; out.adu = adu; out.dphi = dphi;
  }

  // ---- K4b ----
  OBCA_HD_NI static void merit_stage(const QCtx& C, int k, double alpha, MeritPart& out) {
    QLOCALS(C);
    const QuadProblem& P = QCTX_P(C);
    const ProbState& S = *C.S;
    const int N = P.N;
    const double mu_b = S.mu;
    const bool free_x = (k >= 1 && k <= N - 1);
    const double t = S.t + alpha * S.dt;
    double th = 0.0, phi = 0.0;
    bool bad = false;
    LogAcc lacc;
    double x[QNX];
    for (int i = 0; i < QNX; ++i) x[i] = QA(X, i, k) + alpha * QA(dX, i, k);
    phi += QUAD_REG3 * (x[9] * x[9] + x[10] * x[10] + x[11] * x[11]);
    if (free_x)
      for (int i = 0; i < QNX; ++i) {
        const double al = x[i] - P.xlo[i], au = P.xhi[i] - x[i];
        bad |= !(al > 0.0) || !(au > 0.0);
        lacc.add(al); lacc.add(au);
      }
    if (k < N) {
      double u[QNU], c0[3] = {C.in.x0[9], C.in.x0[10], C.in.x0[11]};
      for (int j = 0; j < QNU; ++j) {
        u[j] = QA(U, j, k) + alpha * QA(dU, j, k);
        const double w = k > 0 ? QA(U, j, k - 1) + alpha * QA(dU, j, k - 1) : 0.0;
        const double eh = QUAD_WH - u[j], ed = k > 0 ? (w - u[j]) : 0.0;
        phi += 1e-3 * eh * eh + 1e-2 * ed * ed;
        const double al = u[j] - 1.2, au = 7.8 - u[j];
        bad |= !(al > 0.0) || !(au > 0.0);
        lacc.add(al); lacc.add(au);
      }
      double f[QNX];
      quad_dyn_f(x, u, t, P.Ts, c0, f);
      for (int i = 0; i < QNX; ++i) {
        const double xn = (k + 1 == N) ? C.in.xF[i] : QA(X, i, k + 1) + alpha * QA(dX, i, k + 1);
        th += dabs(f[i] - xn);
      }
    }
    double pos[3] = {x[0], x[1], x[2]};
    for (int o = 0; o < QNOB; ++o) {
      QObsVars Q; QObsGeom G;
      load_obs(C, k, o, Q, alpha);
      qobs_geom<SDV>(pos, Q, G);
      th += dabs(G.cn) + dabs(G.cd);
      for (int i = 0; i < 6; ++i) {
        bad |= !(Q.lam[i] > 0.0);
        phi += QUAD_REG2 * Q.lam[i] * Q.lam[i]; lacc.add(Q.lam[i]);
      }
      if (SDV) { bad |= !(Q.sl > 0.0); phi += 1e2 * Q.sl + 1e3 * Q.sl * Q.sl; lacc.add(Q.sl); }
      const double gap = Q.sd - P.R;
      bad |= !(gap > 0.0);
      lacc.add(gap);
    }
    if (k == 0) {
      const double m = (double)(N + 1);
      const double gl = t - 0.5, gu = 2.0 - t;
      bad |= !(gl > 0.0) || !(gu > 0.0);
      phi += m * (0.25 * t + 5.0 * t * t) - m * mu_b * (log(gl) + log(gu));
    }
    phi -= mu_b * lacc.total();
    out.th = th;
    out.phi = bad ? 1e300 : phi;
  }

  // ---- K4c ----
  OBCA_HD_NI static void update_stage(const QCtx& C, int k) {
    QLOCALS(C);
    const QuadProblem& P = QCTX_P(C);
    const ProbState& S = *C.S;
    const int N = P.N;
    const double mu_b = S.mu, ks = QCTX_O(C).kappa_sigma;
    const double alpha = S.alpha, adu = S.a_du, ay = dmin_(S.alpha, S.a_du);
    const bool free_x = (k >= 1 && k <= N - 1);
    if (free_x)
      for (int i = 0; i < QNX; ++i) {
        double q = QA(X, i, k), zl = QA(ZXL, i, k), zu = QA(ZXU, i, k);
        upd_pair(q, QA(dX, i, k), zl, zu, P.xlo[i], P.xhi[i], alpha, adu, mu_b, ks);
        QA(X, i, k) = q; QA(ZXL, i, k) = zl; QA(ZXU, i, k) = zu;
      }
    if (k < N) {
      for (int j = 0; j < QNU; ++j) {
        double q = QA(U, j, k), zl = QA(ZUL, j, k), zu = QA(ZUU, j, k);
        upd_pair(q, QA(dU, j, k), zl, zu, 1.2, 7.8, alpha, adu, mu_b, ks);
        QA(U, j, k) = q; QA(ZUL, j, k) = zl; QA(ZUU, j, k) = zu;
      }
      for (int i = 0; i < QNX; ++i) QA(PI, i, k) += ay * (QA(PIn, i, k) - QA(PI, i, k));
    }
    for (int r = 0; r < 6 * QNOB; ++r) {
      double q = QA(LAM, r, k), z = QA(ZLAM, r, k);
      const double dq = QA(dLAM, r, k);
      const double dz = dzb(z, q, dq, mu_b);
      q += alpha * dq; z += adu * dz;
      QA(LAM, r, k) = q; QA(ZLAM, r, k) = clipz(z, q, mu_b, ks);
    }
    for (int o = 0; o < QNOB; ++o) {
      if (SDV) {
        double q = QA(SLK, o, k), z = QA(ZSLK, o, k);
        const double dq = QA(dSLK, o, k);
        const double dz = dzb(z, q, dq, mu_b);
        q += alpha * dq; z += adu * dz;
        QA(SLK, o, k) = q; QA(ZSLK, o, k) = clipz(z, q, mu_b, ks);
      }
      QA(YN, o, k) += ay * (QA(YNn, o, k) - QA(YN, o, k));
      double s = QA(SD, o, k), z = QA(VD, o, k);
      const double ds = QA(dSD, o, k);
      const double dz = dzb(z, s - P.R, ds, mu_b);
      s += alpha * ds; z += adu * dz;
      QA(SD, o, k) = s; QA(VD, o, k) = clipz(z, s - P.R, mu_b, ks);
    }
  }

  // phase wrappers of the generic driver (the quadcopter model has no separate item pass: its 5 box blocks are evaluated
  // inside the stage functions)
  OBCA_HD static void eval_phase(const QCtx& C, bool do_err, EvalPart& ep) {
    const int NS = QCTX_P(C).N + 1;
    OBCA_FOR_STAGES(k, NS) { EvalPart e1; stage_eval(C, k, do_err, true, e1); part_merge(ep, e1); }
  }
  OBCA_HD static void recover_phase(const QCtx& C, StepPart& sp) {
    const int NS = QCTX_P(C).N + 1;
    OBCA_FOR_STAGES(k, NS) { StepPart s1; recover_stage(C, k, s1); part_merge(sp, s1); }
  }
  OBCA_HD static void merit_phase(const QCtx& C, double alpha, MeritPart& mp) {
    const int NS = QCTX_P(C).N + 1;
    OBCA_FOR_STAGES(k, NS) { MeritPart m1; merit_stage(C, k, alpha, m1); part_merge(mp, m1); }
  }
  OBCA_HD static void update_phase(const QCtx& C) {
    const int NS = QCTX_P(C).N + 1;
    OBCA_FOR_STAGES(k, NS) update_stage(C, k);
  }

  OBCA_HD static void store_stage(const QCtx& C, int k, const QOutputs& o) {
    QLOCALS(C);
    const int N = QCTX_P(C).N;
    for (int i = 0; i < QNX; ++i) o.xp[(size_t)QNX * k + i] = QA(X, i, k);
    if (k < N) for (int j = 0; j < QNU; ++j) o.up[(size_t)QNU * k + j] = QA(U, j, k);
    o.ts[k] = C.S->t;
    for (int r = 0; r < 6 * QNOB; ++r) o.lp[(size_t)6 * QNOB * k + r] = QA(LAM, r, k);      // lp = [l1; ...; l5] (:296)
    if (SDV && o.slack) for (int j = 0; j < QNOB; ++j) o.slack[(size_t)QNOB * k + j] = QA(SLK, j, k);
  }
};


// ------------------------------------------------------------------------------------------------------------
// K5 (quadcopter): twin of QuadcopterNavigation/constrSatisfaction.jl:25-204, restated verbatim: tolerance 1e-3,
// rows checked for i = 1..N only (:80), state box of the Dist variant ([-1.5, 3] on x10, :107,:119) for both
// variants, norm row one-sided (<= 0, :166-176), quirk Q5 in the body-rate rows (:151-153).
// x 12x(N+1), u 4xN, ts (N+1), lambda 30x(N+1) column-major.  Returns the worst violation per stage in `v`
// (<= 1e-3 everywhere  <=>  the reference returns true).
// ------------------------------------------------------------------------------------------------------------
OBCA_HD double quad_check_stage(const QuadProblem& P, int k, const double* x0, const double* xF, const double* x,
                                const double* u, const double* ts, const double* lam) {
  const int N = P.N;
  double worst = 0.0;
  if (k == 0)
    for (int i = 0; i < QNX; ++i) worst = dmax(worst, dabs(x[i] - x0[i]));                          // :56-61
  if (k == N) {
    for (int i = 0; i < QNX; ++i) worst = dmax(worst, dabs(x[(size_t)QNX * N + i] - xF[i]));       // :63-68
    for (int r = 0; r < 30; ++r) worst = dmax(worst, -lam[(size_t)30 * N + r]);                     // :162 looks at ALL columns
    return worst;
  }
  const double* xk = x + (size_t)QNX * k;
  const double* uk = u + (size_t)QNU * k;
  const double lo[QNX] = {0, 0, 0, -3, -0.2, -0.2, -1, -1, -1, -1.5, -1, -1};                       // :107
  const double hi[QNX] = {10, 10, 5, 3, 0.2, 0.2, 1, 1, 1, 3, 1, 1};                                // :119
  // the box rows are hard (> 0 fails, :84,:97,:108,:120): scale them so that "any violation" exceeds 1e-3
  for (int j = 0; j < QNU; ++j) {
    if (1.2 - uk[j] > 0.0 || uk[j] - 7.8 > 0.0) worst = dmax(worst, 1.0);
  }
  for (int i = 0; i < QNX; ++i)
    if (lo[i] - xk[i] > 0.0 || xk[i] - hi[i] > 0.0) worst = dmax(worst, 1.0);
  double f[QNX];
  const double c0[3] = {x[9], x[10], x[11]};                                                         // linear indexing, :151-153
  quad_dyn_f(xk, uk, ts[k], P.Ts, c0, f);
  for (int i = 0; i < QNX; ++i) worst = dmax(worst, dabs(x[(size_t)QNX * (k + 1) + i] - f[i]));      // :127-153
  worst = dmax(worst, dabs(ts[k] - ts[k + 1]));                                                      // :154
  for (int o = 0; o < QNOB; ++o) {
    const double* l = lam + (size_t)30 * k + 6 * o;
    double bl = 0.0;
    for (int r = 0; r < 6; ++r) { bl += P.obs[o][r] * l[r]; worst = dmax(worst, -l[r]); }            // :162 (lambda >= -1e-3)
    const double p0 = l[0] - l[3], p1 = l[1] - l[4], p2 = l[2] - l[5];
    worst = dmax(worst, p0 * p0 + p1 * p1 + p2 * p2 - 1.0);                                          // :171-180 (one-sided)
    worst = dmax(worst, -(-bl + xk[0] * p0 + xk[1] * p1 + xk[2] * p2 - P.R));                        // :187-200
  }
  return worst;
}

}  // namespace obca

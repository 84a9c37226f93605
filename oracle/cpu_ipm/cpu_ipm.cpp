// oracle/cpu_ipm/cpu_ipm.cpp -- ORACLE / CPU BASELINE (test infrastructure only: never linked into libobca.so, never a fallback).
//
// Compiled stand-in for what the reference does on the CPU at ParkingSignedDist.jl:240 `solve(m)`: a GENERIC sparse
// interior-point solve of the reference-formulation NLP (one row per @NLconstraint, JuMP variable order;
// oracle/parking_nlp.py) -- "IPOPT stand-in, not IPOPT" (BASELINE.md section 4.2):
//   * derivatives: the sympy templates of the oracle, printed as C by oracle/cpu_ipm/gen.py (templates_gen.h)
//     = the AD callbacks JuMP hands to Ipopt (eval_f / grad_f / g / jac_g / h);
//   * algorithm: the published Ipopt algorithm (Waechter & Biegler 2006) exactly as restated in oracle/ipm_ref.py (monotone
//     barrier update, inertia correction by delta_w escalation, filter line search, alpha_for_y = min, the options of
//     ParkingSignedDist.jl:41-43), same "barrier kick" restoration substitute;
//   * linear algebra: generic sparse symmetric-indefinite LDL' of the full augmented system [[W + Sigma + dw I, J'], [J, -dc]]
//     in skyline (envelope) storage with a stage-interleaved ordering, inertia from the signs of D -- the role MUMPS plays
//     for Ipopt.  No problem-specific condensation, no Riccati recursion: nothing of the CUDA solver's structure is used.
// One problem per OpenMP thread.  C-ABI at the bottom (ctypes: oracle/cpu_ipm/__init__.py).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <vector>

#include "templates_gen.h"

namespace {

struct Fam {
  int kind;        // 0 objective, 1 equality, 2 inequality
  int tmpl, n;
  const int* idx;  // n x nv
  int par_off;     // offset of this family's parameter block (n x np doubles) in the per-problem parameter blob
  int row0;
  const double *lo, *hi;   // inequality bounds (n) or null
};

struct Opts {   // oracle/ipm_ref.py IpmOptions (the Ipopt options of ParkingSignedDist.jl:41-43 + Ipopt defaults)
  double tol; int max_iter; double mu_init, mu_min_factor, kappa_eps, kappa_mu, theta_mu, tau_min, kappa1, kappa2, kappa_sigma, s_max;
  double dual_inf_tol, constr_viol_tol, compl_inf_tol, dw_min, dw_first, dw_max, kw_minus, kw_plus, kw_plus_first;
  double gamma_theta, gamma_phi, delta, s_theta, s_phi, eta_phi, gamma_alpha; int max_backtrack; double dc_value; int max_kick;
  double freeze_degenerate;   // > 0: an always-regularised row whose Jacobian row is below this threshold keeps its multiplier this iteration
};

struct Model {
  int n, mE, mI, nf;
  std::vector<Fam> fam;
  const double *zL, *zU;
  const unsigned char* dc_rows;   // mE
  const int* order;               // permutation of the n + mE unknowns: position p holds unknown order[p]
  // derived: symbolic skyline structure
  std::vector<int> pos;           // unknown -> position
  std::vector<int> first;         // first column of every row's envelope
  std::vector<long long> rowp;    // start of row i in the envelope array (entries first[i] .. i)
  long long nenv;
};

const double INF = 1e300;
inline bool fin(double x) { return x > -1e299 && x < 1e299; }

void mark(Model& M, int a, int b) {
  int pa = M.pos[a], pb = M.pos[b];
  if (pa < pb) std::swap(pa, pb);
  if (pb < M.first[pa]) M.first[pa] = pb;
}

void build_structure(Model& M) {
  const int NT = M.n + M.mE;
  M.pos.assign(NT, 0);
  for (int p = 0; p < NT; ++p) M.pos[M.order[p]] = p;
  M.first.resize(NT);
  for (int p = 0; p < NT; ++p) M.first[p] = p;
  for (const Fam& f : M.fam) {
    const TemplateInfo& T = TEMPLATES[f.tmpl];
    for (int r = 0; r < f.n; ++r) {
      const int* ix = f.idx + (size_t)r * T.nv;
      if (f.kind == 1) {
        for (int q = 0; q < T.ngnz; ++q) mark(M, M.n + f.row0 + r, ix[T.gnz[q]]);
      } else if (f.kind == 2) {      // J_I' Sigma_s J_I: dense on the row's support
        for (int a = 0; a < T.ngnz; ++a)
          for (int b = 0; b <= a; ++b) mark(M, ix[T.gnz[a]], ix[T.gnz[b]]);
      }
      for (int q = 0; q < T.nh; ++q) mark(M, ix[T.hi[q]], ix[T.hj[q]]);
    }
  }
  // a skyline factor fills the whole envelope; the envelope of row i must reach back to the smallest `first` of the rows
  // it couples with (standard: env is closed under the LDL' recursion when first[] is taken per row as computed)
  M.rowp.resize(NT + 1);
  long long c = 0;
  for (int i = 0; i < NT; ++i) { M.rowp[i] = c; c += i - M.first[i] + 1; }
  M.rowp[NT] = c; M.nenv = c;
}

struct Work {
  std::vector<double> K, D;            // envelope (row-wise, entries first[i]..i), pivots
  std::vector<double> gf, cE, gI, W_h; // scratch
};

inline double& KE(const Model& M, std::vector<double>& K, int i, int j) { return K[M.rowp[i] + (j - M.first[i])]; }   // i >= j >= first[i]

struct Eval {
  const Model& M; const double* par;
  Eval(const Model& m, const double* p) : M(m), par(p) {}
  double f(const double* z) const {
    double s = 0, v[64], g[64], h[1024], fv;
    for (const Fam& fm : M.fam) if (fm.kind == 0) {
      const TemplateInfo& T = TEMPLATES[fm.tmpl];
      for (int r = 0; r < fm.n; ++r) {
        const int* ix = fm.idx + (size_t)r * T.nv;
        for (int a = 0; a < T.nv; ++a) v[a] = z[ix[a]];
        T.fn(v, par + fm.par_off + (size_t)r * T.np, &fv, g, h);
        s += fv;
      }
    }
    return s;
  }
  void cons(const double* z, double* cE, double* gI) const {
    double v[64], g[64], h[1024], fv;
    for (const Fam& fm : M.fam) if (fm.kind != 0) {
      const TemplateInfo& T = TEMPLATES[fm.tmpl];
      for (int r = 0; r < fm.n; ++r) {
        const int* ix = fm.idx + (size_t)r * T.nv;
        for (int a = 0; a < T.nv; ++a) v[a] = z[ix[a]];
        T.fn(v, par + fm.par_off + (size_t)r * T.np, &fv, g, h);
        (fm.kind == 1 ? cE : gI)[fm.row0 + r] = fv;
      }
    }
  }
};

struct Ipm {
  const Model& M; const Opts& O; const double* par;
  int n, mE, mI, NT;
  std::vector<double> z, s, yE, zLm, zUm, vL, vU, gL, gU, dc;
  std::vector<unsigned char> hasL, hasU, sL, sU;
  // per-evaluation
  std::vector<double> gf, cE, gI, rz, bz, gam, Sz, Ss, rhs, sol, K, D, dz, ds, dzL, dzU, dvL, dvU, yEn, zt, st, cEt, gIt, tmp;
  double n_mult, n_bmult;
  int status, iters; double err;
  Ipm(const Model& m, const Opts& o, const double* p) : M(m), O(o), par(p) {
    n = M.n; mE = M.mE; mI = M.mI; NT = n + mE;
    z.resize(n); s.resize(mI); yE.assign(mE, 0.0); zLm.resize(n); zUm.resize(n); vL.resize(mI); vU.resize(mI); gL.resize(mI); gU.resize(mI);
    hasL.resize(n); hasU.resize(n); sL.resize(mI); sU.resize(mI); dc.assign(mE, 0.0);
    for (int i = 0; i < mE; ++i) if (M.dc_rows[i]) dc[i] = O.dc_value;
    for (const Fam& f : M.fam) if (f.kind == 2)
      for (int r = 0; r < f.n; ++r) { gL[f.row0 + r] = f.lo ? f.lo[r] : -INF; gU[f.row0 + r] = f.hi ? f.hi[r] : INF; }
    gf.resize(n); cE.resize(mE); gI.resize(mI); rz.resize(n); bz.resize(n); gam.resize(mI); Sz.resize(n); Ss.resize(mI); rhs.resize(NT);
    sol.resize(NT); K.resize(M.nenv); D.resize(NT); dz.resize(n); ds.resize(mI); dzL.resize(n); dzU.resize(n); dvL.resize(mI); dvU.resize(mI);
    yEn.resize(mE); zt.resize(n); st.resize(mI); cEt.resize(mE); gIt.resize(mI); tmp.resize(NT);
  }
  static double push(double x, double lo, double hi, double k1, double k2) {
    const bool fl = fin(lo), fu = fin(hi);
    double pl = k1 * std::max(1.0, fabs(lo)), pu = k1 * std::max(1.0, fabs(hi));
    if (fl && fu) { pl = std::min(pl, k2 * (hi - lo)); pu = std::min(pu, k2 * (hi - lo)); }
    if (fl) x = std::max(x, lo + pl);
    if (fu) x = std::min(x, hi - pu);
    return x;
  }
  // gradient of f, Jacobian products and residual of the Lagrangian gradient with multipliers (yE, yI)
  void grad_lag(const double* zz, const double* yE_, const double* yI_, double* gf_out, double* rz_out) {
    double v[64], g[64], h[1024], fv;
    for (int i = 0; i < n; ++i) { gf_out[i] = 0.0; rz_out[i] = 0.0; }
    for (const Fam& fm : M.fam) {
      const TemplateInfo& T = TEMPLATES[fm.tmpl];
      for (int r = 0; r < fm.n; ++r) {
        const int* ix = fm.idx + (size_t)r * T.nv;
        for (int a = 0; a < T.nv; ++a) v[a] = zz[ix[a]];
        T.fn(v, par + fm.par_off + (size_t)r * T.np, &fv, g, h);
        if (fm.kind == 0) { for (int q = 0; q < T.ngnz; ++q) gf_out[ix[T.gnz[q]]] += g[q]; }
        else {
          const double y = fm.kind == 1 ? yE_[fm.row0 + r] : yI_[fm.row0 + r];
          for (int q = 0; q < T.ngnz; ++q) rz_out[ix[T.gnz[q]]] += y * g[q];
        }
      }
    }
    for (int i = 0; i < n; ++i) rz_out[i] += gf_out[i];
  }
  double theta(const double* zz, const double* ss) {
    Eval(M, par).cons(zz, cEt.data(), gIt.data());
    double t = 0;
    for (int i = 0; i < mE; ++i) t += fabs(cEt[i]);
    for (int i = 0; i < mI; ++i) t += fabs(gIt[i] - ss[i]);
    return t;
  }
  double phi(const double* zz, const double* ss, double mu) {
    double lg = 0;
    for (int i = 0; i < n; ++i) {
      if (hasL[i]) { const double a = zz[i] - M.zL[i]; if (!(a > 0)) return INFINITY; lg += log(a); }
      if (hasU[i]) { const double a = M.zU[i] - zz[i]; if (!(a > 0)) return INFINITY; lg += log(a); }
    }
    for (int i = 0; i < mI; ++i) {
      if (sL[i]) { const double a = ss[i] - gL[i]; if (!(a > 0)) return INFINITY; lg += log(a); }
      if (sU[i]) { const double a = gU[i] - ss[i]; if (!(a > 0)) return INFINITY; lg += log(a); }
    }
    return Eval(M, par).f(zz) - mu * lg;
  }
  // Ipopt's scaled optimality error E_mu; also returns the three unscaled parts
  double errors(double mu_t, double* dinf_o, double* cinf_o, double* pinf_o) {
    std::vector<double>& yI = tmp;
    for (int i = 0; i < mI; ++i) yI[i] = vU[i] - vL[i];
    grad_lag(z.data(), yE.data(), yI.data(), gf.data(), rz.data());
    Eval(M, par).cons(z.data(), cE.data(), gI.data());
    double dinf = 0, cinf = 0, pinf = 0, ysum = 0, zsum = 0;
    for (int i = 0; i < n; ++i) {
      const double r = rz[i] - zLm[i] + zUm[i];
      dinf = std::max(dinf, fabs(r));
      if (hasL[i]) { pinf = std::max(pinf, fabs((z[i] - M.zL[i]) * zLm[i] - mu_t)); zsum += zLm[i]; }
      if (hasU[i]) { pinf = std::max(pinf, fabs((M.zU[i] - z[i]) * zUm[i] - mu_t)); zsum += zUm[i]; }
    }
    for (int i = 0; i < mE; ++i) { cinf = std::max(cinf, fabs(cE[i])); ysum += fabs(yE[i]); }
    for (int i = 0; i < mI; ++i) {
      cinf = std::max(cinf, fabs(gI[i] - s[i])); ysum += fabs(yI[i]);
      if (sL[i]) { pinf = std::max(pinf, fabs((s[i] - gL[i]) * vL[i] - mu_t)); zsum += vL[i]; }
      if (sU[i]) { pinf = std::max(pinf, fabs((gU[i] - s[i]) * vU[i] - mu_t)); zsum += vU[i]; }
    }
    const double sd = std::max(O.s_max, (ysum + zsum) / std::max(n_mult, 1.0)) / O.s_max;
    const double sc = std::max(O.s_max, zsum / std::max(n_bmult, 1.0)) / O.s_max;
    if (dinf_o) { *dinf_o = dinf; *cinf_o = cinf; *pinf_o = pinf; }
    return std::max(std::max(dinf / sd, cinf), pinf / sc);
  }
  // assemble K (envelope, permuted) = [[W + Sz + JI' Ss JI + dw I, JE'], [JE, -dc]] at (z, yE, yI); rhs pieces need JI' w products
  void assemble(double dw, const double* yI) {
    std::fill(K.begin(), K.end(), 0.0);
    double v[64], g[64], h[1024], fv;
    for (const Fam& fm : M.fam) {
      const TemplateInfo& T = TEMPLATES[fm.tmpl];
      for (int r = 0; r < fm.n; ++r) {
        const int* ix = fm.idx + (size_t)r * T.nv;
        for (int a = 0; a < T.nv; ++a) v[a] = z[ix[a]];
        T.fn(v, par + fm.par_off + (size_t)r * T.np, &fv, g, h);
        double w = 1.0;
        if (fm.kind == 1) {
          w = yE[fm.row0 + r];
          const int pr = M.pos[n + fm.row0 + r];
          for (int q = 0; q < T.ngnz; ++q) {
            const int pc = M.pos[ix[T.gnz[q]]];
            if (pr >= pc) KE(M, K, pr, pc) += g[q]; else KE(M, K, pc, pr) += g[q];
          }
        } else if (fm.kind == 2) {
          w = yI[fm.row0 + r];
          const double sg = Ss[fm.row0 + r];
          for (int a = 0; a < T.ngnz; ++a)
            for (int b = 0; b <= a; ++b) {
              int pa = M.pos[ix[T.gnz[a]]], pb = M.pos[ix[T.gnz[b]]];
              if (pa < pb) std::swap(pa, pb);
              double val = sg * g[a] * g[b];
              if (a != b && ix[T.gnz[a]] == ix[T.gnz[b]]) val *= 2.0;      // (never: a template's variables are distinct)
              KE(M, K, pa, pb) += val;
            }
        }
        for (int q = 0; q < T.nh; ++q) {
          int pa = M.pos[ix[T.hi[q]]], pb = M.pos[ix[T.hj[q]]];
          if (pa < pb) std::swap(pa, pb);
          KE(M, K, pa, pb) += w * h[q];
        }
      }
    }
    for (int i = 0; i < n; ++i) KE(M, K, M.pos[i], M.pos[i]) += Sz[i] + dw;
    for (int i = 0; i < mE; ++i) KE(M, K, M.pos[n + i], M.pos[n + i]) -= dc[i];
  }
  // skyline LDL' without pivoting; returns true if the inertia is (n, mE, 0) -- counted from the signs of D
  bool factor() {
    const int NTl = NT;
    int npos = 0, nneg = 0;
    for (int i = 0; i < NTl; ++i) {
      const int fi = M.first[i];
      double* Ki = &K[M.rowp[i]] - fi;      // Ki[j] = K(i, j)
      for (int j = fi; j < i; ++j) {
        const int fj = M.first[j];
        const double* Kj = &K[M.rowp[j]] - fj;
        double acc = Ki[j];
        for (int k = std::max(fi, fj); k < j; ++k) acc -= Ki[k] * Kj[k];     // Ki[k] holds L(i,k) D(k) (see below), Kj[k] holds L(j,k)
        Ki[j] = acc;                                                        // = L(i,j) D(j) for now
      }
      double d = Ki[i];
      for (int j = fi; j < i; ++j) {
        const double lij = Ki[j] / D[j];
        d -= lij * Ki[j];
        Ki[j] = lij;                                                        // final L(i, j)
      }
      // rows processed later use Kj[k] = L(j,k) (final) and Ki[k] = L(i,k) D(k): restore that convention for row i's use as "j"
      D[i] = d;
      if (!(fabs(d) > 1e-300) || !fin(d)) return false;
      if (d > 0) ++npos; else ++nneg;
    }
    return npos == n && nneg == mE;
  }
  void ldl_solve(const double* b, double* x) {
    for (int i = 0; i < NT; ++i) {
      const int fi = M.first[i];
      const double* Ki = &K[M.rowp[i]] - fi;
      double acc = b[i];
      for (int j = fi; j < i; ++j) acc -= Ki[j] * x[j];
      x[i] = acc;
    }
    for (int i = 0; i < NT; ++i) x[i] /= D[i];
    for (int i = NT - 1; i >= 0; --i) {
      const int fi = M.first[i];
      const double* Ki = &K[M.rowp[i]] - fi;
      const double xi = x[i];
      for (int j = fi; j < i; ++j) x[j] -= Ki[j] * xi;
    }
  }

  void solve(const double* z0) {
    for (int i = 0; i < n; ++i) { hasL[i] = fin(M.zL[i]); hasU[i] = fin(M.zU[i]); z[i] = push(z0[i], M.zL[i], M.zU[i], O.kappa1, O.kappa2); }
    Eval(M, par).cons(z.data(), cE.data(), gI.data());
    n_bmult = 0;
    for (int i = 0; i < n; ++i) { zLm[i] = hasL[i] ? 1.0 : 0.0; zUm[i] = hasU[i] ? 1.0 : 0.0; n_bmult += hasL[i] + hasU[i]; }
    for (int i = 0; i < mI; ++i) {
      sL[i] = fin(gL[i]); sU[i] = fin(gU[i]);
      s[i] = push(gI[i], gL[i], gU[i], O.kappa1, O.kappa2);
      vL[i] = sL[i] ? 1.0 : 0.0; vU[i] = sU[i] ? 1.0 : 0.0; n_bmult += sL[i] + sU[i];
    }
    n_mult = mE + mI + n_bmult;
    double mu = O.mu_init, tau = std::max(O.tau_min, 1 - mu);
    const double mu_min = O.tol * O.mu_min_factor;
    std::vector<std::pair<double, double> > filt;
    int n_kick = 0;
    const double th0 = theta(z.data(), s.data());
    const double theta_max = 1e4 * std::max(1.0, th0), theta_min = 1e-4 * std::max(1.0, th0);
    double dw_last = 0.0;
    status = 0; err = INFINITY;
    std::vector<double> yI(mI), resid(NT), corr(NT), Kcopy(M.nenv), zero(mE, 0.0), gdummy(n), accv(n);      // (no allocation inside the loop)
    int it = 0;
    for (it = 0; it <= O.max_iter; ++it) {
      double dinf, cinf, pinf;
      const double e0 = errors(0.0, &dinf, &cinf, &pinf);
      err = e0;
      if (e0 <= O.tol && dinf <= O.dual_inf_tol && cinf <= O.constr_viol_tol && pinf <= O.compl_inf_tol) { status = 1; break; }
      if (it == O.max_iter) { status = 0; break; }
      bool changed = false;
      while (mu > mu_min) {
        const double emu = errors(mu, nullptr, nullptr, nullptr);
        if (emu > O.kappa_eps * mu) break;
        mu = std::max(mu_min, std::min(O.kappa_mu * mu, pow(mu, O.theta_mu)));
        tau = std::max(O.tau_min, 1 - mu);
        changed = true;
      }
      if (changed) filt.clear();
      // ---- condensed system ----
      for (int i = 0; i < mI; ++i) yI[i] = vU[i] - vL[i];
      grad_lag(z.data(), yE.data(), yI.data(), gf.data(), rz.data());
      Eval(M, par).cons(z.data(), cE.data(), gI.data());
      for (int i = 0; i < n; ++i) {
        const double a = hasL[i] ? z[i] - M.zL[i] : 1.0, b = hasU[i] ? M.zU[i] - z[i] : 1.0;
        Sz[i] = (hasL[i] ? zLm[i] / a : 0.0) + (hasU[i] ? zUm[i] / b : 0.0);
        bz[i] = gf[i] - (hasL[i] ? mu / a : 0.0) + (hasU[i] ? mu / b : 0.0);
      }
      for (int i = 0; i < mI; ++i) {
        const double c = sL[i] ? s[i] - gL[i] : 1.0, d = sU[i] ? gU[i] - s[i] : 1.0;
        Ss[i] = (sL[i] ? vL[i] / c : 0.0) + (sU[i] ? vU[i] / d : 0.0);
        gam[i] = -(sL[i] ? mu / c : 0.0) + (sU[i] ? mu / d : 0.0);
      }
      // rhs = [-(bz + JI'(gam + Ss cI)), -cE - dc yE]   (in unknown order, permuted below)
      {
        std::vector<double>& w = tmp;      // reuse: weights of the inequality rows
        for (int i = 0; i < mI; ++i) w[i] = gam[i] + Ss[i] * (gI[i] - s[i]);
        grad_lag(z.data(), zero.data(), w.data(), gdummy.data(), accv.data());     // accv = grad f + JI' w
        for (int i = 0; i < n; ++i) rhs[M.pos[i]] = -(bz[i] + (accv[i] - gdummy[i]));
        for (int i = 0; i < mE; ++i) rhs[M.pos[n + i]] = -cE[i] - dc[i] * yE[i];
        if (O.freeze_degenerate > 0) {      // oracle/ipm_ref.py: rows |p|^2 == 1 linearised at p = 0 carry no information
          double v[64], g[64], h[1024], fv;
          for (const Fam& fm : M.fam) if (fm.kind == 1) {
            const TemplateInfo& T = TEMPLATES[fm.tmpl];
            for (int r = 0; r < fm.n; ++r) {
              if (!(dc[fm.row0 + r] > 0)) continue;
              const int* ix = fm.idx + (size_t)r * T.nv;
              for (int a = 0; a < T.nv; ++a) v[a] = z[ix[a]];
              T.fn(v, par + fm.par_off + (size_t)r * T.np, &fv, g, h);
              double rn = 0;
              for (int q = 0; q < T.ngnz; ++q) rn = std::max(rn, fabs(g[q]));
              if (rn < O.freeze_degenerate) rhs[M.pos[n + fm.row0 + r]] = -dc[fm.row0 + r] * yE[fm.row0 + r];
            }
          }
        }
      }
      // ---- inertia correction (Algorithm IC) ----
      double dw = 0.0;
      bool ok = false, first = true;
      for (;;) {
        assemble(dw, yI.data());
        Kcopy = K;      // iterative refinement needs the unfactored matrix
        if (factor()) {
          ldl_solve(rhs.data(), sol.data());
          // two steps of iterative refinement with the unfactored matrix
          for (int rep = 0; rep < 2; ++rep) {
            for (int i = 0; i < NT; ++i) resid[i] = rhs[i];
            for (int i = 0; i < NT; ++i) {
              const int fi = M.first[i];
              const double* Ki = &Kcopy[M.rowp[i]] - fi;
              double acc = 0;
              for (int j = fi; j < i; ++j) { acc += Ki[j] * sol[j]; resid[j] -= Ki[j] * sol[i]; }
              resid[i] -= acc + Ki[i] * sol[i];
            }
            ldl_solve(resid.data(), corr.data());
            for (int i = 0; i < NT; ++i) sol[i] += corr[i];
          }
          ok = true;
          break;
        }
        if (first) { dw = dw_last == 0.0 ? O.dw_first : std::max(O.dw_min, O.kw_minus * dw_last); first = false; }
        else dw *= (dw_last == 0.0 ? O.kw_plus_first : O.kw_plus);
        if (dw > O.dw_max) break;
      }
      if (!ok) { status = -2; break; }
      if (dw > 0) dw_last = dw;
      for (int i = 0; i < n; ++i) dz[i] = sol[M.pos[i]];
      for (int i = 0; i < mE; ++i) yEn[i] = sol[M.pos[n + i]];
      // ds = JI dz + cI
      {
        double v[64], g[64], h[1024], fv;
        for (const Fam& fm : M.fam) if (fm.kind == 2) {
          const TemplateInfo& T = TEMPLATES[fm.tmpl];
          for (int r = 0; r < fm.n; ++r) {
            const int* ix = fm.idx + (size_t)r * T.nv;
            for (int a = 0; a < T.nv; ++a) v[a] = z[ix[a]];
            T.fn(v, par + fm.par_off + (size_t)r * T.np, &fv, g, h);
            double acc = fv - s[fm.row0 + r];
            for (int q = 0; q < T.ngnz; ++q) acc += g[q] * dz[ix[T.gnz[q]]];
            ds[fm.row0 + r] = acc;
          }
        }
      }
      double a_pr = 1.0, a_du = 1.0;
      auto amax = [&](double x, double dx, double& a) { if (dx < 0) a = std::min(a, -tau * x / dx); };
      for (int i = 0; i < n; ++i) {
        const double a = z[i] - M.zL[i], b = M.zU[i] - z[i];
        dzL[i] = hasL[i] ? mu / a - zLm[i] - zLm[i] / a * dz[i] : 0.0;
        dzU[i] = hasU[i] ? mu / b - zUm[i] + zUm[i] / b * dz[i] : 0.0;
        if (hasL[i]) { amax(a, dz[i], a_pr); amax(zLm[i], dzL[i], a_du); }
        if (hasU[i]) { amax(b, -dz[i], a_pr); amax(zUm[i], dzU[i], a_du); }
      }
      for (int i = 0; i < mI; ++i) {
        const double c = s[i] - gL[i], d = gU[i] - s[i];
        dvL[i] = sL[i] ? mu / c - vL[i] - vL[i] / c * ds[i] : 0.0;
        dvU[i] = sU[i] ? mu / d - vU[i] + vU[i] / d * ds[i] : 0.0;
        if (sL[i]) { amax(c, ds[i], a_pr); amax(vL[i], dvL[i], a_du); }
        if (sU[i]) { amax(d, -ds[i], a_pr); amax(vU[i], dvU[i], a_du); }
      }
      // ---- filter line search ----
      const double th_k = theta(z.data(), s.data()), ph_k = phi(z.data(), s.data(), mu);
      double dphi = 0;
      for (int i = 0; i < n; ++i) dphi += bz[i] * dz[i];
      for (int i = 0; i < mI; ++i) dphi += gam[i] * ds[i];
      double a_min;
      if (dphi < 0 && th_k <= theta_min)
        a_min = O.gamma_alpha * std::min(O.gamma_theta, std::min(O.gamma_phi * th_k / (-dphi), O.delta * pow(th_k, O.s_theta) / pow(-dphi, O.s_phi)));
      else if (dphi < 0) a_min = O.gamma_alpha * std::min(O.gamma_theta, O.gamma_phi * th_k / (-dphi));
      else a_min = O.gamma_alpha * O.gamma_theta;
      double alpha = a_pr;
      bool accepted = false, ftype = false;
      int nbt = 0;
      while (alpha >= a_min && nbt < O.max_backtrack) {
        for (int i = 0; i < n; ++i) zt[i] = z[i] + alpha * dz[i];
        for (int i = 0; i < mI; ++i) st[i] = s[i] + alpha * ds[i];
        const double th_t = theta(zt.data(), st.data()), ph_t = phi(zt.data(), st.data(), mu);
        bool in_filter = th_t >= theta_max;
        for (size_t q = 0; q < filt.size() && !in_filter; ++q) in_filter = th_t >= filt[q].first && ph_t >= filt[q].second;
        if (!in_filter && std::isfinite(ph_t)) {
          const bool sw = dphi < 0 && alpha * pow(-dphi, O.s_phi) > O.delta * pow(th_k, O.s_theta);
          if (th_k <= theta_min && sw) {
            if (ph_t <= ph_k + O.eta_phi * alpha * dphi + 10 * 2.220446049250313e-16 * fabs(ph_k)) { accepted = true; ftype = true; }
          } else if (th_t <= (1 - O.gamma_theta) * th_k || ph_t <= ph_k - O.gamma_phi * th_k) accepted = true;
        }
        if (accepted) break;
        alpha *= 0.5; ++nbt;
      }
      if (!accepted) {
        if (n_kick < O.max_kick) { ++n_kick; filt.clear(); mu = std::min(O.mu_init, 10.0 * mu); tau = std::max(O.tau_min, 1 - mu); continue; }
        status = -1;
        break;
      }
      if (!ftype) filt.push_back(std::make_pair((1 - O.gamma_theta) * th_k, ph_k - O.gamma_phi * th_k));
      const double a_y = std::min(alpha, a_du), ks = O.kappa_sigma;
      for (int i = 0; i < n; ++i) {
        z[i] += alpha * dz[i]; zLm[i] += a_du * dzL[i]; zUm[i] += a_du * dzU[i];
        if (hasL[i]) { const double a = z[i] - M.zL[i]; zLm[i] = std::min(std::max(zLm[i], mu / (ks * a)), ks * mu / a); } else zLm[i] = 0.0;
        if (hasU[i]) { const double b = M.zU[i] - z[i]; zUm[i] = std::min(std::max(zUm[i], mu / (ks * b)), ks * mu / b); } else zUm[i] = 0.0;
      }
      for (int i = 0; i < mE; ++i) yE[i] += a_y * (yEn[i] - yE[i]);
      for (int i = 0; i < mI; ++i) {
        s[i] += alpha * ds[i]; vL[i] += a_du * dvL[i]; vU[i] += a_du * dvU[i];
        if (sL[i]) { const double c = s[i] - gL[i]; vL[i] = std::min(std::max(vL[i], mu / (ks * c)), ks * mu / c); } else vL[i] = 0.0;
        if (sU[i]) { const double d = gU[i] - s[i]; vU[i] = std::min(std::max(vU[i], mu / (ks * d)), ks * mu / d); } else vU[i] = 0.0;
      }
    }
    iters = std::min(it, O.max_iter);
  }
};

}  // namespace

extern "C" {

int cpu_ipm_n_templates(void) { return N_TEMPLATES; }
const char* cpu_ipm_template_name(int i) { return (i >= 0 && i < N_TEMPLATES) ? TEMPLATES[i].name : ""; }
int cpu_ipm_opts_size(void) { return (int)sizeof(Opts); }

// One model structure (families, bounds, ordering), B problems that differ in their parameter blob and starting point.
//   fam_desc: nf x 6 ints {kind, tmpl, n, idx_off (into idx_all), par_off (into one problem's blob), bound_off (into lo_all / hi_all, or -1)}
// Outputs per problem: z (n), status, iters, err, seconds (the solve alone).  Returns 0, or -1 on a malformed model.
int cpu_ipm_solve_batch(int n, int mE, int mI, int nf, const int* fam_desc, const int* idx_all, const double* lo_all, const double* hi_all,
                        const double* zL, const double* zU, const unsigned char* dc_rows, const int* order, const void* opts_, int B,
                        int par_len, const double* par_all, const double* z0_all, int nthreads, double* z_out, int* status, int* iters,
                        double* err, double* seconds, double* setup_seconds) {
  const Opts& O = *reinterpret_cast<const Opts*>(opts_);
  auto t0 = std::chrono::steady_clock::now();
  Model M;
  M.n = n; M.mE = mE; M.mI = mI; M.nf = nf; M.zL = zL; M.zU = zU; M.dc_rows = dc_rows; M.order = order;
  int rE = 0, rI = 0;
  for (int f = 0; f < nf; ++f) {
    const int* d = fam_desc + 6 * f;
    if (d[1] < 0 || d[1] >= N_TEMPLATES) return -1;
    Fam F;
    F.kind = d[0]; F.tmpl = d[1]; F.n = d[2]; F.idx = idx_all + d[3]; F.par_off = d[4];
    F.lo = (d[5] >= 0 && lo_all) ? lo_all + d[5] : nullptr; F.hi = (d[5] >= 0 && hi_all) ? hi_all + d[5] : nullptr;
    F.row0 = F.kind == 1 ? rE : (F.kind == 2 ? rI : 0);
    if (F.kind == 1) rE += F.n;
    if (F.kind == 2) rI += F.n;
    M.fam.push_back(F);
  }
  if (rE != mE || rI != mI) return -1;
  build_structure(M);
  if (setup_seconds) *setup_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
#pragma omp parallel for schedule(dynamic) num_threads(nthreads > 0 ? nthreads : 1)
  for (int b = 0; b < B; ++b) {
    auto t1 = std::chrono::steady_clock::now();
    Ipm S(M, O, par_all + (size_t)b * par_len);
    S.solve(z0_all + (size_t)b * n);
    memcpy(z_out + (size_t)b * n, S.z.data(), (size_t)n * sizeof(double));
    status[b] = S.status; iters[b] = S.iters; err[b] = S.err;
    seconds[b] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
  }
  return 0;
}
}

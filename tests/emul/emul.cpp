// tests/emul/emul.cpp -- DEVELOPMENT / TEST TOOL ONLY.
//
// Compiles the OBCA_HD kernel sources (obca_b200/csrc/*.cuh) with g++ and runs ONE emulated CTA sequentially on the
// host, so that the solver logic can be debugged and compared with the oracle in the CPU-only build container.
// It is not linked into libobca.so, it is not a fallback of the product path (libobca.so fails loudly without a
// CUDA device), and nothing in bench.py's measured path uses it.
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../../obca_b200/csrc/obca_host.h"
#include "../../obca_b200/csrc/obca_dualws.cuh"
#include "../../obca_b200/csrc/obca_eval.cuh"
#include "../../obca_b200/csrc/obca_quad.cuh"

using namespace obca;

template <int VM, bool SDV>
static void run_one(const ParkProblem& P, const IpmOpts& O, const PkLay& L, double* W, const PkInputs& in,
                    const PkOutputs& out, ProbState& S) {
  PkCtx C;
  std::vector<double> ric((size_t)(P.N + 2) * RSTRIDE, 0.0);
  std::vector<double> tile(256, 0.0);
  static const bool use_warp = getenv("OBCA_EMUL_SERIAL_KKT") == nullptr;
  C.P = &P; C.O = &O; C.L = L; C.W = W; C.ric = ric.data(); C.pp = ric.data(); C.Wd = W + (size_t)L.dLAM * L.NSP; C.red_scratch = nullptr; C.tile = use_warp ? tile.data() : nullptr; C.S = &S; C.in = in;
  IpmDriver<ParkSolver<VM, SDV> >::solve(C);
  for (int k = 0; k <= P.N; ++k) ParkSolver<VM, SDV>::store_stage(C, k, out);
}

template <int VM>
static void dualws_one(const ParkProblem& P, int j, double X, double Y, double psi, double* lam, double* mu, double* d,
                       int* its) {
  ObsRows<VM> R;
  R.v = P.vOb[j];
  for (int i = 0; i < VM; ++i) {
    const bool on = i < R.v;
    const int r = P.voff[j] + (on ? i : 0);
    R.a1[i] = on ? P.A[r][0] : 0.0; R.a2[i] = on ? P.A[r][1] : 0.0; R.bb[i] = on ? P.b[r] : 0.0;
  }
  *its = dualws_solve<VM>(P, R, X, Y, psi, 1e-5, 100, lam, mu, d);
}

extern "C" {

int emul_default_opts(IpmOpts* o) { *o = default_opts(); return (int)sizeof(IpmOpts); }

int emul_layout_total(int N, int nOb, const int* vOb, int signed_dist) {
  ParkProblem P;
  double A[2 * OBCA_MAX_ROWS] = {0}, b[OBCA_MAX_ROWS] = {0}, ego[4] = {1, 1, 1, 1}, xy[4] = {0, 1, 0, 1};
  if (fill_problem(P, N, nOb, vOb, A, b, 1.0, 1.0, ego, xy, 0, signed_dist)) return -1;
  PkLay L = make_layout(P, nfac_for(P));
  return L.total * L.NSP;
}

// B problems, inputs stacked problem-major in the Julia shapes (see include/obca.h).
int emul_parking_solve_batch(int B, int N, int nOb, const int* vOb, const double* A, const double* b,
                             const double* x0, const double* xF, double Ts, double L, const double* ego,
                             const double* XYbounds, const double* rx, const double* ry, const double* ryaw,
                             const double* xWS, const double* uWS, const double* lWS, const double* nWS, int fixTime,
                             int signed_dist, const IpmOpts* opts, double* xp, double* up, double* ts, double* lp,
                             double* np, double* sl, double* duals, int* status, int* iters, double* kkt_err,
                             int* nfact, double* Wdump) {
  ParkProblem P;
  int rc = fill_problem(P, N, nOb, vOb, A, b, Ts, L, ego, XYbounds, fixTime, signed_dist);
  if (rc) return rc;
  IpmOpts O = opts ? *opts : default_opts();
  PkLay Lay = make_layout(P, nfac_for(P));
  const int NS = N + 1, V = P.V;
  const size_t nd = (size_t)4 * N + (size_t)2 * nOb * NS + (size_t)2 * nOb * NS;
#pragma omp parallel for schedule(dynamic)
  for (int i = 0; i < B; ++i) {
    std::vector<double> W((size_t)Lay.total * Lay.NSP, 0.0);
    PkInputs in;
    in.x0 = x0 + 4 * i; in.xF = xF + 4 * i;
    in.rx = rx + (size_t)NS * i; in.ry = ry + (size_t)NS * i; in.ryaw = ryaw + (size_t)NS * i;
    in.xWS = xWS + (size_t)4 * NS * i; in.ldx = NS;
    in.uWS = uWS + (size_t)2 * N * i; in.ldu = N;
    in.lWS = lWS + (size_t)V * NS * i; in.nWS = nWS + (size_t)4 * nOb * NS * i;
    PkOutputs out;
    out.xp = xp + (size_t)4 * NS * i; out.up = up + (size_t)2 * N * i; out.ts = ts + (size_t)NS * i;
    out.lp = lp + (size_t)V * NS * i; out.np = np + (size_t)4 * nOb * NS * i;
    out.sl = sl ? sl + (size_t)nOb * NS * i : nullptr;
    out.duals = duals ? duals + nd * i : nullptr;
    ProbState S;
    const int vm = max_vob(P) <= 2 ? 2 : 4;
    if (signed_dist) {
      if (vm == 2) run_one<2, true>(P, O, Lay, W.data(), in, out, S);
      else run_one<4, true>(P, O, Lay, W.data(), in, out, S);
    } else {
      if (vm == 2) run_one<2, false>(P, O, Lay, W.data(), in, out, S);
      else run_one<4, false>(P, O, Lay, W.data(), in, out, S);
    }
    status[i] = S.status; iters[i] = S.iters; kkt_err[i] = S.e0;
    if (nfact) nfact[i] = S.n_fact;
    if (Wdump && i == 0) memcpy(Wdump, W.data(), W.size() * sizeof(double));
  }
  return 0;
}

// DualMultWS.jl:29 call surface, batched: lp (N+1) x V and np (N+1) x 4nOb per problem, column-major (as returned, :81-84)
int emul_dualmultws_batch(int B, int N, int nOb, const int* vOb, const double* A, const double* b, const double* ego,
                          const double* rx, const double* ry, const double* ryaw, double* lp, double* np, double* dd,
                          int* its_out) {
  ParkProblem P;
  double xy[4] = {0, 1, 0, 1};
  int rc = fill_problem(P, N, nOb, vOb, A, b, 1.0, 1.0, ego, xy, 0, 1);
  if (rc) return rc;
  const int NS = N + 1, V = P.V;
  const int vm = max_vob(P) <= 2 ? 2 : 4;
#pragma omp parallel for
  for (int i = 0; i < B; ++i) {
    for (int k = 0; k < NS; ++k)
      for (int j = 0; j < nOb; ++j) {
        double lam[4], mu[4], d; int its;
        const size_t o = (size_t)NS * i + k;
        if (vm == 2) dualws_one<2>(P, j, rx[o], ry[o], ryaw[o], lam, mu, &d, &its);
        else dualws_one<4>(P, j, rx[o], ry[o], ryaw[o], lam, mu, &d, &its);
        for (int r = 0; r < P.vOb[j]; ++r) lp[(size_t)V * NS * i + (size_t)(P.voff[j] + r) * NS + k] = lam[r];
        for (int m = 0; m < 4; ++m) np[(size_t)4 * nOb * NS * i + (size_t)(4 * j + m) * NS + k] = mu[m];
        if (dd) dd[(size_t)nOb * NS * i + (size_t)j * NS + k] = d;
        if (its_out) its_out[(size_t)nOb * NS * i + (size_t)j * NS + k] = its;
      }
  }
  return 0;
}

// K1 stand-alone evaluator (obca_eval.cuh) on the host: same arguments as obca_parking_eval_batch_dev
int emul_parking_eval_batch(int B, int N, int nOb, const int* vOb, const double* A, const double* b, const double* x0,
                            const double* xF, double Ts, double L, const double* ego, const double* XYbounds,
                            const double* rx, const double* ry, const double* ryaw, const double* xp, const double* up,
                            const double* ts, const double* lp, const double* np, const double* sl, const double* y,
                            int fixTime, int signed_dist, double* c_out, double* gradL_out, double* fk_out) {
  ParkProblem P;
  int rc = fill_problem(P, N, nOb, vOb, A, b, Ts, L, ego, XYbounds, fixTime, signed_dist);
  if (rc) return rc;
  const size_t NS = N + 1, V = P.V, m = eval_m(P), n = eval_n(P);
  const int vm = max_vob(P) <= 2 ? 2 : 4;
  for (int bb = 0; bb < B; ++bb) {
    EvalIn ib;
    ib.x0 = x0 + 4 * bb; ib.xF = xF + 4 * bb; ib.rx = rx + NS * bb; ib.ry = ry + NS * bb; ib.ryaw = ryaw + NS * bb;
    ib.xp = xp + 4 * NS * bb; ib.up = up + (size_t)2 * N * bb; ib.ts = ts ? ts + NS * bb : nullptr;
    ib.lp = lp + V * NS * bb; ib.np = np + (size_t)4 * nOb * NS * bb; ib.sl = sl ? sl + (size_t)nOb * NS * bb : nullptr;
    ib.y = y ? y + m * bb : nullptr;
    EvalOut ob;
    ob.c = c_out + m * bb; ob.gradL = gradL_out + n * bb; ob.fk = fk_out + NS * bb;
    for (int k = 0; k <= N; ++k) {
      if (signed_dist) { if (vm == 2) eval_stage<2, true>(P, k, ib, ob); else eval_stage<4, true>(P, k, ib, ob); }
      else { if (vm == 2) eval_stage<2, false>(P, k, ib, ob); else eval_stage<4, false>(P, k, ib, ob); }
    }
  }
  return 0;
}

// quadcopter (QuadcopterSignedDist.jl:25 / QuadcopterDist.jl:25), B problems; x0, xF 12 x B; obs 6 x 5 shared;
// xWS 12 x (N+1) per problem; outputs in the reference's shapes
int emul_quadcopter_solve_batch(int B, int N, const double* x0, const double* xF, double Ts, double R, const double* obs,
                                const double* xWS, double timeWS, int signed_dist, const IpmOpts* opts, double* xp,
                                double* up, double* ts, double* lp, double* slack, int* status, int* iters,
                                double* kkt_err, int* nfact) {
  QuadProblem P;
  if (fill_quad_problem(P, N, Ts, R, obs, signed_dist)) return -1;
  IpmOpts O = opts ? *opts : default_opts();
  QLay Lay = make_qlayout(P);
  const size_t NS = N + 1;
#pragma omp parallel for schedule(dynamic)
  for (int i = 0; i < B; ++i) {
    std::vector<double> W((size_t)Lay.total * Lay.NSP, 0.0);
    QCtx C;
    ProbState S;
    C.P = &P; C.O = &O; C.L = Lay; C.W = W.data(); C.red_scratch = nullptr; C.tile = nullptr; C.S = &S;
    C.in.x0 = x0 + 12 * (size_t)i; C.in.xF = xF + 12 * (size_t)i; C.in.xWS = xWS + 12 * NS * i; C.in.timeWS = timeWS;
    QOutputs out;
    out.xp = xp + 12 * NS * i; out.up = up + (size_t)4 * N * i; out.ts = ts + NS * i; out.lp = lp + 30 * NS * i;
    out.slack = slack ? slack + 5 * NS * i : nullptr;
    if (signed_dist) {
      IpmDriver<QuadSolver<true> >::solve(C);
      for (int k = 0; k <= N; ++k) QuadSolver<true>::store_stage(C, k, out);
    } else {
      IpmDriver<QuadSolver<false> >::solve(C);
      for (int k = 0; k <= N; ++k) QuadSolver<false>::store_stage(C, k, out);
    }
    status[i] = S.status; iters[i] = S.iters; kkt_err[i] = S.e0;
    if (nfact) nfact[i] = S.n_fact;
  }
  return 0;
}
}

// obca_phased.cuh -- phase-split ("lock-step") driver of the batched parking solve.
//
// The interior-point loop of IpmDriver<M>::solve (obca_solver.cuh) restated as a per-problem state machine so that
// every phase of an iteration runs as its own kernel over ALL active problems of the batch, each with the launch
// shape that suits it:
//     k_pk_block    flat, thread per (problem, obstacle, stage) : the OBCA constraint blocks of K1 -- evaluation, KKT-error
//                                                        partials, condensation onto the stage pose (streaming, no barriers)
//     k_pk_phaseA   CTA per problem, thread per stage : stage terms of K1 (dynamics, objective, bounds), assembly of the
//                                                        stage models, reductions, convergence test, barrier update
//     (both once more for the problems whose barrier parameter was just reduced)
//     k_pk_sweep    HALF-WARP per problem              : K3 stage-banded KKT sweep, stage slots streamed from HBM
//                                                        through a cp.async ring; inertia-correction bookkeeping
//     k_pk_rblock   flat, thread per (problem, obstacle, stage) : K4 step of the block unknowns from the stored local map
//     k_pk_phaseC   CTA per problem, thread per stage : K4 costates, fraction-to-the-boundary, filter line search, update
// The iterate, the step, the local maps, the block hand-over records and the stage slots of every problem live in HBM (layout: problem-major,
// then [array][stage], so each warp streams contiguous slices); between kernels a problem is described by its
// ProbState record (phase, iteration, barrier parameter, filter, ...).
// When the active set has shrunk below what one wave of resident CTAs can hold, the remaining problems are handed to
//     k_pk_tail     persistent CTA per problem: the same three phase functions in a loop, stage slots in shared memory
// which is also the whole solver for small batches.
//
// The arithmetic and its order are exactly those of IpmDriver<M>::solve, and the library is built without FMA contraction:
// all schedules produce bit-identical iterates (tests/test_gpu_parking.py::test_phased_equals_persistent).
#pragma once
#include "obca_check.cuh"
#include "obca_solver.cuh"

namespace obca {

enum { PH_INIT = 0, PH_EVAL = 1, PH_EVAL2 = 2, PH_REASM = 3, PH_KKT = 4, PH_RECOVER = 5, PH_END = 6, PH_DONE = 7 };
constexpr int GSTRIDE = 80;   // doubles per stage slot in global memory: RSTRIDE rounded up to a 16-byte multiple

struct BatchPtrs {
  const double *x0, *xF, *rx, *ry, *ryaw, *xWS, *uWS, *lWS, *nWS;
  double *xp, *up, *ts, *lp, *np, *sl, *duals;
  int *exitflag, *iters;
  double* kkt_err;
  int B;
  int retry;
  unsigned long long* prof;   // optional: 8 cycle counters summed over the batch (device pointer)
};

#if defined(__CUDACC__)

// ------------------------------------------------------------------------------------------------------------
// the three phase functions (all threads of a CTA call them with the same context; ProbState in shared memory)
// ------------------------------------------------------------------------------------------------------------
template <class M>
struct PhasedDriver {
  typedef typename M::Ctx Ctx;
  typedef IpmDriver<M> D;

  // A: [init] -> evaluate at the current iterate, convergence test, barrier update (+ re-evaluation), or re-assembly
  //    with a larger delta_w.  Leaves phase = PH_KKT (stage models ready) or PH_END (attempt over).
  //    BLK (phase-split rounds): the obstacle blocks come from the flat block kernel through C.bo, so an evaluation can only
  //    run when that kernel has seen the current iterate / barrier parameter: after the initial point the function returns
  //    with PH_EVAL, after a barrier update with PH_EVAL2, and is called again once the blocks have been refreshed.
  template <bool BLK>
  __device__ static void eval_all(const Ctx& C, bool do_err, EvalPart& ep) {
    const int NS = M::n_stages(C);
    part_init(ep);
    OBCA_FOR_STAGES(k, NS) {
      EvalPart e1;
      if (BLK) M::stage_eval_blk(C, k, do_err, true, e1); else M::stage_eval(C, k, do_err, true, e1);
      part_merge(ep, e1);
    }
    OBCA_REDUCE(ep);
  }
  template <bool BLK>
  __device__ static void phase_A(const Ctx& C) {
    const IpmOpts& O = CTX_O(C);
    ProbState& S = *C.S;
    const int NS = M::n_stages(C);
    const int ph = S.phase;
    EvalPart ep;
    if (ph == PH_REASM) {
      eval_all<BLK>(C, false, ep);
      OBCA_SERIAL { S.ok = ep.ok; S.phase = PH_KKT; S.prof[7]++; }
      OBCA_SYNC();
      return;
    }
    if (ph == PH_INIT) {
      const int restart = S.attempt;
      OBCA_SERIAL {
        M::init_scalars(C, restart);
        S.mu = O.mu_init; S.tau = dmax(O.tau_min, 1.0 - O.mu_init);
        S.dw = 0.0; S.dw_last = 0.0; S.nfilt = 0; S.status = 0; S.iters = 0; S.n_fact = 0; S.n_kick = 0;
        S.it = 0; S.first = 1;
      }
      OBCA_SYNC();
      OBCA_FOR_STAGES(k, NS) M::init_stage(C, k, restart);
      OBCA_SYNC();
      OBCA_FOR_STAGES(k, NS) M::init_slacks(C, k);
      // (S.phase may only change behind a barrier: every thread read it on entry)
      OBCA_SERIAL { S.phase = PH_EVAL; }
      OBCA_SYNC();
      if (BLK) return;
    }
    if (ph != PH_EVAL2) {
      // (PH_EVAL: new iterate -- or a barrier kick: new direction from the same iterate)
      OBCA_SERIAL { S.dw = 0.0; }
      OBCA_SYNC();
      eval_all<BLK>(C, true, ep);
      OBCA_SERIAL {
        S.prof[7]++;
        D::apply_errors(C, ep);
        S.ok = ep.ok;
        if (S.first) {
          S.theta_max = 1e4 * dmax(1.0, S.th_k);
          S.theta_min = 1e-4 * dmax(1.0, S.th_k);
          S.first = 0;
        }
        S.e0 = D::err_mu(C, 0.0);
        S.iters = S.it;
        S.flag = 0;
        if (S.e0 <= O.tol && S.e_dual <= O.dual_inf_tol && S.e_pr <= O.constr_viol_tol && S.e_cmax <= O.compl_inf_tol) {
          S.status = 1; S.flag = 1;
        } else if (S.it >= O.max_iter) {
          S.status = 0; S.flag = 1;
        } else {
          bool changed = false;
          while (S.mu > O.mu_min && D::err_mu(C, S.mu) <= O.kappa_eps * S.mu) {
            S.mu = dmax(O.mu_min, dmin_(O.kappa_mu * S.mu, pow(S.mu, O.theta_mu)));
            S.tau = dmax(O.tau_min, 1.0 - S.mu);
            changed = true;
          }
          if (changed) { S.nfilt = 0; S.flag = 2; }
        }
        S.phase = (S.flag == 1) ? PH_END : (S.flag == 2 ? PH_EVAL2 : PH_KKT);
      }
      OBCA_SYNC();
      if (S.flag != 2 || BLK) return;
    }
    // mu changed: the barrier terms of the stage models (and phi) are stale
    eval_all<BLK>(C, true, ep);
    OBCA_SERIAL { D::apply_errors(C, ep); S.ok = ep.ok; S.prof[7]++; S.phase = PH_KKT; }
    OBCA_SYNC();
  }

  // B (serial part, one thread, after the sweep): Ipopt's inertia-correction bookkeeping.  `ok`: 1 if every pivot of
  // the sweep had the required sign, 0 otherwise (also when the local blocks already failed and the sweep was skipped).
  __device__ static void phase_B_serial(ProbState& S, const IpmOpts& O, int ok) {
    S.n_fact++;
    if (!ok) {
      if (S.dw == 0.0) S.dw = (S.dw_last == 0.0) ? O.dw_first : dmax(O.dw_min, O.kw_minus * S.dw_last);
      else S.dw *= (S.dw_last == 0.0) ? O.kw_plus_first : O.kw_plus;
      if (S.dw > O.dw_max) { S.status = -2; ok = -1; }
    } else if (S.dw > 0.0) {
      S.dw_last = S.dw;
    }
    S.ok = ok;
    S.phase = ok > 0 ? PH_RECOVER : (ok == 0 ? PH_REASM : PH_END);
  }

  // C: recover the full step, step lengths, filter line search, and -- when a step is accepted -- the iterate update.
  //    Leaves PH_EVAL (new iterate, or barrier kick: same iterate, new barrier parameter) or PH_END (line-search failure).
  template <bool BLK>
  __device__ static void phase_C(const Ctx& C) {
    const IpmOpts& O = CTX_O(C);
    ProbState& S = *C.S;
    const int NS = M::n_stages(C);
    StepPart sp;
    part_init(sp);
    OBCA_FOR_STAGES(k, NS) {
      StepPart s1;
      if (BLK) M::recover_stage_blk(C, k, s1); else M::recover_stage(C, k, s1);
      part_merge(sp, s1);
    }
    OBCA_REDUCE(sp);
    OBCA_SERIAL {
      const double apr = sp.apr, adu = sp.adu, dphi = sp.dphi;
      S.a_pr = apr; S.a_du = adu; S.dphi = dphi;
      const double th = S.th_k;
      if (dphi < 0.0 && th <= S.theta_min)
        S.a_min = O.gamma_alpha * dmin_(O.gamma_theta, dmin_(O.gamma_phi * th / (-dphi), O.delta * pow(th, O.s_theta) / pow(-dphi, O.s_phi)));
      else if (dphi < 0.0)
        S.a_min = O.gamma_alpha * dmin_(O.gamma_theta, O.gamma_phi * th / (-dphi));
      else
        S.a_min = O.gamma_alpha * O.gamma_theta;
      S.alpha = apr;
      S.flag = 0;
    }
    OBCA_SYNC();
    for (int nbt = 0;; ++nbt) {
      const double alpha = S.alpha;
      MeritPart mp;
      part_init(mp);
      OBCA_FOR_STAGES(k, NS) { MeritPart m1; M::merit_stage(C, k, alpha, m1); part_merge(mp, m1); }
      OBCA_REDUCE(mp);
      OBCA_SERIAL {
        const double th = mp.th, ph = mp.phi;
        S.th_t = th; S.ph_t = ph;
        bool in_filter = th >= S.theta_max || !(ph < 1e299);
        for (int i = 0; i < S.nfilt && !in_filter; ++i) in_filter = (th >= S.filt_th[i] && ph >= S.filt_ph[i]);
        bool accepted = false, ftype = false;
        if (!in_filter) {
          const bool sw = S.dphi < 0.0 && alpha * pow(-S.dphi, O.s_phi) > O.delta * pow(S.th_k, O.s_theta);
          if (S.th_k <= S.theta_min && sw) {
            if (ph <= S.ph_k + O.eta_phi * alpha * S.dphi + 10.0 * 2.220446049250313e-16 * dabs(S.ph_k)) { accepted = true; ftype = true; }
          } else {
            if (th <= (1.0 - O.gamma_theta) * S.th_k || ph <= S.ph_k - O.gamma_phi * S.th_k) accepted = true;
          }
        }
        if (accepted) {
          S.flag = 1;
          if (!ftype && S.nfilt < 64) {
            S.filt_th[S.nfilt] = (1.0 - O.gamma_theta) * S.th_k;
            S.filt_ph[S.nfilt] = S.ph_k - O.gamma_phi * S.th_k;
            S.nfilt++;
          }
        } else {
          S.alpha = 0.5 * alpha;
          if (S.alpha < S.a_min || nbt + 1 >= O.max_backtrack) {
            if (S.n_kick < O.max_kick) {      // barrier kick (restoration substitute), see IpmDriver::solve
              S.n_kick++;
              S.nfilt = 0;
              S.mu = dmin_(O.mu_init, 10.0 * S.mu);
              S.tau = dmax(O.tau_min, 1.0 - S.mu);
              S.flag = -2;
            } else {
              S.flag = -1; S.status = -1;
            }
          }
        }
        if (S.flag == 1) S.phase = PH_EVAL;
        else if (S.flag == -2) { S.it++; S.phase = PH_EVAL; }
        else if (S.flag == -1) S.phase = PH_END;
      }
      OBCA_SYNC();
      if (S.flag != 0) break;
    }
    if (S.flag == 1) {      // accept
      OBCA_FOR_STAGES(k, NS) M::update_stage(C, k);
      OBCA_SERIAL { M::update_scalars(C); S.it++; }
      OBCA_SYNC();
    }
  }
};

// ------------------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------------------
// ProbState <-> global memory, cooperatively, as 8-byte words
static_assert(sizeof(ProbState) % 8 == 0, "ProbState must be a whole number of 8-byte words");
__device__ __forceinline__ void state_load(ProbState& S, const ProbState* g) {
  const unsigned long long* src = reinterpret_cast<const unsigned long long*>(g);
  unsigned long long* dst = reinterpret_cast<unsigned long long*>(&S);
  for (int i = threadIdx.x; i < (int)(sizeof(ProbState) / 8); i += blockDim.x) dst[i] = src[i];
}
__device__ __forceinline__ void state_store(ProbState* g, const ProbState& S) {
  const unsigned long long* src = reinterpret_cast<const unsigned long long*>(&S);
  unsigned long long* dst = reinterpret_cast<unsigned long long*>(g);
  for (int i = threadIdx.x; i < (int)(sizeof(ProbState) / 8); i += blockDim.x) dst[i] = src[i];
}

template <int VM, bool SDV>
__device__ __forceinline__ void pk_make_ctx(PkCtx& C, PkOutputs& out, const ParkProblem& P, const IpmOpts& O, const PkLay& L,
                                            const BatchPtrs& bp, int b, double* W) {
  const int N = P.N, NS = N + 1, V = P.V, nOb = P.nOb;
  C.P = &P; C.O = &O; C.L = L; C.W = W; C.bo = nullptr;
  C.in.x0 = bp.x0 + 4 * (size_t)b; C.in.xF = bp.xF + 4 * (size_t)b;
  C.in.rx = bp.rx + (size_t)NS * b; C.in.ry = bp.ry + (size_t)NS * b; C.in.ryaw = bp.ryaw + (size_t)NS * b;
  C.in.xWS = bp.xWS + (size_t)4 * NS * b; C.in.ldx = NS;
  C.in.uWS = bp.uWS + (size_t)2 * N * b; C.in.ldu = N;
  C.in.lWS = bp.lWS + (size_t)V * NS * b; C.in.nWS = bp.nWS + (size_t)4 * nOb * NS * b;
  out.xp = bp.xp + (size_t)4 * NS * b; out.up = bp.up + (size_t)2 * N * b; out.ts = bp.ts + (size_t)NS * b;
  out.lp = bp.lp + (size_t)V * NS * b; out.np = bp.np + (size_t)4 * nOb * NS * b;
  out.sl = bp.sl ? bp.sl + (size_t)nOb * NS * b : nullptr;
  out.duals = bp.duals ? bp.duals + ((size_t)4 * N + (size_t)4 * nOb * NS) * b : nullptr;
}

struct PkFinalScratch {
  ChkPart chk[4];
  int feas;
};

// End of an attempt (phase PH_END): the reference's status / retry logic around solve(m)
// (ParkingSignedDist.jl:256-283): a failed first attempt is followed by one more solve from the last iterate; after a
// second failure ParkingConstraints decides.  Leaves PH_INIT (second attempt) or PH_DONE (outputs written).
template <int VM, bool SDV>
__device__ void pk_end_of_attempt(const PkCtx& C, const PkOutputs& out, const BatchPtrs& bp, int b, PkFinalScratch& F) {
  ProbState& S = *C.S;
  const ParkProblem& P = CTX_P(C);
  const int NS = P.N + 1;
  const int status = S.status;
  const bool again = !(status == 1 || !bp.retry || S.attempt == 1);
  __syncthreads();      // everybody has read the state before thread 0 changes it
  if (again) {
    if (threadIdx.x == 0) { S.iters_total += S.iters; S.attempt = 1; S.phase = PH_INIT; }
    __syncthreads();
    return;
  }
  for (int k = threadIdx.x; k < NS; k += blockDim.x) ParkSolver<VM, SDV>::store_stage(C, k, out);
  __syncthreads();
  int exitflag = status == 1 ? 1 : 0;
  if (status != 1 && bp.retry) {
    ChkPart c;
    chk_init(c);
    for (int k = threadIdx.x; k < NS; k += blockDim.x) {
      ChkPart ck;
      check_stage(P, k, C.in.x0, C.in.xF, out.xp, out.up, out.lp, out.np, out.ts, out.sl, SDV ? 1 : 0, 0, ck);
      chk_merge(c, ck);
    }
    for (int off = 16; off > 0; off >>= 1) {
      ChkPart o;
#pragma unroll
      for (int i = 0; i < 5; ++i) o.c0[i] = __shfl_down_sync(0xffffffffu, c.c0[i], off);
      o.c1 = __shfl_down_sync(0xffffffffu, c.c1, off); o.c2 = __shfl_down_sync(0xffffffffu, c.c2, off);
      o.c3 = __shfl_down_sync(0xffffffffu, c.c3, off); o.c4 = __shfl_down_sync(0xffffffffu, c.c4, off);
      o.c5 = __shfl_down_sync(0xffffffffu, c.c5, off); o.c6 = __shfl_down_sync(0xffffffffu, c.c6, off);
      o.sbox = __shfl_down_sync(0xffffffffu, c.sbox, off);
      chk_merge(c, o);
    }
    if ((threadIdx.x & 31) == 0) F.chk[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w) chk_merge(c, F.chk[w]);
      int e[7];
      F.feas = check_finish(P, c, out.ts, 0, 5e-5, e);
    }
    __syncthreads();
    exitflag = F.feas ? 1 : 0;
  }
  if (threadIdx.x == 0) {
    bp.exitflag[b] = exitflag;
    bp.iters[b] = S.iters_total + S.iters;
    bp.kkt_err[b] = S.e0;
    S.phase = PH_DONE;
  }
  __syncthreads();
}

// phase A of one problem including the attempt bookkeeping; returns with phase in {PH_KKT, PH_DONE} (BLK: also PH_EVAL /
// PH_EVAL2 when the block kernel has to run first)
template <int VM, bool SDV, bool BLK>
__device__ __forceinline__ void pk_step_A(const PkCtx& C, const PkOutputs& out, const BatchPtrs& bp, int b, PkFinalScratch& F) {
  ProbState& S = *C.S;
  for (;;) {
    if (S.phase == PH_END) pk_end_of_attempt<VM, SDV>(C, out, bp, b, F);
    if (S.phase == PH_DONE) return;
    PhasedDriver<ParkSolver<VM, SDV> >::template phase_A<BLK>(C);
    if (S.phase != PH_END) return;
  }
}

__device__ __forceinline__ void state_fresh(ProbState& S) {
  if (threadIdx.x == 0) {
    S.phase = PH_INIT; S.attempt = 0; S.iters_total = 0; S.it = 0; S.first = 1; S.status = 0; S.iters = 0;
    S.t = 1.0; S.e0 = 0.0; S.ok = 0; S.flag = 0;
  }
}

// ------------------------------------------------------------------------------------------------------------
// K_blk: the OBCA constraint blocks of K1 as a flat streaming kernel -- one thread per (active problem, obstacle, stage),
// consecutive threads = consecutive stages, so every warp reads contiguous slices of the stacked (x, lambda, mu, sl, duals)
// arrays; no shared memory, no barriers.  Each thread evaluates the signed-distance rows of its block, their
// Lagrangian-gradient / KKT-error pieces, condenses the block onto the stage pose and writes the local factor
// (workspace) and the 22-double hand-over record read by k_pk_phaseA.  pass 1: new iterates (PH_EVAL) and
// inertia-correction re-assemblies (PH_REASM); pass 2: problems whose barrier parameter was just reduced (PH_EVAL2).
// ------------------------------------------------------------------------------------------------------------
#ifndef OBCA_MINB_BLK
#define OBCA_MINB_BLK 3
#endif
template <int VM, bool SDV>
__global__ void __launch_bounds__(128, OBCA_MINB_BLK)
k_pk_block(const __grid_constant__ ParkProblem P, const PkLay L, double* __restrict__ Wall, double* __restrict__ BOall,
           const ProbState* __restrict__ Sg, const int* __restrict__ act, const int* __restrict__ n_act, int pass) {
  const int NS = P.N + 1, per = P.nOb * NS;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int a = (int)(idx / per);
  if (a >= *n_act) return;
  const int rem = (int)(idx - (long long)a * per);
  const int j = rem / NS, k = rem - j * NS;
  const int b = act[a];
  const ProbState* S = Sg + b;
  const int ph = S->phase;
  bool do_err;
  double dw;
  if (pass == 1) {
    if (ph == PH_EVAL) { do_err = true; dw = 0.0; }
    else if (ph == PH_REASM) { do_err = false; dw = S->dw; }
    else return;
  } else {
    if (ph != PH_EVAL2) return;
    do_err = true; dw = 0.0;
  }
  const double mu_b = S->mu;
  PkCtx C;
  C.W = Wall + (size_t)b * L.total * L.NSP;
  C.bo = BOall + (size_t)b * P.nOb * BO_N * L.NSP;
  const double X = WA(X, k), Y = WA(Y, k), ps = WA(PS, k);
  double sn_, cs_;
  sincos(ps, &sn_, &cs_);
  BlockOut B;
  ParkSolver<VM, SDV>::block_eval(C, k, j, X, Y, cs_, sn_, mu_b, dw, do_err, true, B);
  ParkSolver<VM, SDV>::block_store(C, k, j, B);
}

// ------------------------------------------------------------------------------------------------------------
// K_A: one CTA per active problem
// ------------------------------------------------------------------------------------------------------------
#ifndef OBCA_MINB_A
#define OBCA_MINB_A 3
#endif
template <int VM, bool SDV>
__global__ void __launch_bounds__(128, OBCA_MINB_A)
k_pk_phaseA(const __grid_constant__ ParkProblem P, const __grid_constant__ IpmOpts O, const PkLay L, const BatchPtrs bp,
            double* __restrict__ Wall, double* __restrict__ slots, ProbState* __restrict__ Sg,
            double* __restrict__ BOall, const int* __restrict__ act_in, const int* __restrict__ n_in, int* __restrict__ act_out,
            int* __restrict__ n_out, int fresh, int pass) {
  extern __shared__ double s_ric[];     // (N+1) x RSTRIDE stage slots, assembled here, streamed out for the sweep kernel
  __shared__ ProbState S;
  __shared__ double s_red[4 * 12];
  __shared__ PkFinalScratch s_fin;
  if ((int)blockIdx.x >= *n_in) return;
  const int b = fresh ? (int)blockIdx.x : act_in[blockIdx.x];
  const int NS = P.N + 1;
  if (pass == 2 && Sg[b].phase != PH_EVAL2) return;      // second pass of a round: only problems whose barrier parameter changed
  if (fresh) state_fresh(S); else state_load(S, Sg + b);
  __syncthreads();
  if (threadIdx.x == 0) S.prof[7] = 0;      // K1 evaluations of this launch (diagnostic counter, see obca_last_profile)
  // the context is identical for every thread: one copy in shared memory (not one per thread in local memory)
  __shared__ PkCtx C; __shared__ PkOutputs out;
  if (threadIdx.x == 0) {
    pk_make_ctx<VM, SDV>(C, out, P, O, L, bp, b, Wall + (size_t)b * L.total * L.NSP);
    C.ric = s_ric; C.pp = s_ric; C.pps = RSTRIDE; C.red_scratch = s_red; C.tile = nullptr; C.S = &S;
    C.bo = BOall + (size_t)b * P.nOb * BO_N * L.NSP;
  }
  __syncthreads();
  pk_step_A<VM, SDV, true>(C, out, bp, b, s_fin);
  if (S.phase == PH_KKT) {
    double* g = slots + (size_t)b * NS * GSTRIDE;
    for (int i = threadIdx.x; i < NS * RSTRIDE; i += blockDim.x) {
      const int k = i / RSTRIDE, o = i - k * RSTRIDE;
      g[(size_t)k * GSTRIDE + o] = s_ric[i];
    }
  }
  state_store(Sg + b, S);
  if (threadIdx.x == 0) {
    if (bp.prof) atomicAdd(bp.prof + (pass == 1 ? 7 : 5), (unsigned long long)S.prof[7]);
    if (pass == 1 && S.phase != PH_DONE) act_out[atomicAdd(n_out, 1)] = b;
  }
}

// ------------------------------------------------------------------------------------------------------------
// K_B: half-warp per problem.  Backward Riccati sweep + forward roll-out with the stage slots streamed from global memory
// (each slot = 640 contiguous bytes) through a 4-deep cp.async ring; gains and the P_{k+1} rows needed by the
// multiplier recovery are written back in place (consumed Q/q space of the slots).
// ------------------------------------------------------------------------------------------------------------
constexpr int SWEEP_DEPTH = 4;

// ------------------------------------------------------------------------------------------------------------
// Two problems per warp: the lane code of the sweep uses 9 of 32 lanes, so each HALF-warp runs one problem (lanes 0..15 /
// 16..31, same instruction stream, own ring / exchange tile / pointers).  Halves the instruction and shared-memory traffic
// per sweep; the kernel is issue / MIO bound (profiles/ncu_summary_r01.md).  A half without work (odd count, problem not
// waiting for a sweep, local pivots already failed) executes along on valid memory with every global write redirected
// to a shared-memory dump area.
// ------------------------------------------------------------------------------------------------------------
constexpr int SWEEP_HALF_DOUBLES = SWEEP_DEPTH * GSTRIDE + 144 + 48 + 4;   // ring + tile + dump + 4 (bank offset between the halves)

__device__ __forceinline__ void sweep_prefetch16(double* ring, const double* gslots, int k, int hl, int n) {
  if (k >= 0 && k < n) {
    double* dst = ring + (k & (SWEEP_DEPTH - 1)) * GSTRIDE;
    const double* src = gslots + (size_t)k * GSTRIDE;
    cp_async16(dst + 2 * hl, src + 2 * hl);
    cp_async16(dst + 32 + 2 * hl, src + 32 + 2 * hl);
    if (hl < 8) cp_async16(dst + 64 + 2 * hl, src + 64 + 2 * hl);
  }
  cp_async_commit();
}

// run: this half has a sweep to do (uniform within the half).  Returns 1 if the half's sweep found inertia (n, m, 0).
template <int VM, bool SDV>
__device__ int pk_sweep_pair(const ParkProblem& Pp, const IpmOpts& O, const PkLay& Lay, double* W, double* gslots,
                             ProbState* Sgl, bool run, double* ring, double* tile, double* dump) {
  typedef ParkSolver<VM, SDV> PS;
  const int N = Pp.N;
  const int lane = threadIdx.x & 31, hl = lane & 15, base = lane & 16;
  const unsigned FULL = 0xffffffffu;
  PkCtx C;
  C.P = &Pp; C.O = &O; C.L = Lay; C.W = W;
  typename PS::KktLane L;
  // kl_init zeroes tile[70 .. 90] with 21 lanes; a half has 16
  PS::kl_init(L, hl, tile, C);
  if (hl < 5) tile[86 + hl] = 0.0;
  sweep_prefetch16(ring, gslots, N - 1, hl, N);
  sweep_prefetch16(ring, gslots, N - 2, hl, N);
  sweep_prefetch16(ring, gslots, N - 3, hl, N);
  __syncwarp();
  for (int k = N - 1; k >= 0; --k) {
    sweep_prefetch16(ring, gslots, k - 3, hl, N);
    cp_async_wait<3>();
    __syncwarp();
    const double* slot = ring + (k & (SWEEP_DEPTH - 1)) * GSTRIDE;
    double* gk = gslots + (size_t)k * GSTRIDE;
    const bool wr = run && L.ok;
    PS::kl_step1(L, hl, slot, tile, wr ? gk + GSTRIDE : dump);
    __syncwarp();
    PS::kl_step2(L, hl, slot, tile);
    __syncwarp();
    PS::kl_step3(L, hl, wr ? gk : dump, tile);
    __syncwarp();
    if (!__any_sync(FULL, run && L.ok)) break;      // both halves have nothing (left) to do
  }
  cp_async_wait<0>();
  const bool good = run && L.ok;
  if (!__any_sync(FULL, good)) return 0;
  int ok = good ? 1 : 0;
  if (hl == IT) { tile[63] = L.Prow[IT]; tile[64] = L.pl; }
  __threadfence();
  __syncwarp();
  double dt = 0.0;
  if (!Pp.fix_time) {
    double ptt = tile[63];
    if (!(ptt > 0.0)) { ok = 0; ptt = 1e300; }
    dt = -tile[64] / ptt;
  }
  if (hl == 0 && good) Sgl->dt = dt;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0, s5 = 0.0;
  const int ur = hl & 1, xr = hl & 3;
  const double selx = (xr == 0) ? 1.0 : 0.0, sely = (xr == 1) ? 1.0 : 0.0;
  double* const dxw = W + (size_t)(Lay.dX + xr) * Lay.NSP;
  double* const duw = W + (size_t)(Lay.dDE + ur) * Lay.NSP;
  __syncwarp();
  sweep_prefetch16(ring, gslots, 0, hl, N); sweep_prefetch16(ring, gslots, 1, hl, N); sweep_prefetch16(ring, gslots, 2, hl, N);
  for (int k = 0; k < N; ++k) {
    sweep_prefetch16(ring, gslots, k + 3, hl, N);
    cp_async_wait<3>();
    __syncwarp();
    const double* const slot = ring + (k & (SWEEP_DEPTH - 1)) * GSTRIDE;
    const double* const kr = slot + RK + ur * NSV;
    double u = slot[RK + 14 + ur] + kr[0] * s0 + kr[1] * s1 + kr[2] * s2 + kr[3] * s3 + kr[4] * s4 + kr[5] * s5 + kr[6] * dt;
    const double u0 = __shfl_sync(FULL, u, base + 0), u1 = __shfl_sync(FULL, u, base + 1);
    const double* const dr = slot + RDYN + 5 * xr;
    double sn = slot[RR4 + xr] + selx * s0 + sely * s1 + dr[0] * s2 + dr[1] * s3 + dr[2] * dt + dr[3] * u0 + dr[4] * u1;
    if (hl < 2 && good) duw[k] = u;
    s0 = __shfl_sync(FULL, sn, base + 0); s1 = __shfl_sync(FULL, sn, base + 1); s2 = __shfl_sync(FULL, sn, base + 2);
    s3 = __shfl_sync(FULL, sn, base + 3);
    s4 = u0; s5 = u1;
    if (hl < 4 && good) {
      if (k + 1 < N) dxw[k + 1] = sn;
      else Sgl->eN[xr] = sn;
    }
    __syncwarp();
  }
  cp_async_wait<0>();
  if (good) {
    if (hl < 4) { dxw[0] = 0.0; dxw[N] = 0.0; }
    if (hl < 2) duw[N] = 0.0;
  }
  return ok;
}

constexpr int SWEEP_WARPS = 4;
template <int VM, bool SDV>
__global__ void __launch_bounds__(32 * SWEEP_WARPS)
k_pk_sweep(const __grid_constant__ ParkProblem P, const __grid_constant__ IpmOpts O, const PkLay L,
           double* __restrict__ Wall, double* __restrict__ slots, ProbState* __restrict__ Sg,
           const int* __restrict__ act, const int* __restrict__ n_act) {
  __shared__ __align__(16) double s_w[SWEEP_WARPS][2 * SWEEP_HALF_DOUBLES];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, half = lane >> 4, hl = lane & 15;
  const int n = *n_act;
  const int first = (blockIdx.x * SWEEP_WARPS + warp) * 2;
  if (first >= n) return;
  const int idx = first + half;
  const bool valid = idx < n;
  const int b = act[valid ? idx : first];              // a half without a problem reads its partner's (valid) memory
  ProbState* S = Sg + b;
  const bool active = valid && S->phase == PH_KKT;
  if (!__any_sync(0xffffffffu, active)) return;
  const int NS = P.N + 1;
  const bool run = active && S->ok != 0;
  double* hw = s_w[warp] + half * SWEEP_HALF_DOUBLES;
  int ok = 0;
  if (__any_sync(0xffffffffu, run))
    ok = pk_sweep_pair<VM, SDV>(P, O, L, Wall + (size_t)b * L.total * L.NSP, slots + (size_t)b * NS * GSTRIDE, S, run, hw,
                                hw + SWEEP_DEPTH * GSTRIDE, hw + SWEEP_DEPTH * GSTRIDE + 144);
  __syncwarp();
  if (hl == 0 && active) PhasedDriver<ParkSolver<VM, SDV> >::phase_B_serial(*S, O, run ? ok : 0);
}

// ------------------------------------------------------------------------------------------------------------
// K_rblk: step recovery of the OBCA blocks as a flat streaming kernel (same thread mapping as k_pk_block): reads the local
// factor written by k_pk_block and the pose step of the sweep, writes the steps of (lambda, mu, sl) and the new multipliers
// of the block's rows into the workspace, and a 5-double record (tightest primal / dual fractions to the boundary, barrier
// directional derivative) for k_pk_phaseC.
// ------------------------------------------------------------------------------------------------------------
template <int VM, bool SDV>
__global__ void __launch_bounds__(128, OBCA_MINB_BLK)
k_pk_rblock(const __grid_constant__ ParkProblem P, const PkLay L, double* __restrict__ Wall, double* __restrict__ BOall,
            const ProbState* __restrict__ Sg, const int* __restrict__ act, const int* __restrict__ n_act) {
  const int NS = P.N + 1, per = P.nOb * NS;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int a = (int)(idx / per);
  if (a >= *n_act) return;
  const int rem = (int)(idx - (long long)a * per);
  const int j = rem / NS, k = rem - j * NS;
  const int b = act[a];
  const ProbState* S = Sg + b;
  if (S->phase != PH_RECOVER) return;
  const double mu_b = S->mu, dw = S->dw;
  PkCtx C;
  C.W = Wall + (size_t)b * L.total * L.NSP;
  C.bo = BOall + (size_t)b * P.nOb * BO_N * L.NSP;
  const double X = WA(X, k), Y = WA(Y, k), ps = WA(PS, k);
  const double dX = WA(dX, k), dY = WA(dY, k), dP = WA(dPS, k);
  double sn_, cs_;
  sincos(ps, &sn_, &cs_);
  typename ParkSolver<VM, SDV>::RBlockOut rb;
  ParkSolver<VM, SDV>::block_recover(C, k, j, X, Y, cs_, sn_, dX, dY, dP, mu_b, dw, rb);
  ParkSolver<VM, SDV>::rblock_store(C, k, j, rb);
}

// ------------------------------------------------------------------------------------------------------------
// K_C: one CTA per active problem
// ------------------------------------------------------------------------------------------------------------
#ifndef OBCA_MINB_C
#define OBCA_MINB_C 4
#endif
template <int VM, bool SDV>
__global__ void __launch_bounds__(128, OBCA_MINB_C)
k_pk_phaseC(const __grid_constant__ ParkProblem P, const __grid_constant__ IpmOpts O, const PkLay L, const BatchPtrs bp,
            double* __restrict__ Wall, double* __restrict__ slots, ProbState* __restrict__ Sg, double* __restrict__ BOall,
            const int* __restrict__ act, const int* __restrict__ n_act) {
  __shared__ ProbState S;
  __shared__ double s_red[4 * 12];
  if ((int)blockIdx.x >= *n_act) return;
  const int b = act[blockIdx.x];
  if (Sg[b].phase != PH_RECOVER) return;
  const int NS = P.N + 1;
  state_load(S, Sg + b);
  __shared__ PkCtx C; __shared__ PkOutputs out;
  if (threadIdx.x == 0) {
    pk_make_ctx<VM, SDV>(C, out, P, O, L, bp, b, Wall + (size_t)b * L.total * L.NSP);
    C.ric = nullptr; C.pp = slots + (size_t)b * NS * GSTRIDE; C.pps = GSTRIDE; C.red_scratch = s_red; C.tile = nullptr; C.S = &S;
    C.bo = BOall + (size_t)b * P.nOb * BO_N * L.NSP;
  }
  __syncthreads();
  PhasedDriver<ParkSolver<VM, SDV> >::template phase_C<true>(C);
  state_store(Sg + b, S);
}

// ------------------------------------------------------------------------------------------------------------
// tail / small-batch kernel: persistent CTAs pull problems (fresh, or handed over by the phase kernels at a round
// boundary) and run the three phases in a loop with the stage slots in shared memory.
// ------------------------------------------------------------------------------------------------------------
#ifndef OBCA_MIN_BLOCKS
#define OBCA_MIN_BLOCKS 3
#endif
template <int VM, bool SDV>
__global__ void __launch_bounds__(128, OBCA_MIN_BLOCKS)
k_pk_tail(const __grid_constant__ ParkProblem P, const __grid_constant__ IpmOpts O, const PkLay L, const BatchPtrs bp,
          double* __restrict__ Wall, ProbState* __restrict__ Sg, const int* __restrict__ act, const int* __restrict__ n_act,
          int* __restrict__ counter, int fresh) {
  extern __shared__ double s_ric[];
  __shared__ ProbState S;
  __shared__ double s_tile[144];
  __shared__ double s_red[4 * 12];
  __shared__ PkFinalScratch s_fin;
  __shared__ int s_i;
  __shared__ PkCtx C; __shared__ PkOutputs out;
  typedef ParkSolver<VM, SDV> PS;
  const int n = *n_act;
  for (;;) {
    if (threadIdx.x == 0) s_i = atomicAdd(counter, 1);
    __syncthreads();
    const int i = s_i;
    if (i >= n) break;
    const int b = fresh ? i : act[i];
    if (fresh) state_fresh(S); else state_load(S, Sg + b);
    if (threadIdx.x == 0) {
      S.prof[7] = 0;
      pk_make_ctx<VM, SDV>(C, out, P, O, L, bp, b, Wall + (size_t)b * L.total * L.NSP);
      C.ric = s_ric; C.pp = s_ric; C.pps = RSTRIDE; C.red_scratch = s_red; C.tile = s_tile; C.S = &S;
    }
    __syncthreads();
    for (;;) {
      pk_step_A<VM, SDV, false>(C, out, bp, b, s_fin);
      if (S.phase == PH_DONE) break;
      if (threadIdx.x < 32) {
        int ok = S.ok;
        if (ok) ok = PS::kkt_solve_warp(C, s_tile);
        __syncwarp();
        if (threadIdx.x == 0) PhasedDriver<PS>::phase_B_serial(S, O, ok);
      }
      __syncthreads();
      if (S.phase == PH_RECOVER) PhasedDriver<PS>::template phase_C<false>(C);
    }
    if (threadIdx.x == 0 && bp.prof) atomicAdd(bp.prof + 6, (unsigned long long)S.prof[7]);
    __syncthreads();
  }
}

#endif  // __CUDACC__

}  // namespace obca

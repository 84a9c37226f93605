// obca_lib.cu -- sm_100a kernels + the C-ABI of include/obca.h.
//
// Kernels
//   obca_phased.cuh          the parking solver: phase-split rounds (k_pk_eval, k_pk_sweep, k_pk_step) over all active
//                            problems + the persistent tail kernel k_pk_tail
//   k_quad_solve<SDV>        quadcopter model (config 4): persistent CTA per problem, block-cooperative KKT sweep on FP64 tensor cores (DMMA)
//   k_dualws<VM>             K2: one thread per (problem, stage, obstacle) micro interior-point solve.
//   k_check                  K5: ParkingConstraints twin + strict audit, one CTA per problem.
// There is NO CPU fallback: every compute entry point returns OBCA_ERR_NO_DEVICE without a CUDA device.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/obca.h"
#include "obca_check.cuh"
#include "obca_dualws.cuh"
#include "obca_eval.cuh"
#include "obca_host.h"
#include "obca_phased.cuh"

using namespace obca;

// ---------------------------------------------------------------------------------------------------------
// device-side batch description
// ---------------------------------------------------------------------------------------------------------
template <int VM>
__global__ void k_dualws(const __grid_constant__ ParkProblem P, int B, const double* __restrict__ rx,
                         const double* __restrict__ ry, const double* __restrict__ ryaw, double tol, int max_iter,
                         double* __restrict__ lp, double* __restrict__ np, double* __restrict__ dd) {
  const int NS = P.N + 1, nOb = P.nOb, V = P.V;
  const size_t total = (size_t)B * nOb * NS;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % NS);
    const int j = (int)((idx / NS) % nOb);
    const size_t b = idx / ((size_t)NS * nOb);
    ObsRows<VM> R;
    R.v = P.vOb[j];
#pragma unroll
    for (int i = 0; i < VM; ++i) {
      const bool on = i < R.v;
      const int r = P.voff[j] + (on ? i : 0);
      R.a1[i] = on ? P.A[r][0] : 0.0; R.a2[i] = on ? P.A[r][1] : 0.0; R.bb[i] = on ? P.b[r] : 0.0;
    }
    double lam[VM], mu[4], d;
    const size_t o = (size_t)NS * b + k;
    dualws_solve<VM>(P, R, rx[o], ry[o], ryaw[o], tol, max_iter, lam, mu, &d);
#pragma unroll
    for (int i = 0; i < VM; ++i)
      if (i < R.v) lp[(size_t)V * NS * b + (size_t)(P.voff[j] + i) * NS + k] = lam[i];
#pragma unroll
    for (int m = 0; m < 4; ++m) np[(size_t)4 * nOb * NS * b + (size_t)(4 * j + m) * NS + k] = mu[m];
    if (dd) dd[(size_t)nOb * NS * b + (size_t)j * NS + k] = d;
  }
}

// K1 stand-alone (obca_eval.cuh): one thread per (problem, stage); consecutive threads = consecutive stages, so every
// warp streams contiguous slices of the stacked (x, u, l, n, sl, y) arrays.
template <int VM, bool SDV>
__global__ void __launch_bounds__(128)
k_parking_eval(const __grid_constant__ ParkProblem P, int B, EvalIn in, EvalOut out) {
  const int N = P.N, NS = N + 1, nOb = P.nOb, V = P.V;
  const size_t m = eval_m(P), n = eval_n(P);
  const size_t total = (size_t)B * NS;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t b = idx / NS;
    const int k = (int)(idx - b * NS);
    EvalIn ib;
    ib.x0 = in.x0 + 4 * b; ib.xF = in.xF + 4 * b; ib.rx = in.rx + NS * b; ib.ry = in.ry + NS * b; ib.ryaw = in.ryaw + NS * b;
    ib.xp = in.xp + (size_t)4 * NS * b; ib.up = in.up + (size_t)2 * N * b; ib.ts = in.ts ? in.ts + NS * b : nullptr;
    ib.lp = in.lp + (size_t)V * NS * b; ib.np = in.np + (size_t)4 * nOb * NS * b;
    ib.sl = in.sl ? in.sl + (size_t)nOb * NS * b : nullptr;
    ib.y = in.y ? in.y + m * b : nullptr;
    EvalOut ob;
    ob.c = out.c + m * b; ob.gradL = out.gradL + n * b; ob.fk = out.fk + NS * b;
    eval_stage<VM, SDV>(P, k, ib, ob);
  }
}

__global__ void k_check(const __grid_constant__ ParkProblem P, int B, const double* __restrict__ x0,
                        const double* __restrict__ xF, const double* __restrict__ x, const double* __restrict__ u,
                        const double* __restrict__ l, const double* __restrict__ n, const double* __restrict__ ts,
                        const double* __restrict__ sl, int sd, int* __restrict__ feasible, int* __restrict__ e_out,
                        int* __restrict__ strict_out) {
  __shared__ ChkPart s_c[2][4];
  const int N = P.N, NS = N + 1, V = P.V, nOb = P.nOb;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const double* xb = x + (size_t)4 * NS * b; const double* ub = u + (size_t)2 * N * b;
    const double* lb = l + (size_t)V * NS * b; const double* nb = n + (size_t)4 * nOb * NS * b;
    const double* tb = ts + (size_t)NS * b; const double* sb = sl ? sl + (size_t)nOb * NS * b : nullptr;
    ChkPart c[2];
    chk_init(c[0]); chk_init(c[1]);
    for (int k = threadIdx.x; k < NS; k += blockDim.x) {
      for (int st = 0; st < 2; ++st) {
        ChkPart ck;
        check_stage(P, k, x0 + 4 * (size_t)b, xF + 4 * (size_t)b, xb, ub, lb, nb, tb, sb, sd, st, ck);
        chk_merge(c[st], ck);
      }
    }
    for (int st = 0; st < 2; ++st) {
      for (int off = 16; off > 0; off >>= 1) {
        ChkPart o;
#pragma unroll
        for (int i = 0; i < 5; ++i) o.c0[i] = __shfl_down_sync(0xffffffffu, c[st].c0[i], off);
        o.c1 = __shfl_down_sync(0xffffffffu, c[st].c1, off); o.c2 = __shfl_down_sync(0xffffffffu, c[st].c2, off);
        o.c3 = __shfl_down_sync(0xffffffffu, c[st].c3, off); o.c4 = __shfl_down_sync(0xffffffffu, c[st].c4, off);
        o.c5 = __shfl_down_sync(0xffffffffu, c[st].c5, off); o.c6 = __shfl_down_sync(0xffffffffu, c[st].c6, off);
        o.sbox = __shfl_down_sync(0xffffffffu, c[st].sbox, off);
        chk_merge(c[st], o);
      }
      if ((threadIdx.x & 31) == 0) s_c[st][threadIdx.x >> 5] = c[st];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int st = 0; st < 2; ++st) {
        ChkPart m = s_c[st][0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) chk_merge(m, s_c[st][w]);
        int e[7];
        const int f = check_finish(P, m, tb, st, 5e-5, e);
        if (st == 0) {
          feasible[b] = f;
          if (e_out) for (int i = 0; i < 7; ++i) e_out[7 * (size_t)b + i] = e[i];
        } else if (strict_out) strict_out[b] = f;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------
// quadcopter (obca_quad.cuh): persistent solver kernel + constrSatisfaction twin
// ---------------------------------------------------------------------------------------------------------
struct QBatchPtrs {
  const double *x0, *xF, *xWS;
  double timeWS;
  double *xp, *up, *ts, *lp, *slack;
  int *exitflag, *iters;
  double* kkt_err;
  int B;
  unsigned long long* prof;   // 8 per-phase cycle counters summed over the batch (see obca_last_profile)
};

#ifndef OBCA_MINB_Q
#define OBCA_MINB_Q 2
#endif
template <bool SDV>
__global__ void __launch_bounds__(128, OBCA_MINB_Q)
k_quad_solve(const __grid_constant__ QuadProblem P, const __grid_constant__ IpmOpts O, const QLay L, const QBatchPtrs bp,
             double* __restrict__ Wall, int* __restrict__ counter) {
  extern __shared__ double s_kkt[];      // QuadSolver::smem_doubles(N): block-cooperative KKT sweep
  __shared__ ProbState S;
  __shared__ double s_red[4 * 12];
  __shared__ int s_b;
  __shared__ double s_sum[4];
  const int N = P.N, NS = N + 1;
  double* W = Wall + (size_t)blockIdx.x * L.total * L.NSP;
  for (;;) {
    if (threadIdx.x == 0) s_b = atomicAdd(counter, 1);
    __syncthreads();
    const int b = s_b;
    if (b >= bp.B) break;
    QCtx C;
    C.P = &P; C.O = &O; C.L = L; C.W = W; C.red_scratch = s_red; C.tile = s_kkt; C.S = &S;
    C.in.x0 = bp.x0 + 12 * (size_t)b; C.in.xF = bp.xF + 12 * (size_t)b; C.in.xWS = bp.xWS + (size_t)12 * NS * b;
    C.in.timeWS = bp.timeWS;
    QOutputs out;
    out.xp = bp.xp + (size_t)12 * NS * b; out.up = bp.up + (size_t)4 * N * b; out.ts = bp.ts + (size_t)NS * b;
    out.lp = bp.lp + (size_t)30 * NS * b; out.slack = bp.slack ? bp.slack + (size_t)5 * NS * b : nullptr;
    IpmDriver<QuadSolver<SDV> >::solve(C, 0);      // flag = 1 in the reference: a single attempt (QuadcopterSignedDist.jl:227-235)
    __syncthreads();
    double ssum = 0.0;
    for (int k = threadIdx.x; k < NS; k += blockDim.x) {
      QuadSolver<SDV>::store_stage(C, k, out);
      if (SDV) for (int j = 0; j < QNOB; ++j) ssum += W[(size_t)(L.SLK + j) * L.NSP + k];
    }
    for (int off = 16; off > 0; off >>= 1) ssum += __shfl_down_sync(0xffffffffu, ssum, off);
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = ssum;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_sum[w];
      int ef = S.status == 1 ? 1 : 0;
      if (SDV && ef == 1 && tot > 1e-3) ef = 2;      // sum-slack gate, QuadcopterSignedDist.jl:283-288
      bp.exitflag[b] = ef; bp.iters[b] = S.iters; bp.kkt_err[b] = S.e0;
      if (bp.prof)
        for (int i = 0; i < 8; ++i) atomicAdd(bp.prof + i, (unsigned long long)S.prof[i]);
    }
    __syncthreads();
  }
}

__global__ void k_quad_check(const __grid_constant__ QuadProblem P, int B, const double* __restrict__ x0,
                             const double* __restrict__ xF, const double* __restrict__ x, const double* __restrict__ u,
                             const double* __restrict__ ts, const double* __restrict__ lam, int* __restrict__ feasible,
                             double* __restrict__ worst_out) {
  __shared__ double s_w[4];
  const int N = P.N, NS = N + 1;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    double w = 0.0;
    for (int k = threadIdx.x; k < NS; k += blockDim.x)
      w = fmax(w, quad_check_stage(P, k, x0 + 12 * (size_t)b, xF + 12 * (size_t)b, x + (size_t)12 * NS * b, u + (size_t)4 * N * b,
                                   ts + (size_t)NS * b, lam + (size_t)30 * NS * b));
    for (int off = 16; off > 0; off >>= 1) w = fmax(w, __shfl_down_sync(0xffffffffu, w, off));
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = w;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int i = 1; i < (int)(blockDim.x >> 5); ++i) w = fmax(w, s_w[i]);
      feasible[b] = w <= 1e-3 ? 1 : 0;
      if (worst_out) worst_out[b] = w;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static void set_err(const std::string& s) { g_err = s; }
#define CK(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) {                                                                         \
      set_err(std::string(#call) + ": " + cudaGetErrorString(e_));                                   \
      return OBCA_ERR_CUDA;                                                                          \
    }                                                                                                \
  } while (0)

struct DevCtx {
  bool init = false;
  cudaStream_t st = nullptr;
  cudaEvent_t ev0 = nullptr, evm = nullptr, ev1 = nullptr;
  int sms = 0;
  double* W = nullptr; size_t Wbytes = 0;
  int* counter = nullptr;
  unsigned long long* prof = nullptr;
  unsigned long long prof_host[8] = {0};
  char* stage = nullptr; size_t stage_bytes = 0;   // device staging for the host-pointer API
  double* wsd = nullptr; size_t wsd_bytes = 0;     // warm-start duals of DualMultWS when the caller passes lWS = nWS = NULL
  // phase-split driver (obca_phased.cuh)
  double* slots = nullptr; size_t slots_bytes = 0;  // stage slots of every problem, B x (N+1) x RSTRIDE
  char* pstate = nullptr; size_t pstate_bytes = 0;  // ProbState per problem
  int* act = nullptr; size_t act_bytes = 0;         // two active lists
  int* ncnt = nullptr;                              // device: three rotating active counts, [3] tail work counter
  int* h_n = nullptr;                               // pinned ring of active counts read back per round
  cudaEvent_t evr[16] = {nullptr};
  std::vector<cudaEvent_t> tev;                     // OBCA_PHASE_TIMING=1: pool of timing events
  int last_rounds = 0, last_tail = 0;
  double phase_ms[5] = {0, 0, 0, 0, 0};             // OBCA_PHASE_TIMING=1: summed event times of eval, sweep, step, tail; [4] DualMultWS
  double last_dualws_s = 0.0, last_solve_s = 0.0;
  std::recursive_mutex mu;
};
static DevCtx g_dev[64];
static std::mutex g_init_mu;

static int get_ctx(int dev, DevCtx** out) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { set_err("no CUDA device visible"); return OBCA_ERR_NO_DEVICE; }
  if (dev < 0 || dev >= n || dev >= 64) { set_err("bad device ordinal"); return OBCA_ERR_ARG; }
  DevCtx& c = g_dev[dev];
  CK(cudaSetDevice(dev));
  std::lock_guard<std::mutex> lk(g_init_mu);
  if (!c.init) {
    CK(cudaStreamCreateWithFlags(&c.st, cudaStreamNonBlocking));
    CK(cudaEventCreate(&c.ev0)); CK(cudaEventCreate(&c.evm)); CK(cudaEventCreate(&c.ev1));
    cudaDeviceProp pr;
    CK(cudaGetDeviceProperties(&pr, dev));
    c.sms = pr.multiProcessorCount;
    CK(cudaMalloc(&c.counter, sizeof(int)));
    CK(cudaMalloc(&c.prof, 8 * sizeof(unsigned long long)));
    CK(cudaMalloc(&c.ncnt, 4 * sizeof(int)));
    CK(cudaMallocHost(&c.h_n, 16 * sizeof(int)));
    for (int i = 0; i < 16; ++i) CK(cudaEventCreateWithFlags(&c.evr[i], cudaEventDisableTiming));
    c.init = true;
  }
  *out = &c;
  return 0;
}
static int ensure(void** p, size_t* have, size_t need) {
  if (*have >= need) return 0;
  if (*p) cudaFree(*p);
  *p = nullptr; *have = 0;
  CK(cudaMalloc(p, need));
  *have = need;
  return 0;
}

static IpmOpts to_ipm(const obca_opts* o) {
  IpmOpts r = default_opts();
  if (!o) return r;
  r.tol = o->tol; r.max_iter = o->max_iter; r.mu_init = o->mu_init; r.mu_min = o->mu_min;
  r.kappa_eps = o->kappa_eps; r.kappa_mu = o->kappa_mu; r.theta_mu = o->theta_mu; r.tau_min = o->tau_min;
  r.kappa1 = o->kappa1; r.kappa2 = o->kappa2; r.kappa_sigma = o->kappa_sigma; r.s_max = o->s_max;
  r.dual_inf_tol = o->dual_inf_tol; r.constr_viol_tol = o->constr_viol_tol; r.compl_inf_tol = o->compl_inf_tol;
  r.dw_min = o->dw_min; r.dw_first = o->dw_first; r.dw_max = o->dw_max; r.kw_minus = o->kw_minus;
  r.kw_plus = o->kw_plus; r.kw_plus_first = o->kw_plus_first;
  r.gamma_theta = o->gamma_theta; r.gamma_phi = o->gamma_phi; r.delta = o->delta; r.s_theta = o->s_theta;
  r.s_phi = o->s_phi; r.eta_phi = o->eta_phi; r.gamma_alpha = o->gamma_alpha; r.max_backtrack = o->max_backtrack;
  r.dc = o->dc; r.max_kick = o->max_kick; r.quad_dual_ws = o->quad_dual_ws;
  return r;
}

template <int VM>
static int launch_dualws(DevCtx& c, const ParkProblem& P, int B, const double* rx, const double* ry,
                         const double* ryaw, double* lp, double* np, double* dd) {
  const size_t total = (size_t)B * P.nOb * (P.N + 1);
  const int threads = 128;
  const int blocks = (int)((total + threads - 1) / threads);
  k_dualws<VM><<<blocks, threads, 0, c.st>>>(P, B, rx, ry, ryaw, 1e-5, 100, lp, np, dd);
  CK(cudaGetLastError());
  return 0;
}
static int run_dualws(DevCtx& c, const ParkProblem& P, int B, const double* rx, const double* ry, const double* ryaw,
                      double* lp, double* np, double* dd) {
  return max_vob(P) <= 2 ? launch_dualws<2>(c, P, B, rx, ry, ryaw, lp, np, dd)
                         : launch_dualws<4>(c, P, B, rx, ry, ryaw, lp, np, dd);
}

// ---------------------------------------------------------------------------------------------------------
// phase-split driver (obca_phased.cuh): rounds of [K_A, K_B, K_C] over the active problems, then the tail kernel
// ---------------------------------------------------------------------------------------------------------
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

template <int VM, bool SDV>
static int launch_phased(DevCtx& c, const ParkProblem& P, const IpmOpts& O, const BatchPtrs& bp, int mode) {
  PkLay L = make_layout(P, LocalDims<VM, SDV>::NFAC);
  const int B = bp.B, NS = P.N + 1;
  if (NS > 128) { set_err("horizon too long for this build (N+1 <= 128)"); return OBCA_ERR_UNSUPPORTED; }
  const size_t smem_slots = (size_t)NS * RSTRIDE * sizeof(double);
  const size_t smem_step = (size_t)(L.dRS + 1 - L.dLAM) * L.NSP * sizeof(double);
  CK(cudaFuncSetAttribute(k_pk_eval<VM, SDV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_slots));
  CK(cudaFuncSetAttribute(k_pk_tail<VM, SDV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_slots));
  CK(cudaFuncSetAttribute(k_pk_step<VM, SDV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_step));
  const size_t smem_sweep = (size_t)SWEEP_WARPS * SWEEP_WARP_DOUBLES * sizeof(double);
  CK(cudaFuncSetAttribute(k_pk_sweep<VM, SDV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sweep));
  int occ_t = 0, occ_e = 0, occ_s = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_t, k_pk_tail<VM, SDV>, PK_THREADS, smem_slots));
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_e, k_pk_eval<VM, SDV>, PK_THREADS, smem_slots));
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_s, k_pk_step<VM, SDV>, PK_THREADS, smem_step));
  if (occ_t < 1) occ_t = 1;
  if (occ_e < 1) occ_e = 1;
  if (occ_s < 1) occ_s = 1;
  {   // development: fewer resident CTAs per SM than the occupancy allows (how much do co-resident CTAs slow each other down?)
    const int oe = env_int("OBCA_EVAL_OCC", occ_e), os = env_int("OBCA_STEP_OCC", occ_s);
    if (oe >= 1 && oe < occ_e) occ_e = oe;
    if (os >= 1 && os < occ_s) occ_s = os;
  }
  const int tail_cap = c.sms * occ_t, eval_cap = c.sms * occ_e, step_cap = c.sms * occ_s;
  int rc = ensure((void**)&c.W, &c.Wbytes, (size_t)B * L.total * L.NSP * sizeof(double));
  if (rc) return rc;
  rc = ensure((void**)&c.pstate, &c.pstate_bytes, (size_t)B * sizeof(ProbState));
  if (rc) return rc;
  ProbState* Sg = (ProbState*)c.pstate;
  cudaStream_t st = c.st;
  // mode 1: tail kernel only; mode 2: rounds only (until the active set is empty, or OBCA_TAIL_THRESH if set); auto: rounds for
  // large batches until the active set has shrunk to the hand-over point, then the tail kernel.  Hand-over point: measured on
  // B200 (config 2, B = 4096, this build) the solve time is flat (113-116 ms) between 1200 and 2500 remaining problems and
  // 4-15 % higher outside (rounds only 133 ms, tail kernel only 133 ms): lock-step rounds do 1.65x the evaluations per second
  // while all SMs are busy, the persistent kernel finishes the stragglers without a launch-bound round per iteration.
  const int thresh = env_int("OBCA_TAIL_THRESH", mode == 2 ? 0 : (B / 2 > tail_cap ? B / 2 : tail_cap));
  const bool rounds = mode == 2 || (mode == 0 && B > 2 * tail_cap);
  c.last_rounds = 0; c.last_tail = B;
  int h_init[4] = {B, 0, 0, 0};
  CK(cudaMemcpyAsync(c.ncnt, h_init, 4 * sizeof(int), cudaMemcpyHostToDevice, st));
  if (!rounds) {
    const int grid = B < tail_cap ? B : tail_cap;
    k_pk_tail<VM, SDV><<<grid, PK_THREADS, smem_slots, st>>>(P, O, L, bp, c.W, Sg, nullptr, c.ncnt, c.ncnt + 3, 1);
    CK(cudaGetLastError());
    return 0;
  }
  rc = ensure((void**)&c.slots, &c.slots_bytes, (size_t)B * NS * RSTRIDE * sizeof(double));
  if (rc) return rc;
  rc = ensure((void**)&c.act, &c.act_bytes, 2 * (size_t)B * sizeof(int));
  if (rc) return rc;
  int* act[2] = {c.act, c.act + B};
  int n_bound = B, done_r = 0, r = 0;
  const int max_rounds = 8 * (O.max_iter + 8);
  const bool timing = env_int("OBCA_PHASE_TIMING", 0) != 0;      // development / bench: per-kernel event times, summed per solve
  size_t n_ev = 0;
  auto mark = [&]() {
    if (!timing) return;
    if (n_ev == c.tev.size()) { cudaEvent_t e; cudaEventCreate(&e); c.tev.push_back(e); }
    cudaEventRecord(c.tev[n_ev++], st);
  };
  auto cap = [](int n, int m) { return n < m ? (n > 0 ? n : 1) : m; };
  // one round = three launches + the read-back of the active count; grids from an upper bound nb of the active count
  auto launch_round = [&](int rr, int nb, bool readback) -> int {
    const int* a_in = act[rr & 1]; int* a_out = act[(rr + 1) & 1];
    int* n_in = c.ncnt + (rr % 3); int* n_out = c.ncnt + ((rr + 1) % 3); int* n_zero = c.ncnt + ((rr + 2) % 3);
    mark();
    k_pk_eval<VM, SDV><<<cap(nb, eval_cap), PK_THREADS, smem_slots, st>>>(P, O, L, bp, c.W, c.slots, Sg, a_in, n_in, a_out, n_out, rr == 0 ? 1 : 0);
    mark();
    k_pk_sweep<VM, SDV><<<(nb + SWEEP_WARPS - 1) / SWEEP_WARPS, 32 * SWEEP_WARPS, smem_sweep, st>>>(P, O, L, c.W, c.slots, Sg, a_out, n_out, n_zero);
    mark();
    k_pk_step<VM, SDV><<<cap(nb, step_cap), PK_THREADS, smem_step, st>>>(P, O, L, bp, c.W, c.slots, Sg, a_out, n_out);
    mark();
    CK(cudaGetLastError());
    if (readback) CK(cudaMemcpyAsync(c.h_n + (rr & 15), n_out, sizeof(int), cudaMemcpyDeviceToHost, st));
    return 0;
  };
  // Rounds as a CUDA graph: the buffer pattern of a round (active lists: period 2, counters: period 3) repeats every 6 rounds, so
  // rounds 1..6 are captured once per call and the graph is replayed -- about 14 driver calls per solve instead of ~420
  // (3 launches + copy + event per round), which is what paces the host when 8 ranks share one box.  The grids inside the graph
  // are those of the count at capture time (an upper bound later on: surplus CTAs / warps exit at once); the host looks at the
  // active count once per replay, one replay behind, so the hand-over point is overshot by up to 12 rounds (the solve time is
  // flat there, see above).  OBCA_GRAPH=0 or OBCA_PHASE_TIMING=1: plain launches, count read back every round.
  const bool use_graph = !timing && env_int("OBCA_GRAPH", 1) != 0;
  if (use_graph) {
    rc = launch_round(0, B, true);
    if (rc) return rc;
    r = 1;
    cudaGraph_t graph = nullptr; cudaGraphExec_t gexec = nullptr;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    for (int q = 0; q < 6 && rc == 0; ++q) rc = launch_round(1 + q, B, q == 5);
    cudaError_t ce = cudaStreamEndCapture(st, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    CK(ce);
    CK(cudaGraphInstantiate(&gexec, graph, 0));
    int g = 0;
    for (; r < max_rounds; ++g) {
      CK(cudaGraphLaunch(gexec, st));
      CK(cudaEventRecord(c.evr[g & 1], st));
      r += 6;
      if (g >= 1) {
        CK(cudaEventSynchronize(c.evr[(g - 1) & 1]));
        n_bound = c.h_n[(1 + 5) & 15];      // count after the last round of a replay (the replay in flight may refresh it: still a bound)
        if (n_bound <= thresh) break;
      }
    }
    CK(cudaStreamSynchronize(st));
    n_bound = c.h_n[(1 + 5) & 15];
    cudaGraphExecDestroy(gexec); cudaGraphDestroy(graph);
  } else {
    for (; r < max_rounds; ++r) {
      rc = launch_round(r, n_bound, true);
      if (rc) return rc;
      CK(cudaEventRecord(c.evr[r & 15], st));
      // the active set only shrinks: the count of any completed round bounds every later round
      if (r - done_r >= 8) CK(cudaEventSynchronize(c.evr[done_r & 15]));
      while (done_r <= r && cudaEventQuery(c.evr[done_r & 15]) == cudaSuccess) { n_bound = c.h_n[done_r & 15]; ++done_r; }
      if (n_bound <= thresh) {
        CK(cudaStreamSynchronize(st));
        n_bound = c.h_n[r & 15];
        ++r;
        break;
      }
    }
  }
  c.last_rounds = r; c.last_tail = n_bound;
  if (timing) {
    CK(cudaStreamSynchronize(st));
    double t[3] = {0, 0, 0};      // four marks per round: | eval | sweep | step |
    for (size_t i = 0; i + 3 < n_ev; i += 4)
      for (int j = 0; j < 3; ++j) { float ms = 0.f; cudaEventElapsedTime(&ms, c.tev[i + j], c.tev[i + j + 1]); t[j] += ms; }
    for (int j = 0; j < 3; ++j) c.phase_ms[j] += t[j];
  }
  if (n_bound > 0) {
    // after r rounds the active list is act[r & 1] with count ncnt[r % 3]
    const int grid = n_bound < tail_cap ? n_bound : tail_cap;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (timing) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, st); }
    k_pk_tail<VM, SDV><<<grid, PK_THREADS, smem_slots, st>>>(P, O, L, bp, c.W, Sg, act[r & 1], c.ncnt + (r % 3), c.ncnt + 3, 0);
    CK(cudaGetLastError());
    if (timing) {
      cudaEventRecord(e1, st); cudaEventSynchronize(e1);
      float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1); c.phase_ms[3] += ms;
      cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
  }
  return 0;
}

static int solve_dev_impl(DevCtx& c, const ParkProblem& P, const IpmOpts& O, BatchPtrs bp, double* seconds) {
  const int vm = max_vob(P) <= 2 ? 2 : 4;
  {  // problem, options and workspace layout -> __constant__ memory (stream-ordered with the kernels below)
    const PkLay Lc = make_layout(P, nfac_for(P));
    CK(cudaMemcpyToSymbolAsync(c_pkP, &P, sizeof(P), 0, cudaMemcpyHostToDevice, c.st));
    CK(cudaMemcpyToSymbolAsync(c_pkO, &O, sizeof(O), 0, cudaMemcpyHostToDevice, c.st));
    CK(cudaMemcpyToSymbolAsync(c_pkL, &Lc, sizeof(Lc), 0, cudaMemcpyHostToDevice, c.st));
  }
  CK(cudaMemsetAsync(c.prof, 0, 8 * sizeof(unsigned long long), c.st));
  bp.prof = c.prof;
  for (int i = 0; i < 4; ++i) c.phase_ms[i] = 0.0;
  int rc;
  const int mode = env_int("OBCA_MODE", 0);
  // large batches are solved in chunks so that the per-problem workspace (~0.25 MB) stays within a few GB
  const int chunk = env_int("OBCA_CHUNK", 16384);
  const int NSc = P.N + 1, Vc = P.V, nObc = P.nOb;
  rc = 0;
  for (int b0 = 0; b0 < bp.B && rc == 0; b0 += chunk) {
    BatchPtrs q = bp;
    q.B = bp.B - b0 < chunk ? bp.B - b0 : chunk;
    const size_t o = (size_t)b0;
    q.x0 += 4 * o; q.xF += 4 * o; q.rx += NSc * o; q.ry += NSc * o; q.ryaw += NSc * o; q.xWS += 4 * NSc * o; q.uWS += 2 * (size_t)P.N * o;
    q.lWS += (size_t)Vc * NSc * o; q.nWS += (size_t)4 * nObc * NSc * o;
    q.xp += 4 * NSc * o; q.up += 2 * (size_t)P.N * o; q.ts += NSc * o; q.lp += (size_t)Vc * NSc * o; q.np += (size_t)4 * nObc * NSc * o;
    if (q.sl) q.sl += (size_t)nObc * NSc * o;
    if (q.duals) q.duals += ((size_t)4 * P.N + (size_t)4 * nObc * NSc) * o;
    q.exitflag += o; q.iters += o; q.kkt_err += o;
#ifdef OBCA_FAST_BUILD   // development builds: only the config-2 instantiation
    if (!(P.signed_dist && vm == 2)) { set_err("fast build: only <2,true>"); return OBCA_ERR_UNSUPPORTED; }
    rc = launch_phased<2, true>(c, P, O, q, mode);
#else
    // OBCA_MODE: 0 auto (phase-split rounds for large batches + tail kernel for the stragglers, tail kernel alone for small
    // batches), 1 tail kernel only, 2 rounds only
    if (P.signed_dist) rc = vm == 2 ? launch_phased<2, true>(c, P, O, q, mode) : launch_phased<4, true>(c, P, O, q, mode);
    else rc = vm == 2 ? launch_phased<2, false>(c, P, O, q, mode) : launch_phased<4, false>(c, P, O, q, mode);
#endif
  }
  if (rc) return rc;
  CK(cudaEventRecord(c.ev1, c.st));
  CK(cudaMemcpyAsync(c.prof_host, c.prof, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c.st));
  CK(cudaStreamSynchronize(c.st));
  // the reference's `time` is the wall time of solve(m) alone (ParkingSignedDist.jl:239-241,297): DualMultWS (:219) and
  // the model build are outside it.  ev0 .. evm = DualMultWS (when the library runs it), evm .. ev1 = the solve.
  float ms = 0.f, ms_ws = 0.f;
  CK(cudaEventElapsedTime(&ms, c.evm, c.ev1));
  CK(cudaEventElapsedTime(&ms_ws, c.ev0, c.evm));
  c.last_solve_s = ms * 1e-3; c.last_dualws_s = ms_ws * 1e-3; c.phase_ms[4] = ms_ws;
  if (seconds) *seconds = ms * 1e-3;
  return 0;
}

template <int VM, bool SDV>
static int launch_eval(DevCtx& c, const ParkProblem& P, int B, const EvalIn& in, const EvalOut& out, int reps) {
  const size_t total = (size_t)B * (P.N + 1);
  const int threads = 128;
  int blocks = (int)((total + threads - 1) / threads);
  for (int r = 0; r < reps; ++r) k_parking_eval<VM, SDV><<<blocks, threads, 0, c.st>>>(P, B, in, out);
  CK(cudaGetLastError());
  return 0;
}

template <bool SDV>
static int launch_quad(DevCtx& c, const QuadProblem& P, const IpmOpts& O, const QBatchPtrs& bp) {
  QLay L = make_qlayout(P);
  if (L.NSP > 128) { set_err("horizon too long for this build (N+1 <= 128)"); return OBCA_ERR_UNSUPPORTED; }
  const size_t smem = (size_t)QuadSolver<SDV>::smem_doubles(P.N) * sizeof(double);
  if (smem > 200 * 1024) { set_err("horizon too long for the quadcopter kernel's shared memory"); return OBCA_ERR_UNSUPPORTED; }
  CK(cudaFuncSetAttribute(k_quad_solve<SDV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_quad_solve<SDV>, L.NSP, smem));
  if (occ < 1) occ = 1;
  int grid = c.sms * occ;
  if (grid > bp.B) grid = bp.B;
  const size_t need = (size_t)grid * L.total * L.NSP * sizeof(double);
  int rc = ensure((void**)&c.W, &c.Wbytes, need);
  if (rc) return rc;
  CK(cudaMemcpyToSymbolAsync(c_pkO, &O, sizeof(O), 0, cudaMemcpyHostToDevice, c.st));   // IpmDriver reads the options here
  CK(cudaMemcpyToSymbolAsync(c_qP, &P, sizeof(P), 0, cudaMemcpyHostToDevice, c.st));
  CK(cudaMemcpyToSymbolAsync(c_qL, &L, sizeof(L), 0, cudaMemcpyHostToDevice, c.st));
  CK(cudaMemsetAsync(c.counter, 0, sizeof(int), c.st));
  k_quad_solve<SDV><<<grid, L.NSP, smem, c.st>>>(P, O, L, bp, c.W, c.counter);
  CK(cudaGetLastError());
  return 0;
}

extern "C" {

#ifdef OBCA_QPROF   // development builds only: cycles of the quadcopter sweep's phases, summed over CTAs; resets the counters
int obca_debug_qprof(unsigned long long* out8) {
  unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (cudaMemcpyFromSymbol(out8, g_qprof, sizeof(z)) != cudaSuccess) return -1;
  return cudaMemcpyToSymbol(g_qprof, z, sizeof(z)) == cudaSuccess ? 0 : -1;
}
#endif

int obca_version(void) { return OBCA_VERSION; }
int obca_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}
const char* obca_last_error(void) { return g_err.c_str(); }
void obca_default_opts(obca_opts* o) {
  IpmOpts d = default_opts();
  o->tol = d.tol; o->max_iter = d.max_iter; o->mu_init = d.mu_init; o->mu_min = d.mu_min;
  o->kappa_eps = d.kappa_eps; o->kappa_mu = d.kappa_mu; o->theta_mu = d.theta_mu; o->tau_min = d.tau_min;
  o->kappa1 = d.kappa1; o->kappa2 = d.kappa2; o->kappa_sigma = d.kappa_sigma; o->s_max = d.s_max;
  o->dual_inf_tol = d.dual_inf_tol; o->constr_viol_tol = d.constr_viol_tol; o->compl_inf_tol = d.compl_inf_tol;
  o->dw_min = d.dw_min; o->dw_first = d.dw_first; o->dw_max = d.dw_max; o->kw_minus = d.kw_minus;
  o->kw_plus = d.kw_plus; o->kw_plus_first = d.kw_plus_first;
  o->gamma_theta = d.gamma_theta; o->gamma_phi = d.gamma_phi; o->delta = d.delta; o->s_theta = d.s_theta;
  o->s_phi = d.s_phi; o->eta_phi = d.eta_phi; o->gamma_alpha = d.gamma_alpha; o->max_backtrack = d.max_backtrack;
  o->dc = d.dc; o->max_kick = d.max_kick; o->quad_dual_ws = d.quad_dual_ws;
  o->device = 0; o->retry = 1; o->q4 = 0;
}

int obca_parking_solve_batch_dev(int B, int N, int nOb, const int* vOb, const double* A, const double* b,
                                 const double* x0, const double* xF, double Ts, double L, const double* ego,
                                 const double* XYbounds, const double* rx, const double* ry, const double* ryaw,
                                 const double* xWS, const double* uWS, const double* lWS, const double* nWS,
                                 int fixTime, int signed_dist, const obca_opts* opts, double* xp, double* up,
                                 double* ts, double* lp, double* np, double* sl, int* exitflag, int* iters,
                                 double* kkt_err, double* solve_seconds) {
  if (B <= 0 || !vOb || !A || !b || !x0 || !xF || !ego || !XYbounds || !rx || !ry || !ryaw || !xWS || !uWS || !xp ||
      !up || !ts || !lp || !np || !exitflag || !iters || !kkt_err) { set_err("null argument"); return OBCA_ERR_ARG; }
  ParkProblem P;
  if (fill_problem(P, N, nOb, vOb, A, b, Ts, L, ego, XYbounds, fixTime, signed_dist)) {
    set_err("unsupported problem shape"); return OBCA_ERR_ARG;
  }
  DevCtx* c;
  int rc = get_ctx(opts ? opts->device : 0, &c);
  if (rc) return rc;
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  IpmOpts O = to_ipm(opts);
  BatchPtrs bp;
  bp.x0 = x0; bp.xF = xF; bp.rx = rx; bp.ry = ry; bp.ryaw = ryaw; bp.xWS = xWS; bp.uWS = uWS;
  bp.xp = xp; bp.up = up; bp.ts = ts; bp.lp = lp; bp.np = np; bp.sl = sl; bp.duals = nullptr;
  bp.exitflag = exitflag; bp.iters = iters; bp.kkt_err = kkt_err; bp.B = B; bp.retry = opts ? opts->retry : 1;
  bp.q4 = opts ? opts->q4 : 0;
  CK(cudaEventRecord(c->ev0, c->st));
  if (lWS && nWS) { bp.lWS = lWS; bp.nWS = nWS; }
  else {
    // the reference runs DualMultWS inside the NLP driver (ParkingSignedDist.jl:219); its result goes to a scratch of
    // the library (never to the caller's output arrays, which stay untouched until a problem has finished)
    const size_t NS = (size_t)N + 1, nl = (size_t)P.V * NS * B, nn = 4 * (size_t)nOb * NS * B;
    rc = ensure((void**)&c->wsd, &c->wsd_bytes, (nl + nn) * sizeof(double));
    if (rc) return rc;
    rc = run_dualws(*c, P, B, rx, ry, ryaw, c->wsd, c->wsd + nl, nullptr);
    if (rc) return rc;
    bp.lWS = c->wsd; bp.nWS = c->wsd + nl;
  }
  CK(cudaEventRecord(c->evm, c->st));      // `time` = the solve alone, as in the reference (:239-241); see obca_last_times
  return solve_dev_impl(*c, P, O, bp, solve_seconds);
}

int obca_parking_solve_batch(int B, int N, int nOb, const int* vOb, const double* A, const double* b,
                             const double* x0, const double* xF, double Ts, double L, const double* ego,
                             const double* XYbounds, const double* rx, const double* ry, const double* ryaw,
                             const double* xWS, const double* uWS, const double* lWS, const double* nWS,
                             int fixTime, int signed_dist, const obca_opts* opts, double* xp, double* up, double* ts,
                             double* lp, double* np, double* sl, int* exitflag, int* iters, double* kkt_err,
                             double* solve_seconds) {
  if (B <= 0 || !vOb || !x0 || !xF || !rx || !ry || !ryaw || !xWS || !uWS || !xp || !up || !ts || !lp || !np ||
      !exitflag || !iters || !kkt_err) { set_err("null argument"); return OBCA_ERR_ARG; }
  if (N < 2 || nOb < 1 || nOb > OBCA_MAX_OB) { set_err("unsupported problem shape"); return OBCA_ERR_ARG; }
  int V = 0;
  for (int j = 0; j < nOb; ++j) V += vOb[j];
  DevCtx* c;
  int rc = get_ctx(opts ? opts->device : 0, &c);
  if (rc) return rc;
  const size_t NS = N + 1;
  const size_t n_x0 = 4, n_r = NS, n_xw = 4 * NS, n_uw = 2 * (size_t)N, n_l = V * NS, n_n = 4 * (size_t)nOb * NS, n_s = nOb * NS;
  const bool have_ws = lWS && nWS;
  // staging layout (doubles): inputs then outputs
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off += n * (size_t)B; return o; };
  const size_t o_x0 = take(n_x0), o_xF = take(n_x0), o_rx = take(n_r), o_ry = take(n_r), o_ryaw = take(n_r);
  const size_t o_xw = take(n_xw), o_uw = take(n_uw), o_lw = take(n_l), o_nw = take(n_n);
  const size_t o_xp = take(n_xw), o_up = take(n_uw), o_ts = take(n_r), o_lp = take(n_l), o_np = take(n_n), o_sl = take(n_s);
  const size_t o_err = take(1);
  const size_t dbytes = off * sizeof(double);
  const size_t ibytes = 2 * (size_t)B * sizeof(int);
  // one staging buffer per device: the whole call (copies in, solve, copies out) runs under the device lock
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  rc = ensure((void**)&c->stage, &c->stage_bytes, dbytes + ibytes);
  if (rc) return rc;
  double* d = (double*)c->stage;
  int* di = (int*)(c->stage + dbytes);
  cudaStream_t st = c->st;
#define H2D(dst, src, n) CK(cudaMemcpyAsync(d + (dst), (src), (n) * (size_t)B * sizeof(double), cudaMemcpyHostToDevice, st))
  H2D(o_x0, x0, n_x0); H2D(o_xF, xF, n_x0); H2D(o_rx, rx, n_r); H2D(o_ry, ry, n_r); H2D(o_ryaw, ryaw, n_r);
  H2D(o_xw, xWS, n_xw); H2D(o_uw, uWS, n_uw);
  if (have_ws) { H2D(o_lw, lWS, n_l); H2D(o_nw, nWS, n_n); }
#undef H2D
  obca_opts o2;
  if (opts) o2 = *opts; else obca_default_opts(&o2);
  rc = obca_parking_solve_batch_dev(B, N, nOb, vOb, A, b, d + o_x0, d + o_xF, Ts, L, ego, XYbounds, d + o_rx, d + o_ry,
                                    d + o_ryaw, d + o_xw, d + o_uw, have_ws ? d + o_lw : nullptr,
                                    have_ws ? d + o_nw : nullptr, fixTime, signed_dist, &o2, d + o_xp, d + o_up,
                                    d + o_ts, d + o_lp, d + o_np, d + o_sl, di, di + B, d + o_err, solve_seconds);
  if (rc) return rc;
#define D2H(dst, src, n) CK(cudaMemcpyAsync((dst), d + (src), (n) * (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, st))
  D2H(xp, o_xp, n_xw); D2H(up, o_up, n_uw); D2H(ts, o_ts, n_r); D2H(lp, o_lp, n_l); D2H(np, o_np, n_n);
  if (sl) D2H(sl, o_sl, n_s);
  D2H(kkt_err, o_err, 1);
#undef D2H
  CK(cudaMemcpyAsync(exitflag, di, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(iters, di + B, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return 0;
}

int obca_dualmultws_batch(int B, int N, int nOb, const int* vOb, const double* A, const double* b, const double* ego,
                          const double* rx, const double* ry, const double* ryaw, const obca_opts* opts, double* lp,
                          double* np, double* dd) {
  if (B <= 0 || !vOb || !A || !b || !ego || !rx || !ry || !ryaw || !lp || !np) { set_err("null argument"); return OBCA_ERR_ARG; }
  ParkProblem P;
  double xy[4] = {0, 1, 0, 1};
  if (fill_problem(P, N, nOb, vOb, A, b, 1.0, 1.0, ego, xy, 0, 1)) { set_err("unsupported problem shape"); return OBCA_ERR_ARG; }
  DevCtx* c;
  int rc = get_ctx(opts ? opts->device : 0, &c);
  if (rc) return rc;
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  const size_t NS = N + 1, nr = NS * B, nl = (size_t)P.V * NS * B, nn = 4 * (size_t)nOb * NS * B, ndd = (size_t)nOb * NS * B;
  rc = ensure((void**)&c->stage, &c->stage_bytes, (3 * nr + nl + nn + ndd) * sizeof(double));
  if (rc) return rc;
  double* d = (double*)c->stage;
  CK(cudaMemcpyAsync(d, rx, nr * sizeof(double), cudaMemcpyHostToDevice, c->st));
  CK(cudaMemcpyAsync(d + nr, ry, nr * sizeof(double), cudaMemcpyHostToDevice, c->st));
  CK(cudaMemcpyAsync(d + 2 * nr, ryaw, nr * sizeof(double), cudaMemcpyHostToDevice, c->st));
  double* dl = d + 3 * nr; double* dn = dl + nl; double* ddv = dn + nn;
  rc = run_dualws(*c, P, B, d, d + nr, d + 2 * nr, dl, dn, ddv);
  if (rc) return rc;
  CK(cudaMemcpyAsync(lp, dl, nl * sizeof(double), cudaMemcpyDeviceToHost, c->st));
  CK(cudaMemcpyAsync(np, dn, nn * sizeof(double), cudaMemcpyDeviceToHost, c->st));
  if (dd) CK(cudaMemcpyAsync(dd, ddv, ndd * sizeof(double), cudaMemcpyDeviceToHost, c->st));
  CK(cudaStreamSynchronize(c->st));
  return 0;
}

int obca_check_parking(int B, int N, int nOb, const int* vOb, const double* A, const double* b, const double* x0,
                       const double* xF, double Ts, double L, const double* ego, const double* XYbounds,
                       const double* x, const double* u, const double* l, const double* n, const double* timeScale,
                       const double* sl, int fixTime, int sd, const obca_opts* opts, int* feasible, int* e,
                       int* strict) {
  if (B <= 0 || !vOb || !A || !b || !x0 || !xF || !ego || !XYbounds || !x || !u || !l || !n || !timeScale || !feasible) {
    set_err("null argument"); return OBCA_ERR_ARG;
  }
  ParkProblem P;
  if (fill_problem(P, N, nOb, vOb, A, b, Ts, L, ego, XYbounds, fixTime, sd)) { set_err("unsupported problem shape"); return OBCA_ERR_ARG; }
  DevCtx* c;
  int rc = get_ctx(opts ? opts->device : 0, &c);
  if (rc) return rc;
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  const size_t NS = N + 1;
  const size_t nx = 4 * NS * B, nu = 2 * (size_t)N * B, nl = (size_t)P.V * NS * B, nn = 4 * (size_t)nOb * NS * B, nt = NS * B,
               ns = (size_t)nOb * NS * B, n0 = 4 * (size_t)B;
  const size_t dbytes = (2 * n0 + nx + nu + nl + nn + nt + ns) * sizeof(double);
  rc = ensure((void**)&c->stage, &c->stage_bytes, dbytes + 9 * (size_t)B * sizeof(int));
  if (rc) return rc;
  double* d = (double*)c->stage;
  double *d0 = d, *dF = d0 + n0, *dx = dF + n0, *du = dx + nx, *dl = du + nu, *dn = dl + nl, *dt = dn + nn, *ds = dt + nt;
  int* di = (int*)(c->stage + dbytes);
  cudaStream_t st = c->st;
  CK(cudaMemcpyAsync(d0, x0, n0 * 8, cudaMemcpyHostToDevice, st)); CK(cudaMemcpyAsync(dF, xF, n0 * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dx, x, nx * 8, cudaMemcpyHostToDevice, st)); CK(cudaMemcpyAsync(du, u, nu * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dl, l, nl * 8, cudaMemcpyHostToDevice, st)); CK(cudaMemcpyAsync(dn, n, nn * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dt, timeScale, nt * 8, cudaMemcpyHostToDevice, st));
  if (sl) CK(cudaMemcpyAsync(ds, sl, ns * 8, cudaMemcpyHostToDevice, st));
  const int grid = B < 4 * c->sms ? B : 4 * c->sms;
  k_check<<<grid, 128, 0, st>>>(P, B, d0, dF, dx, du, dl, dn, dt, sl ? ds : nullptr, sd, di, di + B, di + 8 * (size_t)B);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(feasible, di, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (e) CK(cudaMemcpyAsync(e, di + B, 7 * (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (strict) CK(cudaMemcpyAsync(strict, di + 8 * (size_t)B, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return 0;
}

int obca_quadcopter_solve_batch(int B, int N, const double* x0, const double* xF, double Ts, double R, const double* ob,
                                const double* xWS, double timeWS, int signed_dist, const obca_opts* opts, double* xp,
                                double* up, double* ts, double* lp, double* slack, int* exitflag, int* iters,
                                double* kkt_err, double* solve_seconds) {
  if (B <= 0 || !x0 || !xF || !ob || !xWS || !xp || !up || !ts || !lp || !exitflag || !iters || !kkt_err) {
    set_err("null argument"); return OBCA_ERR_ARG;
  }
  QuadProblem P;
  if (fill_quad_problem(P, N, Ts, R, ob, signed_dist)) { set_err("unsupported problem shape"); return OBCA_ERR_ARG; }
  DevCtx* c;
  int rc = get_ctx(opts ? opts->device : 0, &c);
  if (rc) return rc;
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  IpmOpts O = to_ipm(opts);
  if (!opts) O.max_iter = 3000;                        // the reference sets no max_iter here (Ipopt default)
  const size_t NS = N + 1;
  const size_t n0 = 12 * (size_t)B, nx = 12 * NS * B, nu = 4 * (size_t)N * B, nt = NS * B, nl = 30 * NS * B, ns = 5 * NS * B;
  const size_t dbytes = (2 * n0 + 2 * nx + nu + nt + nl + ns + B) * sizeof(double);
  rc = ensure((void**)&c->stage, &c->stage_bytes, dbytes + 2 * (size_t)B * sizeof(int));
  if (rc) return rc;
  double* d = (double*)c->stage;
  double *d0 = d, *dF = d0 + n0, *dW = dF + n0, *dxp = dW + nx, *dup = dxp + nx, *dts = dup + nu, *dlp = dts + nt, *dsl = dlp + nl,
         *derr = dsl + ns;
  int* di = (int*)(c->stage + dbytes);
  cudaStream_t st = c->st;
  CK(cudaMemcpyAsync(d0, x0, n0 * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dF, xF, n0 * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dW, xWS, nx * 8, cudaMemcpyHostToDevice, st));
  QBatchPtrs bp;
  bp.x0 = d0; bp.xF = dF; bp.xWS = dW; bp.timeWS = timeWS; bp.xp = dxp; bp.up = dup; bp.ts = dts; bp.lp = dlp; bp.slack = dsl;
  bp.exitflag = di; bp.iters = di + B; bp.kkt_err = derr; bp.B = B; bp.prof = c->prof;
  CK(cudaMemsetAsync(c->prof, 0, 8 * sizeof(unsigned long long), st));
  CK(cudaEventRecord(c->ev0, st));
  rc = signed_dist ? launch_quad<true>(*c, P, O, bp) : launch_quad<false>(*c, P, O, bp);
  if (rc) return rc;
  CK(cudaEventRecord(c->ev1, st));
  CK(cudaMemcpyAsync(xp, dxp, nx * 8, cudaMemcpyDeviceToHost, st)); CK(cudaMemcpyAsync(up, dup, nu * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(ts, dts, nt * 8, cudaMemcpyDeviceToHost, st)); CK(cudaMemcpyAsync(lp, dlp, nl * 8, cudaMemcpyDeviceToHost, st));
  if (slack && signed_dist) CK(cudaMemcpyAsync(slack, dsl, ns * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(kkt_err, derr, (size_t)B * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(exitflag, di, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(iters, di + B, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(c->prof_host, c->prof, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  if (solve_seconds) *solve_seconds = ms * 1e-3;
  return 0;
}

int obca_check_quadcopter(int B, int N, const double* x, const double* u, const double* timeScale, const double* x0,
                          const double* xF, double Ts, const double* lambda, const double* ob, double R,
                          const obca_opts* opts, int* feasible, double* worst) {
  if (B <= 0 || !x || !u || !timeScale || !x0 || !xF || !lambda || !ob || !feasible) { set_err("null argument"); return OBCA_ERR_ARG; }
  QuadProblem P;
  if (fill_quad_problem(P, N, Ts, R, ob, 0)) { set_err("unsupported problem shape"); return OBCA_ERR_ARG; }
  DevCtx* c;
  int rc = get_ctx(opts ? opts->device : 0, &c);
  if (rc) return rc;
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  const size_t NS = N + 1;
  const size_t n0 = 12 * (size_t)B, nx = 12 * NS * B, nu = 4 * (size_t)N * B, nt = NS * B, nl = 30 * NS * B;
  const size_t dbytes = (2 * n0 + nx + nu + nt + nl + B) * sizeof(double);
  rc = ensure((void**)&c->stage, &c->stage_bytes, dbytes + (size_t)B * sizeof(int));
  if (rc) return rc;
  double* d = (double*)c->stage;
  double *d0 = d, *dF = d0 + n0, *dx = dF + n0, *du = dx + nx, *dt = du + nu, *dl = dt + nt, *dw = dl + nl;
  int* di = (int*)(c->stage + dbytes);
  cudaStream_t st = c->st;
  CK(cudaMemcpyAsync(d0, x0, n0 * 8, cudaMemcpyHostToDevice, st)); CK(cudaMemcpyAsync(dF, xF, n0 * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dx, x, nx * 8, cudaMemcpyHostToDevice, st)); CK(cudaMemcpyAsync(du, u, nu * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dt, timeScale, nt * 8, cudaMemcpyHostToDevice, st)); CK(cudaMemcpyAsync(dl, lambda, nl * 8, cudaMemcpyHostToDevice, st));
  const int grid = B < 4 * c->sms ? B : 4 * c->sms;
  k_quad_check<<<grid, 128, 0, st>>>(P, B, d0, dF, dx, du, dt, dl, di, dw);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(feasible, di, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (worst) CK(cudaMemcpyAsync(worst, dw, (size_t)B * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return 0;
}

int obca_last_profile(int device, unsigned long long* out8) {
  if (device < 0 || device >= 64 || !out8 || !g_dev[device].init) { set_err("no profile"); return OBCA_ERR_ARG; }
  for (int i = 0; i < 8; ++i) out8[i] = g_dev[device].prof_host[i];
  return 0;
}

int obca_last_schedule(int device, int* rounds, int* handed_over, double* kernel_ms5) {
  if (device < 0 || device >= 64 || !g_dev[device].init) { set_err("no schedule"); return OBCA_ERR_ARG; }
  if (rounds) *rounds = g_dev[device].last_rounds;
  if (handed_over) *handed_over = g_dev[device].last_tail;
  if (kernel_ms5) for (int i = 0; i < 5; ++i) kernel_ms5[i] = g_dev[device].phase_ms[i];
  return 0;
}

int obca_last_times(int device, double* dualws_seconds, double* solve_seconds) {
  if (device < 0 || device >= 64 || !g_dev[device].init) { set_err("no solve yet"); return OBCA_ERR_ARG; }
  if (dualws_seconds) *dualws_seconds = g_dev[device].last_dualws_s;
  if (solve_seconds) *solve_seconds = g_dev[device].last_solve_s;
  return 0;
}

int obca_parking_eval_sizes(int N, int nOb, const int* vOb, int signed_dist, long long* n_out, long long* m_out) {
  ParkProblem P;
  double A[2 * OBCA_MAX_ROWS] = {0}, b[OBCA_MAX_ROWS] = {0}, ego[4] = {1, 1, 1, 1}, xy[4] = {0, 1, 0, 1};
  if (fill_problem(P, N, nOb, vOb, A, b, 1.0, 1.0, ego, xy, 0, signed_dist)) { set_err("unsupported problem shape"); return OBCA_ERR_ARG; }
  if (n_out) *n_out = (long long)eval_n(P);
  if (m_out) *m_out = (long long)eval_m(P);
  return 0;
}

int obca_parking_eval_batch_dev(int B, int N, int nOb, const int* vOb, const double* A, const double* b,
                                const double* x0, const double* xF, double Ts, double L, const double* ego,
                                const double* XYbounds, const double* rx, const double* ry, const double* ryaw,
                                const double* xp, const double* up, const double* ts, const double* lp,
                                const double* np, const double* sl, const double* y, int fixTime, int signed_dist,
                                const obca_opts* opts, double* c_out, double* gradL_out, double* fk_out, int reps,
                                double* kernel_ms) {
  if (B <= 0 || !vOb || !A || !b || !x0 || !xF || !ego || !XYbounds || !rx || !ry || !ryaw || !xp || !up || !lp || !np ||
      !c_out || !gradL_out || !fk_out || (signed_dist && !sl) || (!fixTime && !ts)) { set_err("null argument"); return OBCA_ERR_ARG; }
  ParkProblem P;
  if (fill_problem(P, N, nOb, vOb, A, b, Ts, L, ego, XYbounds, fixTime, signed_dist)) { set_err("unsupported problem shape"); return OBCA_ERR_ARG; }
  DevCtx* c;
  int rc = get_ctx(opts ? opts->device : 0, &c);
  if (rc) return rc;
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  EvalIn in; in.x0 = x0; in.xF = xF; in.rx = rx; in.ry = ry; in.ryaw = ryaw; in.xp = xp; in.up = up; in.ts = ts; in.lp = lp;
  in.np = np; in.sl = sl; in.y = y;
  EvalOut out; out.c = c_out; out.gradL = gradL_out; out.fk = fk_out;
  if (reps < 1) reps = 1;
  const int vm = max_vob(P) <= 2 ? 2 : 4;
  CK(cudaEventRecord(c->ev0, c->st));
  if (P.signed_dist) rc = vm == 2 ? launch_eval<2, true>(*c, P, B, in, out, reps) : launch_eval<4, true>(*c, P, B, in, out, reps);
  else rc = vm == 2 ? launch_eval<2, false>(*c, P, B, in, out, reps) : launch_eval<4, false>(*c, P, B, in, out, reps);
  if (rc) return rc;
  CK(cudaEventRecord(c->ev1, c->st));
  CK(cudaStreamSynchronize(c->st));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  if (kernel_ms) *kernel_ms = ms / reps;
  return 0;
}
}

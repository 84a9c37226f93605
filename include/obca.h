/* obca.h -- C-ABI of libobca.so, the B200-native replacement for the JuMP + Ipopt solve inside the OBCA
 * reference's NLP drivers.  Plain C, plain pointers and sizes; no allocation escapes the library.
 *
 * The reference (XiaojingGeorgeZhang/OBCA) has no FFI of its own: its boundary is the set of Julia functions that
 * AutonomousParking/main.jl calls positionally.  Each entry point below is what the Julia shim of the same name
 * (julia/*.jl) `ccall`s in place of the reference's `Model(solver=IpoptSolver(...))` ...
 * `solve(m)` ... `getvalue(...)` sequence:
 *
 *   obca_parking_solve_batch  <- ParkingSignedDist.jl:29-314  (signed_dist=1)  and  ParkingDist.jl:29-315 (=0)
 *                                 called from main.jl:269 / main.jl:258
 *   obca_dualmultws_batch     <- DualMultWS.jl:29-86           called from ParkingSignedDist.jl:219, ParkingDist.jl:221
 *   obca_check_parking        <- ParkingConstraints.jl:29-149  called from ParkingSignedDist.jl:253,278, ParkingDist.jl:259,277
 *
 * Conventions
 *   - every array is float64, COLUMN-MAJOR exactly as the Julia caller holds it; a batch of B problems is the B
 *     per-problem arrays stored one after the other (problem index slowest).  B = 1 reproduces the reference call.
 *   - vOb[nOb] are HALF-SPACE counts (main.jl:101 passes vObMPC = vOb-1); A is sum(vOb) x 2, b is sum(vOb)
 *     (obstHrep.jl:39-40), shared by the whole batch.
 *   - return value: 0 = executed (per-problem success is in exitflag[]), < 0 = usage / CUDA error (see
 *     obca_last_error()).  The library never throws and never falls back to a CPU path: without a CUDA device
 *     every compute entry point returns OBCA_ERR_NO_DEVICE.
 *   - `*_dev` variants take DEVICE pointers (cudaMalloc'ed by the caller on opts->device) and enqueue on the
 *     library's stream of that device; they return after the stream has been synchronised.
 */
#ifndef OBCA_H
#define OBCA_H

#ifdef __cplusplus
extern "C" {
#endif

#define OBCA_VERSION 200
#define OBCA_ERR_ARG (-1)
#define OBCA_ERR_NO_DEVICE (-2)
#define OBCA_ERR_CUDA (-3)
#define OBCA_ERR_UNSUPPORTED (-4)

/* Interior-point options: the Ipopt options set at ParkingSignedDist.jl:41-43 (tol, max_iter,
 * min_hessian_perturbation -> dw_min, jacobian_regularization_value is replaced by the always-on dc) and the Ipopt
 * defaults the reference leaves untouched.  obca_default_opts() fills in exactly those values. */
typedef struct obca_opts {
  double tol;
  int max_iter;
  double mu_init, mu_min;
  double kappa_eps, kappa_mu, theta_mu, tau_min;
  double kappa1, kappa2, kappa_sigma, s_max;
  double dual_inf_tol, constr_viol_tol, compl_inf_tol;
  double dw_min, dw_first, dw_max, kw_minus, kw_plus, kw_plus_first;
  double gamma_theta, gamma_phi, delta, s_theta, s_phi, eta_phi, gamma_alpha;
  int max_backtrack;
  double dc;
  int max_kick;      /* restoration substitute: barrier kicks per attempt after a failed line search (default 3) */
  int quad_dual_ws;  /* quadcopter: 1 (default) closed-form dual warm start; 0 the reference's l = 0.05 start */
  /* execution */
  int device;        /* CUDA device ordinal used by this call */
  int retry;         /* 1: the reference's status / retry logic: ParkingSignedDist.jl:256-290 (a failed first attempt is always
                        followed by one more solve from the last iterate, then ParkingConstraints decides) and
                        ParkingDist.jl:245-289 (after a failed first attempt ParkingConstraints decides whether the point is
                        accepted or solved again); 0: single attempt, exitflag = converged */
  int q4;            /* ParkingDist only: 1 reproduces the reference's inverted test after a SECOND failure
                        (ParkingDist.jl:278-282: Feasible == 0 -> exitflag 1); 0 (default): exitflag = Feasible */
} obca_opts;

int obca_version(void);
int obca_device_count(void);                 /* number of visible CUDA devices (0 if none) */
const char* obca_last_error(void);
void obca_default_opts(obca_opts* o);

/* Batched ParkingSignedDist / ParkingDist.
 * inputs (per problem):  x0[4], xF[4], rx/ry/ryaw[N+1], xWS (N+1)x4, uWS Nx2 (first N rows of the reference's uWS,
 *                        ParkingSignedDist.jl:217), lWS (N+1)xV and nWS (N+1)x4nOb as returned by DualMultWS
 *                        (pass lWS = nWS = NULL to have the library run DualMultWS itself, as the reference does
 *                        at ParkingSignedDist.jl:219)
 * outputs (per problem): xp 4x(N+1), up 2xN, ts (N+1) [ones if fixTime], lp Vx(N+1), np 4nOb x(N+1)
 *                        (ParkingSignedDist.jl:302-313), sl nOb x(N+1) (may be NULL),
 *                        exitflag (1 = converged, 0 = not), iters, kkt_err (Ipopt's scaled NLP error E_0),
 *                        solve_seconds[1] = device time of the solve alone -- the reference's `time` (:239-241, :297) is
 *                        the wall time of solve(m), which excludes DualMultWS (:219); obca_last_times() gives both.  */
int obca_parking_solve_batch(int B, int N, int nOb, const int* vOb, const double* A, const double* b,
                             const double* x0, const double* xF, double Ts, double L, const double* ego,
                             const double* XYbounds, const double* rx, const double* ry, const double* ryaw,
                             const double* xWS, const double* uWS, const double* lWS, const double* nWS,
                             int fixTime, int signed_dist, const obca_opts* opts, double* xp, double* up, double* ts,
                             double* lp, double* np, double* sl, int* exitflag, int* iters, double* kkt_err,
                             double* solve_seconds);

/* Same, all batch arrays (x0 ... nWS, xp ... kkt_err) are device pointers on opts->device; vOb, A, b, ego,
 * XYbounds, opts and solve_seconds stay host pointers. */
int obca_parking_solve_batch_dev(int B, int N, int nOb, const int* vOb, const double* A, const double* b,
                                 const double* x0, const double* xF, double Ts, double L, const double* ego,
                                 const double* XYbounds, const double* rx, const double* ry, const double* ryaw,
                                 const double* xWS, const double* uWS, const double* lWS, const double* nWS,
                                 int fixTime, int signed_dist, const obca_opts* opts, double* xp, double* up,
                                 double* ts, double* lp, double* np, double* sl, int* exitflag, int* iters,
                                 double* kkt_err, double* solve_seconds);

/* Batched DualMultWS (DualMultWS.jl:29; `ego` is a global there, :39).  lp (N+1)xV, np (N+1)x4nOb per problem
 * (already transposed like :81-84); d (N+1)xnOb optional (optimal objective = ego/obstacle distance). */
int obca_dualmultws_batch(int B, int N, int nOb, const int* vOb, const double* A, const double* b, const double* ego,
                          const double* rx, const double* ry, const double* ryaw, const obca_opts* opts, double* lp,
                          double* np, double* d);

/* Batched ParkingConstraints (ParkingConstraints.jl:29-149), restated verbatim including its quirks
 * (SURVEY.md A.4-Q3).  x 4x(N+1), u 2xN, l Vx(N+1), n 4nOb x(N+1), timeScale (N+1).
 * feasible[B] = the reference's return value (0/1); e[7*B] = its seven pass flags; strict[B] (optional) = a
 * checker without the quirks: every dynamics row, every obstacle, box bounds, slack-aware distance row. */
int obca_check_parking(int B, int N, int nOb, const int* vOb, const double* A, const double* b, const double* x0,
                       const double* xF, double Ts, double L, const double* ego, const double* XYbounds,
                       const double* x, const double* u, const double* l, const double* n, const double* timeScale,
                       const double* sl, int fixTime, int sd, const obca_opts* opts, int* feasible, int* e,
                       int* strict);

/* Batched QuadcopterSignedDist (signed_dist=1, QuadcopterNavigation/QuadcopterSignedDist.jl:25-300, called from
 * mainQuadcopter.jl:152) / QuadcopterDist (signed_dist=0, QuadcopterDist.jl:25-282, mainQuadcopter.jl:145).
 * inputs:  x0, xF 12 per problem; ob = ob1..ob5 as a 6 x 5 column-major block shared by the batch; xWS 12x(N+1) per
 *          problem; timeWS scalar.  (uWS is ignored by the reference, QuadcopterSignedDist.jl:202.)
 * outputs: xp 12x(N+1), up 4xN, ts (N+1), lp 30x(N+1) (= [l1;l2;l3;l4;l5], :296), slack 5x(N+1) (SD only, may be
 *          NULL), exitflag 1 / 0 / 2 (2 = converged but sum(slack) > 1e-3, :283-288), iters, kkt_err, solve_seconds.
 * A single solve attempt, like the reference (flag = 1, :227-235).  opts == NULL: defaults with max_iter = 3000
 * (the reference leaves Ipopt's default). */
int obca_quadcopter_solve_batch(int B, int N, const double* x0, const double* xF, double Ts, double R, const double* ob,
                                const double* xWS, double timeWS, int signed_dist, const obca_opts* opts, double* xp,
                                double* up, double* ts, double* lp, double* slack, int* exitflag, int* iters,
                                double* kkt_err, double* solve_seconds);

/* Batched constrSatisfaction (QuadcopterNavigation/constrSatisfaction.jl:25-204, called from mainQuadcopter.jl:147,154),
 * restated verbatim (tolerance 1e-3, rows for i = 1..N, Dist-variant box on x10, one-sided norm row, quirk Q5).
 * feasible[B] = the reference's Bool; worst[B] (optional) = largest violation found. */
int obca_check_quadcopter(int B, int N, const double* x, const double* u, const double* timeScale, const double* x0,
                          const double* xF, double Ts, const double* lambda, const double* ob, double R,
                          const obca_opts* opts, int* feasible, double* worst);

/* Device times of the last obca_parking_solve_batch[_dev] on `device`: DualMultWS (0 when the caller passed lWS / nWS) and the
 * solve alone (= what solve_seconds reports; the reference's `time`, ParkingSignedDist.jl:297). */
int obca_last_times(int device, double* dualws_seconds, double* solve_seconds);

/* Counters of the last solve on `device`, summed over the batch.  Parking (phase-split schedule): out8[7] = K1 evaluations in
 * the round kernel k_pk_eval (of which out8[5] are re-evaluations after a barrier update), out8[6] = merit-function
 * evaluations in k_pk_step, out8[4] / out8[3] = K1 / merit evaluations inside the tail kernel.  Quadcopter kernel: per-phase
 * device cycle counters {eval, kkt, recover, merit, update, serial, #merit evaluations, #K1 evaluations} (thread 0 of every
 * CTA, clock64).  Diagnostic only. */
int obca_last_profile(int device, unsigned long long* out8);

/* How the last obca_parking_solve_batch[_dev] on `device` was scheduled (last chunk of the batch): number of phase-split
 * rounds ([k_pk_eval, k_pk_sweep, k_pk_step] kernel triples over all active problems) and the number of problems handed to
 * the persistent tail kernel afterwards (= the batch size when no rounds were run).  Environment overrides for experiments:
 * OBCA_MODE (0 auto, 1 tail kernel only, 2 rounds only), OBCA_TAIL_THRESH (hand-over point), OBCA_CHUNK (problems per
 * chunk).  kernel_ms5 (may be NULL): with OBCA_PHASE_TIMING=1 in the environment the solve records CUDA events around
 * every kernel and reports the summed times of {k_pk_eval (K1 + decisions), k_pk_sweep (K3), k_pk_step (K4), tail kernel,
 * DualMultWS} in milliseconds.  Diagnostic only. */
int obca_last_schedule(int device, int* rounds, int* handed_over, double* kernel_ms5);

/* K1 stand-alone: fused evaluation of the parking NLP in the REFERENCE's formulation at B given points (no solve);
 * what JuMP's eval_g / eval_grad_f / eval_jac_g' * y do for Ipopt (ParkingSignedDist.jl:240), one thread per
 * (problem, stage).  Per problem it reads z = (xp, ts, up, lp, np[, sl]) (n values), the row multipliers y (m values,
 * may be NULL = zeros) and rx, ry, ryaw, and writes c (m rows; inequality rows = their bodies),
 * gradL = grad_z [f + y'c] (n values) and fk (objective per stage, N+1 values).
 *   rows of c / y:  start 4 | end 4 | dyn 4xN | chain N | rate N | norm nOb x(N+1) | rot 2nOb x(N+1) | dist nOb x(N+1)
 *   entries of gradL: x 4x(N+1) | timeScale (N+1) | u 2xN | l Vx(N+1) | n 4nOb x(N+1) | sl nOb x(N+1) (SD only)
 * Algorithmic HBM bytes per problem per call: 8 * (2n + 2m + 3(N+1))  (fused variant).
 * All batch arrays are DEVICE pointers.  `reps` launches are timed together; kernel_ms = mean per launch. */
int obca_parking_eval_sizes(int N, int nOb, const int* vOb, int signed_dist, long long* n_out, long long* m_out);
int obca_parking_eval_batch_dev(int B, int N, int nOb, const int* vOb, const double* A, const double* b,
                                const double* x0, const double* xF, double Ts, double L, const double* ego,
                                const double* XYbounds, const double* rx, const double* ry, const double* ryaw,
                                const double* xp, const double* up, const double* ts, const double* lp,
                                const double* np, const double* sl, const double* y, int fixTime, int signed_dist,
                                const obca_opts* opts, double* c_out, double* gradL_out, double* fk_out, int reps,
                                double* kernel_ms);

#ifdef __cplusplus
}
#endif
#endif

/* obca_planner.h -- C-ABI of the host-side warm-start producer (libobca_planner.so, plain C++, no CUDA).
 *
 * Replaces, for a caller that has no Julia-0.6 environment, the part of AutonomousParking/main.jl that runs BEFORE the NLP:
 *   hybrid_a_star.calc_hybrid_astar_path   hybrid_a_star.jl:104-190   (with reeds_shepp.jl, collision_check.jl, a_star.jl:47-128)
 *   veloSmooth                              veloSmooth.jl:29-109
 *   the warm-start extraction               main.jl:215-248
 * Same algorithms as the Python restatement obca_b200/planner/ (which documents the reference line by line); the two are tested to
 * produce the same paths (tests/test_planner_native.py).  The reference's planner is host code as well and runs once per problem.
 *
 * All arrays are Float64, owned by the caller, column-major where two-dimensional (Julia layout).
 */
#ifndef OBCA_PLANNER_H
#define OBCA_PLANNER_H
#ifdef __cplusplus
extern "C" {
#endif

#define OBCA_PLAN_OK 0
#define OBCA_PLAN_NO_PATH 1        /* the search exhausted its open list / expansion budget (the reference prints an error and returns nothing) */
#define OBCA_PLAN_CAPACITY 2       /* the caller's arrays are too short; *n_out holds the required length */
#define OBCA_PLAN_BAD_ARG 3

int obca_planner_version(void);

/* Hybrid A* (hybrid_a_star.jl:104): start (sx, sy, syaw) -> goal (gx, gy, gyaw) among the obstacle points (ox[i], oy[i]), i < n_ob.
 * xyreso / yawreso: grid resolutions (main.jl passes XY_GRID_RESOLUTION = 0.3 and YAW_GRID_RESOLUTION = 5 deg; <= 0 selects them).
 * Output: the path sampled every MOTION_RESOLUTION = 0.1 m, *n_out points written to rx, ry, ryaw (capacity cap). */
int obca_hybrid_astar(double sx, double sy, double syaw, double gx, double gy, double gyaw, const double* ox, const double* oy, int n_ob,
                      double xyreso, double yawreso, int max_expansions, int cap, double* rx, double* ry, double* ryaw, int* n_out);

/* The obstacle point cloud of the reference's two demo scenarios (main.jl:111-142 "backwards" = 0, :170-205 "parallel" = 1). */
int obca_scenario_obstacle_points(int scenario, int cap, double* ox, double* oy, int* n_out);

/* main.jl:215-248 in one call: Hybrid A* from x0 = (x, y, yaw) to xF, speed profile, veloSmooth, steering, down-sampling by sampleN.
 * Ts <= 0 selects the scenario's variable-time sampling time (main.jl:43-58).  On success *N_out = N and
 *   rx, ry, ryaw   N+1 values each (the down-sampled path, main.jl:237-239)
 *   xWS            (N+1) x 4 column-major: x, y, yaw, v        (main.jl:247)
 *   uWS            N x 2 column-major: steering, acceleration  (main.jl:248)
 * cap = capacity of the arrays in stages (N+1 <= cap). */
int obca_plan_warmstart(const double* x0, const double* xF, int scenario, double Ts, double L, int sampleN, int cap, double* rx, double* ry,
                        double* ryaw, double* xWS, double* uWS, int* N_out);

/* The same for B start poses (x0: 3 x B column-major) towards one goal -- the randomised sweeps of main.jl:165-168 --, on nthreads host
 * threads (<= 0: hardware concurrency).  Problem i writes N[i], status[i] (OBCA_PLAN_OK / NO_PATH / CAPACITY) and, with the stride cap,
 * rx / ry / ryaw [i*cap ..], xWS [i*4*cap ..] as (N[i]+1) x 4 column-major, uWS [i*2*cap ..] as N[i] x 2 column-major. */
int obca_plan_warmstart_batch(int B, const double* x0, const double* xF, int scenario, double Ts, double L, int sampleN, int cap, int nthreads,
                              double* rx, double* ry, double* ryaw, double* xWS, double* uWS, int* N, int* status);

/* Reeds-Shepp shortest path length between two poses for maximum curvature maxc (reeds_shepp.jl:79-96); used by tests. */
double obca_reeds_shepp_length(double sx, double sy, double syaw, double gx, double gy, double gyaw, double maxc);

#ifdef __cplusplus
}
#endif
#endif

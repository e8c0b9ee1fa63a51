/*
 * dmpc_b200.h -- C-ABI of libdmpc_b200.so: the B200-native DMPC per-agent QP hot path.
 *
 * Drop-in boundary for carlosluis/multiagent_planning (reference paths relative to the
 * reference repository root).  The reference has no FFI for this path: the path sits behind
 * plain MATLAB function signatures (and the C++ class DMPC).  Each entry point below cites the
 * reference interface it replaces.  Plain pointers and sizes only; all arrays are caller-owned,
 * column-major fp64 exactly like MATLAB mxArray real data: a horizon buffer `l` is 3 x K x N,
 * element (d,k,n) at l[d + 3*(k + K*n)] (0-based), identical to the C++ reference's
 * std::vector<MatrixXd(3,K)> blocks (dmpc/cpp/dmpc.cpp:1630-1631).
 *
 * Return value: 0 = ok, <0 = API / CUDA error (text via dmpcb200_last_error).  Solver outcomes
 * are NOT errors (the reference returns flags and empty arrays, solveSoftDMPCbound.m:26-31,
 * 136-139); they are reported in the per-agent status word.
 *
 * There is no CPU fallback: every compute entry point needs a CUDA device and fails loudly
 * (DMPCB200_ERR_CUDA) without one.
 */
#ifndef DMPC_B200_H
#define DMPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMPCB200_ABI_VERSION 2

/* solver variants (which reference function the step reproduces) */
enum {
    DMPCB200_SOFT_BOUND = 0,    /* dmpc/matlab/solveSoftDMPCbound.m   (k_ctr = k,   slack >= -0.05) */
    DMPCB200_SOFT_BOUND2 = 1,   /* dmpc/matlab/solveSoftDMPCbound2.m  (k_ctr = k-1, slack >= -0.01) */
    DMPCB200_HARD = 2,          /* dmpc/matlab/solveHardDMPC.m         (all k, dist < 1, no slack)  */
    DMPCB200_HARD_ONDEMAND = 3  /* dmpc/matlab/solveHardDMPCOnDemand.m (first violating k, no slack)*/
};

/* per-agent status word: low byte = flags, bits 8..15 = number of infeasible-retries taken */
enum {
    DMPCB200_ST_SOLVED = 1,      /* p, v, a valid                                                  */
    DMPCB200_ST_COLL = 2,        /* k==1 violation deeper than coll_tol: reference returns coll=1   */
    DMPCB200_ST_INFEASIBLE = 4,  /* QP infeasible after all retries: reference returns feasible=0   */
    DMPCB200_ST_OUTBOUND = 8,    /* first predicted position outside workspace (is_inbounds.m)      */
    DMPCB200_ST_QPFAIL = 16,     /* internal: iteration cap                                         */
    DMPCB200_ST_OVERFLOW = 32    /* internal: constraint-row or active-set capacity exceeded        */
};

enum {
    DMPCB200_OK = 0,
    DMPCB200_ERR_ARG = -1,
    DMPCB200_ERR_CUDA = -2,
    DMPCB200_ERR_STATE = -3
};

/* Parameters.  Mirrors `struct Params` of dmpc/cpp/dmpc.h:50-63 plus the constants the MATLAB
 * scripts hard-code (solveSoftDMPCbound.m:25,43-52,78; CheckCollSoftDMPC.m:12; is_inbounds.m:2;
 * initDMPC.m:7; CollConstrHardDMPC.m:19; test/failure_rate.m:7-27).  dmpcb200_default_params
 * fills the reference's values. */
typedef struct dmpcb200_params {
    int32_t K;            /* k_hor, horizon length (<= 32)                                  */
    int32_t variant;      /* DMPCB200_SOFT_BOUND ...                                        */
    int32_t max_tries;    /* 30 (solveSoftDMPCbound.m:102)                                  */
    int32_t neigh_mode;   /* 0: dist < neigh_factor*rmin (MATLAB); 1: rmin*(1+k/K) (C++)    */
    double h;             /* time step                                                      */
    double rmin;          /* protection radius                                              */
    double c;             /* ellipsoid z scaling, E = diag(1,1,c)                           */
    double alim;          /* |a| bound                                                      */
    double Q1, S1;        /* weights when a collision constraint was added                  */
    double term;          /* linear slack penalty (negative)                                */
    double Q_far, Q_near; /* 1000 / 10000                                                   */
    double S_free;        /* 10                                                             */
    double near_radius;   /* 1.0                                                            */
    double slack_lb;      /* -0.05 (bound) / -0.01 (bound2)                                 */
    double neigh_factor;  /* 3.0                                                            */
    double coll_tol;      /* 0.05                                                           */
    double inb_tol;       /* 0.05                                                           */
    double hard_radius;   /* 1.0                                                            */
    double init_div;      /* 10                                                             */
    double goal_tol;      /* 0.01 (error_tol, test/failure_rate.m:24; ReachedGoal.m)        */
} dmpcb200_params;

typedef struct dmpcb200_handle dmpcb200_t;

/* per-agent diagnostics written by the QP kernel (optional output) */
typedef struct dmpcb200_diag {
    int32_t kstar;    /* 1-based first violating horizon step, 0 = none      */
    int32_t nv;       /* collision rows handed to the QP                     */
    int32_t iters;    /* dual active-set iterations (all tries)              */
    int32_t nact;     /* active-set size at the optimum                      */
} dmpcb200_diag;

int dmpcb200_abi_version(void);
const char* dmpcb200_last_error(void);
int dmpcb200_device_count(void);

/* defaults = the reference scripts' values for the given variant */
void dmpcb200_default_params(dmpcb200_params* p, int variant);

/* The semantics of the C++ port (solveQPv2, dmpc/cpp/dmpc.cpp:803-1287) as a parameter preset: k_ctr = k +
 * k_factor (0 -> SOFT_BOUND, -1 -> SOFT_BOUND2 incl. the `k == 0` skip, :516,:890), neighbour threshold
 * rmin (1 + k/k_hor) (:418), slack bound lim = 0.01 doubled together with the penalty for at most 20 retries
 * (:1077-1109: max_tries = 21 solves), penalty `int term = -1e6` (:846), hard-coded collision weights 1000 /
 * 100 (:940-945), struct Params defaults (dmpc.h:65-67: h 0.2, k_hor 12, c 1.5, rmin 0.5, alim 2, goal_tol
 * 0.05).  Acceleration limits as rows instead of bounds (:993-996) describe the same feasible set.  NOT
 * mirrored: the overflow of the 32-bit `term` after 11 doublings (undefined behaviour in the reference).
 * Un-commanded agents as static obstacles (:1633-1649): dmpcb200_set_static_obstacles below. */
void dmpcb200_default_params_cpp(dmpcb200_params* p, int k_factor);

/* getPosMat.m:1-23, getDeltaMat.m:1-9, precompute dmpc_soft_bound.m:81-108
 * (C++: get_lambda_A_v_mat dmpc.cpp:83-108, get_delta_mat :110-134, get_A0_mat :136-155).
 * Host computation (runs once per scenario in the reference too). Any output may be NULL.
 * A_p, A_v, Delta: 3K x 3K; A_initp: 3K x 6; column-major. */
int dmpcb200_model_mats(double h, int K, double* A_p, double* A_v, double* A_initp, double* Delta);

/* Create a solver for N agents (whole swarm) of which this handle solves agents [n0, n1)
 * on CUDA device `device` (one handle per GPU/process; dmpc.cpp:1600-1625 clusters).  n0 == n1 (an empty
 * block) is valid: such a handle only holds a replica of the horizon buffer.
 * n_scenarios >= 1: number of INDEPENDENT swarms of N agents solved side by side by every kernel launch --
 * the trial loops of test/failure_rate.m:61-68 (50 random trials per swarm size) as one batch.  Scenarios
 * never interact (own horizons, own workspace box, own loop control); they are sharded WHOLE across GPUs
 * (one handle with n0 = 0, n1 = N per GPU, no collective).  n_scenarios = 1 is the single swarm.
 * max_rows: capacity of the per-agent constraint-row buffer, 0 = automatic. */
int dmpcb200_create(const dmpcb200_params* p, int N, int n0, int n1, int n_scenarios, int device, int max_rows,
                    dmpcb200_t** out);
void dmpcb200_destroy(dmpcb200_t* h);

/* workspace box (set_boundaries dmpc.h:126) and goals (set_final_pts dmpc.h:130); pf is 3 x N */
int dmpcb200_set_bounds(dmpcb200_t* h, const double* pmin, const double* pmax);
int dmpcb200_set_goals(dmpcb200_t* h, const double* pf);

/* Agents [n_cmd, N) are STATIC OBSTACLES (the C++ port's N_cmd < N, dmpc.cpp:1633-1649: set_final_pts with fewer
 * columns than set_initial_pts): they are never solved, their horizon is constant at their start point, the goal
 * test and the failure scan cover agents [0, n_cmd) only.  Single-scenario handle with n0 = 0, n1 = N; call
 * before dmpcb200_init_horizons.  set_goals still takes 3 x N (the obstacles' columns are ignored). */
int dmpcb200_set_static_obstacles(dmpcb200_t* h, int n_cmd);

/* initDMPC.m:1-13 for all N agents on the device: l (3 x K x N), p1,v1,a1 (3 x N) host outputs
 * (any may be NULL); also seeds the device-resident state for dmpcb200_run. po is 3 x N. */
int dmpcb200_init_horizons(dmpcb200_t* h, const double* po, double* l, double* p1, double* v1,
                           double* a1);

/* One Jacobi MPC step for agents [n0,n1): the body of `for n = 1:N` in test/failure_rate.m:100-119
 * / dmpc_soft_bound.m:116-135 (C++ cluster_solvev2 dmpc.cpp:1792-1841), HOST buffers.
 * Inputs: pk,vk,ak 3 x N current states (only columns n0..n1-1 are read), l_prev 3 x K x N.
 * Outputs (rows of agents n0..n1-1 written, others untouched; any may be NULL):
 *   l_new 3 x K x N, p1,v1,a1 3 x N (first columns), v_hor,a_hor 3 x K x N (full v,a horizons),
 *   status[N], diag[N].  Agents that are not SOLVED keep l_new(:,:,n) = l_prev(:,:,n).
 * first_fail: lowest agent index without SOLVED or with OUTBOUND, -1 if none.
 * Caller arrays that are page-locked (cudaHostAlloc / cudaHostRegister, MATLAB: none) are used in place: DMA
 * straight from the inputs, kernel writes straight into l_new / p1 / v1 / a1; pageable arrays go through the
 * handle's pinned staging block (one copy in, none out).  The call returns when the results are in the arrays. */
int dmpcb200_step(dmpcb200_t* h, const double* pk, const double* vk, const double* ak,
                  const double* l_prev, double* l_new, double* p1, double* v1, double* a1,
                  double* v_hor, double* a_hor, int32_t* status, dmpcb200_diag* diag,
                  int32_t* first_fail);

/* The same call with the argument list bound once (a tight closed loop over preallocated, ideally pinned,
 * caller arrays: MATLAB would hold them in the MEX gateway's persistent state).  bind returns a slot (0..63),
 * step_bound runs dmpcb200_step on the arrays bound to it.  The arrays stay owned by the caller. */
int dmpcb200_bind_step(dmpcb200_t* h, const double* pk, const double* vk, const double* ak,
                       const double* l_prev, double* l_new, double* p1, double* v1, double* a1,
                       double* v_hor, double* a_hor, int32_t* status, dmpcb200_diag* diag, int32_t* slot);
int dmpcb200_step_bound(dmpcb200_t* h, int32_t slot, int32_t* first_fail);

/* Same step on DEVICE pointers, asynchronous on `stream` (a cudaStream_t passed as void*).
 * State arrays are 3 x N, horizons 3 x K x N, on the handle's device.  No host sync. */
int dmpcb200_step_dev(dmpcb200_t* h, const double* d_pk, const double* d_vk, const double* d_ak,
                      const double* d_l_prev, double* d_l_new, double* d_p1, double* d_v1,
                      double* d_a1, double* d_v_hor, double* d_a_hor, int32_t* d_status,
                      dmpcb200_diag* d_diag, void* stream);

/* ReachedGoal.m:1-11 / reached_goalv2 dmpc.cpp:1868-1882 on device pointers:
 * d_out[0] = max_n ||p(:,n)-pf(:,n)||, d_out[1] = (that < goal_tol).  Agent n's position is at
 * d_p + ld*n: ld = 3 for a packed 3 x N array, ld = 3K to read the first column of every horizon
 * of an l buffer (after the all-gather every rank holds all of them). */
int dmpcb200_goal_dev(dmpcb200_t* h, const double* d_p, int ld, double* d_out, void* stream);

/* ReachedGoal.m:1-11 on HOST arrays: p, pf 3 x N; *pass = (max_n ||p-pf|| < tol). */
int dmpcb200_reached_goal(dmpcb200_t* h, const double* p, const double* pf, double tol,
                          double* max_dist, int32_t* pass);

/* Device-resident closed loop for single-GPU handles (n0 = 0, n1 = N): the
 * `while ~reached_goal && k < max_K` loop of test/failure_rate.m:99-127 after
 * dmpcb200_init_horizons (or dmpcb200_set_state).  Runs at most max_steps further MPC steps with
 * no host synchronisation inside the loop (CUDA graph + device control word); stops at the goal
 * (ReachedGoal.m) and, if stop_on_fail = 1, at the first step in which an agent fails (any failure status);
 * stop_on_fail = 2 stops only when an agent's QP is infeasible -- the reference's driver (test/failure_rate.m:
 * 112-124 breaks the trial on ~feasible; `coll` / `outbound` leave feasible = 1 in solveSoftDMPCbound.m:29-30,
 * 125-128 and the trial goes on); otherwise agents that fail keep their state.
 * mode: 0 = CUDA graph; bit 0 = per-kernel CUDA-event timing (plain launches);
 *       bit 1 = plain launches without per-kernel events.
 * traj_p/v/a: optional host outputs 3 x (max_steps+1) x N (column 0 = initial state).
 * status_hist: optional max_steps x N.  *steps_done, *reached, *first_fail_step (-1 none). */
int dmpcb200_run(dmpcb200_t* h, int max_steps, int stop_on_fail, int mode, double* traj_p,
                 double* traj_v, double* traj_a, int32_t* status_hist, int32_t* steps_done,
                 int32_t* reached, int32_t* first_fail_step, int32_t* first_fail_agent);

/* ---- scenario batching (n_scenarios of dmpcb200_create) --------------------------------------------
 * set_scenario: start points po, goals pf (3 x N), workspace box of scenario s, and initDMPC.m:1-13 for its
 * agents (the scenario's loop state is reset).  On a single-scenario handle (s = 0) it is
 * set_bounds + set_goals + init_horizons in one call. */
int dmpcb200_set_scenario(dmpcb200_t* h, int s, const double* po, const double* pf, const double* pmin,
                          const double* pmax);
/* Scenario generation ON THE DEVICE for all n_scenarios of the handle at once (one CTA per scenario):
 * mode 0 = randomTest.m:1-57 (start and goal sets drawn independently, pairwise ellipsoidal distance
 * ||E1 (p - q)|| > rmin_init, a point that cannot be placed in 200000 tries restarts its set, :9-27; C++:
 * gen_rand_pts dmpc.cpp:188-227), mode 1 = randomExchange.m:1-56 (Euclidean distance, goals = a random
 * permutation of the starts in which every agent moves; C++: gen_rand_perm :229-265).  All scenarios share the
 * arena pmin / pmax (test/failure_rate.m:63-64).  The random stream is counter based (splitmix64 of seed,
 * scenario, set, draw index): reproducible, and restated bit for bit by the CPU oracle -- MATLAB's own stream is
 * never seeded by the reference and cannot be.  Also sets the bounds and goals and runs initDMPC.m: the handle
 * is ready for dmpcb200_run_batch (or dmpcb200_run when n_scenarios = 1).
 * po_out, pf_out: optional host copies, 3 x N x n_scenarios.  At most 4096 agents per scenario. */
int dmpcb200_gen_scenarios(dmpcb200_t* h, uint64_t seed, int mode, double rmin_init, const double* pmin,
                           const double* pmax, double* po_out, double* pf_out);

/* The closed loop of EVERY scenario, test/failure_rate.m:99-133 per trial: at most max_steps MPC steps; a
 * scenario stops at its goal (ReachedGoal.m) and, if stop_on_fail, at its first failing agent (the reference
 * `break`s the trial, failure_rate.m:112-123); the others go on.  Three launches per step for the whole batch
 * (scan, QP, per-scenario tail), CUDA graph of two steps, no host synchronisation inside the loop.
 * mode: 0 = CUDA graph, non-zero = plain launches.
 * Per-scenario outputs (any may be NULL): steps_done, reached, first_fail_step (-1 none), first_fail_agent,
 * goal_dist (max_n ||p - pf|| after the last step): n_scenarios entries each.
 * traj_p/v/a: optional, n_scenarios blocks of 3 x (max_steps+1) x N (column 0 = initial state). */
int dmpcb200_run_batch(dmpcb200_t* h, int max_steps, int stop_on_fail, int mode, double* traj_p, double* traj_v,
                       double* traj_a, int32_t* steps_done, int32_t* reached, int32_t* first_fail_step,
                       int32_t* first_fail_agent, double* goal_dist);
/* loop state of scenario s (any output may be NULL): l 3 x K x N, pk,vk,ak 3 x N, status / diag of its agents */
int dmpcb200_get_scenario(dmpcb200_t* h, int s, double* l, double* pk, double* vk, double* ak, int32_t* status,
                          dmpcb200_diag* diag);
/* device time (CUDA events) of the last dmpcb200_run_batch and the agent-steps it solved (sum over the
 * scenarios of steps_done x N) */
int dmpcb200_last_batch_timing(dmpcb200_t* h, double* total_ms, int64_t* agent_steps);

/* read / overwrite the device-resident loop state (any output may be NULL):
 * l 3 x K x N, pk,vk,ak 3 x N, status / diag of the last step. */
int dmpcb200_get_state(dmpcb200_t* h, double* l, double* pk, double* vk, double* ak,
                       int32_t* status, dmpcb200_diag* diag);
int dmpcb200_set_state(dmpcb200_t* h, const double* l, const double* pk, const double* vk,
                       const double* ak);

/* ---- per-agent drop-ins (batch of one; keep the reference's helper semantics) ------------- */

/* solveSoftDMPCbound.m / solveSoftDMPCbound2.m / solveHardDMPC.m / solveHardDMPCOnDemand.m
 * (selected by params.variant at create): po,pf,vo,ao 3-vectors, n 0-based agent index,
 * l 3 x K x N.  p,v,a: 3 x K outputs.  *status as above. */
int dmpcb200_solve_agent(dmpcb200_t* h, const double* po, const double* pf, const double* vo,
                         const double* ao, int n, const double* l, double* p, double* v,
                         double* a, int32_t* status, dmpcb200_diag* diag);

/* CheckCollSoftDMPC.m:1-17 (C++ check_collisionsv2 dmpc.cpp:395-448): p3 = own predicted
 * position at horizon step k (1-based), l 3 x K x N.  violation/viol_constr: N bytes. */
int dmpcb200_check_coll(dmpcb200_t* h, const double* p3, const double* l, int n, int k,
                        uint8_t* violation, uint8_t* viol_constr, double* min_dist,
                        int32_t* any_violation);

/* CollConstrSoftDMPC.m / CollConstrSoftDMPC2.m / CollConstrHardDMPC.m / ...OnDemand.m
 * (C++ build_collconstraintv2 dmpc.cpp:496-546): dense rows like the reference.
 * Ain: cap x 3K column-major (leading dimension cap), bin, prev_dist: cap.  mask = viol_constr
 * (N bytes; ignored for the HARD variant).  *nrows rows written, neighbour order ascending.
 * n = -1: no own agent, every column of l is an obstacle (dec-iSCP/CollConstr.m:1-23 with the SOFT_BOUND2
 * variant: rows on block k-1, right-hand side without the velocity term when vo = 0). */
int dmpcb200_coll_constr(dmpcb200_t* h, const double* p3, const double* po, const double* vo,
                         int n, int k, const double* l, const uint8_t* mask, int cap, double* Ain,
                         double* bin, double* prev_dist, int32_t* nrows);

/* propStatedmpc.m:1-8 for a batch of B agents on the device: a 3K x B -> p, v 3K x B */
int dmpcb200_prop_state(dmpcb200_t* h, int B, const double* po, const double* vo, const double* a,
                        double* p, double* v);

/* Post-processing of a finished transition -- the rest of the reference's t_dmpc, test/failure_rate.m:134-195
 * (C++: dmpc.cpp:1912-2086): time scaling of the trajectory to the velocity / acceleration limits (:136-162,
 * r_factor, h_scaled), 100 Hz cubic-spline interpolation (MATLAB `spline`, not-a-knot, :164-168), the O(N^2 T)
 * pairwise collision check on the interpolated positions (:170-181), travelled distance (:183) and trajectory
 * time (:185-194).
 * pk, vk, ak: HOST 3 x S x N as the MPC loop left them (column k = state after step k; S >= 4); they are scaled
 * IN PLACE like the reference does.  p, v, a: optional HOST outputs 3 x nt_cap x N for the interpolated
 * trajectories (written only if nt <= nt_cap; pass NULL / 0 to get the figures only -- res->nt tells the size).
 * time_index: optional int32[N] (failure_rate.m:186-193). */
typedef struct dmpcb200_post {
    double r_factor, h_scaled, T;  /* failure_rate.m:144,145,148 */
    double min_dist;               /* min over pairs and samples of ||E1 (p_i - p_j)||  (:174-175) */
    double totdist, traj_time;     /* :183, :194 */
    int32_t nt;                    /* number of 100 Hz samples, length(0:Ts:T) */
    int32_t violation;             /* min_dist < rmin - coll_tol  (:176-179) */
    double device_ms;              /* device time of the whole post-processing (CUDA events) */
} dmpcb200_post;
int dmpcb200_postprocess(dmpcb200_t* h, int S, double* pk, double* vk, double* ak, double vmax, double amax,
                         double Ts, double goal_radius, double* p, double* v, double* a, int nt_cap,
                         int32_t* time_index, dmpcb200_post* res);

/* ---- the reference's on-disk trajectory format (host code) ------------------------------------------
 * DMPC::trajectories2file (dmpc/cpp/dmpc.cpp:2088-2126), read back by dmpc/cpp_results/read_result.m:1-42:
 *   line 1: N N_cmd h_scaled pmin(3) pmax(3); po (3 lines of N); pf (3 lines of N_cmd); then per commanded
 *   agent 3 lines of T positions, then all velocities, then all accelerations -- every matrix in Eigen's default
 *   stream format (6 significant digits, columns right-aligned to the widest coefficient of that matrix).
 * po 3 x N, pf 3 x N_cmd, pos / vel / acc 3 x T x N_cmd, column-major like everything else here.
 * read: first call with the array pointers NULL returns the sizes (*N, *N_cmd, *T), second call fills them. */
int dmpcb200_write_trajectories(const char* path, int N, int N_cmd, int T, double h_scaled, const double* pmin,
                                const double* pmax, const double* po, const double* pf, const double* pos,
                                const double* vel, const double* acc);
int dmpcb200_read_trajectories(const char* path, int32_t* N, int32_t* N_cmd, int32_t* T, double* h_scaled,
                               double* pmin, double* pmax, double* po, double* pf, double* pos, double* vel,
                               double* acc);
/* one rows x cols column-major matrix in that stream format into buf (NUL-terminated, truncated to cap);
 * returns the full length */
int dmpcb200_format_matrix(int rows, int cols, const double* col_major, char* buf, int cap);

/* the same for scenario `scen` of a batched handle (goals of that scenario) */
int dmpcb200_postprocess_scenario(dmpcb200_t* h, int scen, int S, double* pk, double* vk, double* ak, double vmax,
                                  double amax, double Ts, double goal_radius, double* p, double* v, double* a,
                                  int nt_cap, int32_t* time_index, dmpcb200_post* res);

/* timing of the last dmpcb200_step / dmpcb200_run, CUDA events on the launch stream:
 * ms[0] = neighbour-scan kernel, ms[1] = QP kernel, ms[2] = whole step (device), averaged over
 * the steps of the call; launches[0] = number of kernels launched. */
int dmpcb200_last_timing(dmpcb200_t* h, double* ms, int64_t* launches);

/* host-side phases of the last dmpcb200_step, microseconds: us4[0] pack the inputs into the pinned staging
 * block, us4[1] submit (one H2D copy, two kernels, one D2H copy), us4[2] wait for the stream, us4[3] hand the
 * rows to the caller's arrays. */
int dmpcb200_last_host_timing(dmpcb200_t* h, double* us4);

/* raw device pointers of the handle's resident state (for host frameworks that own streams):
 * which: 0 l_cur, 1 l_next, 2 pk, 3 vk, 4 ak, 5 pf, 6 status, 7 goal_out(2 doubles),
 *        8 pk_next, 9 vk_next, 10 ak_next, 11 diag */
void* dmpcb200_device_ptr(dmpcb200_t* h, int which);
/* swap l_cur/l_next after a step + exchange (Jacobi `l = new_l`, failure_rate.m:124) */
int dmpcb200_swap_horizons(dmpcb200_t* h);

/* launch configuration chosen at create: out8 = {agents (warps) per QP block, QMAX (on-chip
 * active-set capacity), RCAP (rows held on chip), RMAX (row capacity), QBIG (rescue capacity),
 * rescue slots, QP kernel dynamic shared memory bytes, scan kernel shared memory bytes} */
int dmpcb200_config(dmpcb200_t* h, int32_t* out8);

#ifdef __cplusplus
}
#endif
#endif

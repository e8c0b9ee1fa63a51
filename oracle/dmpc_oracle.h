/*
 * dmpc_oracle.h -- CPU fp64 ORACLE for the DMPC per-agent QP hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path
 * (multiagent_planning_b200/ + libdmpc_b200.so) never calls it.
 *
 * It restates, in plain C, the algorithm of carlosluis/multiagent_planning
 * dmpc/matlab (file:line cited at every function in dmpc_oracle.c).  The QP
 * arithmetic itself is NOT in the reference (MATLAB quadprog / eigen-quadprog /
 * OOQP / CPLEX, none vendored, none pinned); it is replaced here by an exact
 * dense Goldfarb-Idnani dual active-set solver with a KKT certificate.  The
 * QPs are strictly convex, so the optimum is unique and solver independent.
 *
 * Parity pinning: matrix known-answers and the single-step known-answer
 * workspaces data/failure_rate/failure_rate2.mat (solveSoftDMPCbound) and
 * data/comp_kctr/comp_kctr_3.mat (solveSoftDMPCbound2) -> tests/golden/.
 * The hard variants have no saved reference data: parity unpinned for them.
 *
 * All arrays are column-major fp64 exactly like MATLAB: a horizon buffer
 * l is 3 x K x N, element (d,k,n) at l[d + 3*(k + K*n)], indices 0-based here.
 */
#ifndef DMPC_ORACLE_H
#define DMPC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    ORC_VARIANT_SOFT_BOUND = 0,   /* solveSoftDMPCbound.m  (k_ctr = k,   slack >= -0.05) */
    ORC_VARIANT_SOFT_BOUND2 = 1,  /* solveSoftDMPCbound2.m (k_ctr = k-1, slack >= -0.01) */
    ORC_VARIANT_HARD = 2,         /* solveHardDMPC.m          (all k, dist < 1, no slack)  */
    ORC_VARIANT_HARD_ONDEMAND = 3 /* solveHardDMPCOnDemand.m  (first violating k, no slack)*/
};

/* per-agent status word */
enum {
    ORC_ST_SOLVED = 1,      /* p,v,a are valid                                            */
    ORC_ST_COLL = 2,        /* k==1 violation deeper than coll_tol: reference returns []  */
    ORC_ST_INFEASIBLE = 4,  /* QP infeasible after all retries: reference returns []      */
    ORC_ST_OUTBOUND = 8,    /* first predicted position outside workspace (+inb_tol)      */
    ORC_ST_QPFAIL = 16      /* oracle-internal: iteration cap / numerical failure         */
};

typedef struct orc_params {
    int32_t K;            /* horizon length k_hor                                          */
    int32_t variant;      /* ORC_VARIANT_*                                                 */
    int32_t max_tries;    /* 30 (solveSoftDMPCbound.m:102)                                 */
    int32_t neigh_mode;   /* 0: dist < neigh_factor*rmin (MATLAB); 1: rmin*(1+k/K) (C++)   */
    double h;             /* time step                                                     */
    double rmin;          /* protection radius                                             */
    double c;             /* ellipsoid z scaling, E = diag(1,1,c)                          */
    double alim;          /* |a| bound                                                     */
    double Q1, S1;        /* weights when a collision constraint was added                 */
    double term;          /* linear slack penalty (negative)                               */
    double Q_far, Q_near; /* 1000 / 10000 (solveSoftDMPCbound.m:43-52)                     */
    double S_free;        /* 10                                                            */
    double near_radius;   /* 1.0: ||po-pf|| threshold between Q_far and Q_near             */
    double slack_lb;      /* -0.05 (bound) / -0.01 (bound2)                                */
    double neigh_factor;  /* 3.0 (CheckCollSoftDMPC.m:12)                                  */
    double coll_tol;      /* 0.05 (solveSoftDMPCbound.m:25)                                */
    double inb_tol;       /* 0.05 (is_inbounds.m:2)                                        */
    double hard_radius;   /* 1.0  (CollConstrHardDMPC.m:19)                                */
    double init_div;      /* 10   (initDMPC.m:7)                                           */
} orc_params;

/* per-solve diagnostics (all optional outputs) */
typedef struct orc_diag {
    int32_t kstar;     /* 1-based first violating horizon step, 0 = none                  */
    int32_t nv;        /* number of collision rows                                         */
    int32_t tries;     /* retries performed                                                */
    int32_t qp_iters;  /* active-set iterations of the last solve                          */
    double kkt_stat;   /* ||Hx+f+A'lam||_inf / (1+||f||_inf)                               */
    double kkt_prim;   /* max constraint violation                                         */
    double kkt_comp;   /* max |lam_i * slack_i|                                            */
    double kkt_dual;   /* most negative multiplier (as positive number)                    */
    double min_dist;   /* min neighbour distance at kstar                                  */
    double objective;  /* 1/2 x'Hx + f'x at the optimum                                    */
    int32_t n_act_box; /* active acceleration bounds at the optimum                        */
    int32_t n_act_pos; /* active workspace (position) rows                                 */
    int32_t n_act_row; /* active collision rows                                            */
    int32_t n_act_eps; /* slack variables strictly below 0                                 */
} orc_diag;

void orc_default_params(orc_params* p, int variant);

/* a1-a3: model matrices (getPosMat.m, getDeltaMat.m, dmpc_soft_bound.m:81-108) */
void orc_model_mats(double h, int K, double* A_p /*3K x 3K*/, double* A_v /*3K x 3K*/,
                    double* A_initp /*3K x 6*/, double* Delta /*3K x 3K*/);

/* a4: initDMPC.m */
void orc_init_dmpc(const double* po, const double* pf, double h, int K, double init_div,
                   double* p /*3xK*/, double* v, double* a);

/* a5: CheckCollSoftDMPC.m ; k is 1-based like MATLAB, n 0-based. Returns any(violation). */
int orc_check_coll(const orc_params* P, const double* p3, const double* l, int N, int n, int k,
                   uint8_t* violation /*N*/, uint8_t* viol_constr /*N*/, double* min_dist);

/* a6/a7: CollConstr{Soft,Soft2,Hard,HardOnDemand}DMPC.m  -> dense rows like the reference.
 * Ain is nrows x 3K column-major, returns nrows written.  For ORC_VARIANT_HARD the
 * zero (vacuous) rows the reference keeps are NOT emitted.  mask may be NULL for HARD. */
int orc_coll_constr(const orc_params* P, const double* p3, const double* po, const double* vo,
                    int n, int k, const double* l, int N, const uint8_t* mask,
                    double* Ain, int ld, double* bin, double* prev_dist, int32_t* neigh_idx);

/* a8-a10: one agent's solve. p,v,a are 3xK.  Returns the status word. */
int orc_solve_agent(const orc_params* P, const double* po, const double* pf, const double* vo,
                    const double* ao, int n, const double* l, int N, const double* pmin,
                    const double* pmax, double* p, double* v, double* a, orc_diag* diag);

/* a14: one Jacobi MPC step for agents [n0,n1) against the full horizon buffer l_prev.
 * pk,vk,ak: 3xN current state (first column of the previous solution); pf 3xN.
 * Outputs for agent n land at index n of l_new / p1 / v1 / a1 / status (other agents untouched).
 * nthreads>1 splits the agent range into contiguous clusters (dmpc.cpp:1600-1625).
 * Returns the lowest failing agent index (status without SOLVED or with OUTBOUND), or -1. */
int orc_step(const orc_params* P, int N, int n0, int n1, const double* pk, const double* vk,
             const double* ak, const double* pf, const double* l_prev, const double* pmin,
             const double* pmax, double* l_new, double* p1, double* v1, double* a1,
             int32_t* status, orc_diag* diags /*N or NULL*/, int nthreads);

/* a13: ReachedGoal.m */
int orc_reached_goal(const double* p /*3xN*/, const double* pf, int N, double tol, double* max_dist);

/* a12: is_inbounds.m */
int orc_is_inbounds(const double* p3, const double* pmin, const double* pmax, double tol);

/* a11: propStatedmpc.m (structured evaluation, same values as the dense product) */
void orc_prop_state(double h, int K, const double* po, const double* vo, const double* a /*3K*/,
                    double* p /*3K*/, double* v /*3K*/);

/* a16 stand-in: exact dense QP  min 1/2 x'Hx + f'x  s.t. A x <= b, lb <= x <= ub.
 * H n x n col-major (ld n), A m x n col-major with leading dimension lda.
 * lam: m + 2n multipliers (rows, lower bounds, upper bounds) or NULL.
 * returns 0 ok, 1 infeasible, 2 iteration cap / numerical failure. */
int orc_qp_gi(int n, const double* H, const double* f, int m, const double* A, int lda,
              const double* b, const double* lb, const double* ub, double* x, double* lam,
              int32_t* iters, orc_diag* kkt);

/* randomTest.m:1-57 (mode 0) / randomExchange.m:1-56 (mode 1) for scenario `scen` of a batch: po, pf 3 x N.
 * Counter-based random stream (splitmix64 of seed, scenario, set, draw index) -- see dmpc_oracle.c. */
void orc_gen_scenario(uint64_t seed, int scen, int mode, int N, const double* pmin, const double* pmax,
                      double rmin, double c, double* po, double* pf);

#ifdef __cplusplus
}
#endif
#endif

/*
 * dmpc_oracle.c -- CPU fp64 ORACLE (test infrastructure, see dmpc_oracle.h).
 *
 * Behavioural restatement of carlosluis/multiagent_planning dmpc/matlab.
 * Every function cites the reference file:line it follows.  Written from the
 * algorithm, not copied: MATLAB sources are scripts over dense matrices, this is
 * plain C over flat column-major arrays.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: the model matrices are
 * compared bit-for-bit with the reference's saved workspaces, so no FMA contraction).
 */
#include "dmpc_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* parameters                                                                */
/* ------------------------------------------------------------------------- */

/* Defaults = the values hard-coded in test/failure_rate.m:7-27,80-92 and
 * solveSoftDMPCbound.m:25,43-52,78 / solveSoftDMPCbound2.m:77 / dmpc_hard.m:75-80. */
void orc_default_params(orc_params* p, int variant) {
    memset(p, 0, sizeof(*p));
    p->K = 15;
    p->variant = variant;
    p->max_tries = 30;
    p->neigh_mode = 0;
    p->h = 0.2;
    p->rmin = 0.35;
    p->c = 2.0;
    p->alim = 1.0;
    p->Q1 = 1000.0;
    p->S1 = (variant == ORC_VARIANT_HARD || variant == ORC_VARIANT_HARD_ONDEMAND) ? 10.0 : 100.0;
    p->term = -5.0e4;
    p->Q_far = 1000.0;
    p->Q_near = 10000.0;
    p->S_free = 10.0;
    p->near_radius = 1.0;
    p->slack_lb = (variant == ORC_VARIANT_SOFT_BOUND2) ? -0.01 : -0.05;
    p->neigh_factor = 3.0;
    p->coll_tol = 0.05;
    p->inb_tol = 0.05;
    p->hard_radius = 1.0;
    p->init_div = 10.0;
}

/* ------------------------------------------------------------------------- */
/* a1-a3 model matrices                                                      */
/* ------------------------------------------------------------------------- */

/* getPosMat.m:1-23, dmpc_soft_bound.m:81-108 (A_p, A_v, A_initp), getDeltaMat.m:1-9.
 * The reference iterates new_row = Aux*prev_row + add_b with Aux = [I hI; 0 I] and
 * A_init = Aux*A_init.  Per scalar that is  p <- p + h*v (+ h^2/2 on the diagonal block),
 * v <- v (+ h on the diagonal block), t <- t + h; evaluated in that order without FMA it
 * reproduces the saved workspaces bit for bit (tests/test_oracle_golden.py). */
void orc_model_mats(double h, int K, double* A_p, double* A_v, double* A_initp, double* Delta) {
    const int n = 3 * K;
    double* pc = (double*)calloc((size_t)K, sizeof(double)); /* scalar (per-axis) rows */
    double* vc = (double*)calloc((size_t)K, sizeof(double));
    if (A_p) memset(A_p, 0, sizeof(double) * n * n);
    if (A_v) memset(A_v, 0, sizeof(double) * n * n);
    if (A_initp) memset(A_initp, 0, sizeof(double) * n * 6);
    if (Delta) memset(Delta, 0, sizeof(double) * n * n);
    double t = 0.0;
    const double hh2 = h * h / 2;
    for (int k = 0; k < K; ++k) {
        for (int j = 0; j < K; ++j) {
            volatile double hv = h * vc[j];
            pc[j] = pc[j] + hv;
        }
        pc[k] = pc[k] + hh2;
        vc[k] = vc[k] + h;
        t = t + h;
        for (int j = 0; j < K; ++j)
            for (int d = 0; d < 3; ++d) {
                if (A_p) A_p[(3 * k + d) + (size_t)n * (3 * j + d)] = pc[j];
                if (A_v) A_v[(3 * k + d) + (size_t)n * (3 * j + d)] = vc[j];
            }
        if (A_initp)
            for (int d = 0; d < 3; ++d) {
                A_initp[(3 * k + d) + (size_t)n * d] = 1.0;
                A_initp[(3 * k + d) + (size_t)n * (3 + d)] = t;
            }
    }
    if (Delta) {
        for (int i = 0; i < n; ++i) Delta[i + (size_t)n * i] = 1.0;
        for (int i = 3; i < n; ++i) Delta[i + (size_t)n * (i - 3)] = -1.0;
    }
    free(pc);
    free(vc);
}

/* scalar model coefficients used by the structured evaluations below:
 * lam[k][j] = A_p(3k+d,3j+d), t[k] = A_initp(3k+d,3+d), built exactly as above. */
typedef struct {
    int K;
    double h;
    double* lam; /* K x K row-major, lower triangular */
    double* tt;  /* K */
} orc_model;

static void model_build(orc_model* M, double h, int K) {
    M->K = K;
    M->h = h;
    M->lam = (double*)calloc((size_t)K * K, sizeof(double));
    M->tt = (double*)calloc((size_t)K, sizeof(double));
    double* pc = (double*)calloc((size_t)K, sizeof(double));
    double* vc = (double*)calloc((size_t)K, sizeof(double));
    double t = 0.0;
    const double hh2 = h * h / 2;
    for (int k = 0; k < K; ++k) {
        for (int j = 0; j < K; ++j) {
            volatile double hv = h * vc[j];
            pc[j] = pc[j] + hv;
        }
        pc[k] = pc[k] + hh2;
        vc[k] = vc[k] + h;
        t = t + h;
        for (int j = 0; j < K; ++j) M->lam[k * K + j] = pc[j];
        M->tt[k] = t;
    }
    free(pc);
    free(vc);
}
static void model_free(orc_model* M) {
    free(M->lam);
    free(M->tt);
}

/* propStatedmpc.m:1-8   p = A_p a + A_initp [po;vo],  v = A_v a + repmat(vo) */
static void prop_state_m(const orc_model* M, const double* po, const double* vo, const double* a,
                         double* p, double* v) {
    const int K = M->K;
    for (int k = 0; k < K; ++k)
        for (int d = 0; d < 3; ++d) {
            double sp = 0.0, sv = 0.0;
            for (int j = 0; j <= k; ++j) {
                sp += M->lam[k * K + j] * a[3 * j + d];
                sv += M->h * a[3 * j + d];
            }
            p[3 * k + d] = sp + (po[d] + M->tt[k] * vo[d]);
            v[3 * k + d] = sv + vo[d];
        }
}

void orc_prop_state(double h, int K, const double* po, const double* vo, const double* a,
                    double* p, double* v) {
    orc_model M;
    model_build(&M, h, K);
    prop_state_m(&M, po, vo, a, p, v);
    model_free(&M);
}

/* initDMPC.m:1-13 */
void orc_init_dmpc(const double* po, const double* pf, double h, int K, double init_div,
                   double* p, double* v, double* a) {
    for (int i = 0; i < K; ++i) {
        /* t = 0:h:(K-1)*h  -> t(i) = i*h (MATLAB colon: multiples of the increment) */
        double t = i * h;
        for (int d = 0; d < 3; ++d) {
            p[3 * i + d] = po[d] + 1 * t * (pf[d] - po[d]) / init_div;
            v[3 * i + d] = 0.0;
            a[3 * i + d] = 0.0;
        }
    }
}

/* is_inbounds.m:1-6 */
int orc_is_inbounds(const double* p, const double* pmin, const double* pmax, double tol) {
    int up = p[0] < pmax[0] + tol && p[1] < pmax[1] + tol && p[2] < pmax[2] + tol;
    int down = p[0] > pmin[0] - tol && p[1] > pmin[1] - tol && p[2] > pmin[2] - tol;
    return up && down;
}

/* ReachedGoal.m:1-11 */
int orc_reached_goal(const double* p, const double* pf, int N, double tol, double* max_dist) {
    double md = 0.0;
    for (int n = 0; n < N; ++n) {
        double s = 0.0;
        for (int d = 0; d < 3; ++d) {
            double e = p[3 * n + d] - pf[3 * n + d];
            s += e * e;
        }
        s = sqrt(s);
        if (s > md) md = s;
    }
    if (max_dist) *max_dist = md;
    return md < tol;
}

/* ------------------------------------------------------------------------- */
/* a5-a7 neighbour scan and constraint rows                                  */
/* ------------------------------------------------------------------------- */

/* norm(E1*(p-pj),2) with E1 = diag(1,1,1/c)  (CheckCollSoftDMPC.m:10) */
static inline double ell_dist(const double* p, const double* pj, double c, double* dvec) {
    double dx = p[0] - pj[0], dy = p[1] - pj[1], dz = p[2] - pj[2];
    if (dvec) {
        dvec[0] = dx;
        dvec[1] = dy;
        dvec[2] = dz;
    }
    double ez = dz / c;
    return sqrt(dx * dx + dy * dy + ez * ez);
}

static inline double neigh_threshold(const orc_params* P, int k1) {
    if (P->neigh_mode == 1) /* dmpc.cpp:418, k zero-based there */
        return P->rmin * (1.0 + (double)(k1 - 1) / P->K);
    return P->rmin * P->neigh_factor; /* CheckCollSoftDMPC.m:12 */
}

/* CheckCollSoftDMPC.m:1-17 */
int orc_check_coll(const orc_params* P, const double* p, const double* l, int N, int n, int k,
                   uint8_t* violation, uint8_t* viol_constr, double* min_dist) {
    const int K = P->K;
    const double thr = neigh_threshold(P, k);
    int any = 0;
    double md = INFINITY;
    for (int i = 0; i < N; ++i) {
        if (violation) violation[i] = 0;
        if (viol_constr) viol_constr[i] = 0;
        if (i == n) continue;
        const double* pj = l + 3 * ((k - 1) + (size_t)K * i);
        double dist = ell_dist(p, pj, P->c, NULL);
        int vi = dist < P->rmin;
        if (violation) violation[i] = (uint8_t)vi;
        if (viol_constr) viol_constr[i] = (uint8_t)(dist < thr);
        any |= vi;
        if (dist < md) md = dist;
    }
    if (min_dist) *min_dist = md;
    return any;
}

/* CollConstrSoftDMPC.m:1-32 (k_ctr = k), CollConstrSoftDMPC2.m:8 (k_ctr = k-1),
 * CollConstrHardDMPC.m:1-34 (all i != n with dist < hard_radius),
 * CollConstrHardDMPCOnDemand.m:1-32.  order = 2 only:
 *   dist = ||E1 d||, diff = E2 d, r = dist*(rmin - dist) + diff.p - diff.A_initp[k_ctr][po;vo]
 *   row = -[0.. diff@k_ctr ..0]*A_p,  b = -r */
static int coll_rows(const orc_params* P, const orc_model* M, const double* p, const double* po,
                     const double* vo, int n, int k, int k_ctr, const double* l, int N,
                     const uint8_t* mask, double* Ain, int ld, double* bin, double* prev_dist,
                     int32_t* neigh_idx, double* diff_out) {
    const int K = P->K;
    const double c2 = P->c * P->c;
    const int hard = (P->variant == ORC_VARIANT_HARD);
    int idx = 0;
    for (int i = 0; i < N; ++i) {
        if (i == n) continue;
        if (!hard && !(mask && mask[i])) continue;
        const double* pj = l + 3 * ((k - 1) + (size_t)K * i);
        double dv[3];
        double dist = ell_dist(p, pj, P->c, dv);
        if (hard && !(dist < P->hard_radius)) continue;
        double diff[3] = {dv[0], dv[1], dv[2] / c2};
        double dp = diff[0] * p[0] + diff[1] * p[1] + diff[2] * p[2];
        /* A_initp rows of block k_ctr (1-based): [I, t_kctr I] */
        double tk = M->tt[k_ctr - 1];
        double d0 = diff[0] * (po[0] + tk * vo[0]) + diff[1] * (po[1] + tk * vo[1]) +
                    diff[2] * (po[2] + tk * vo[2]);
        double r = dist * ((P->rmin) - dist + dp / dist) - d0;
        if (Ain)
            for (int j = 0; j < K; ++j)
                for (int d = 0; d < 3; ++d)
                    Ain[idx + (size_t)ld * (3 * j + d)] = -diff[d] * M->lam[(k_ctr - 1) * K + j];
        if (bin) bin[idx] = -r;
        if (prev_dist) prev_dist[idx] = dist;
        if (neigh_idx) neigh_idx[idx] = i;
        if (diff_out) {
            diff_out[3 * idx + 0] = diff[0];
            diff_out[3 * idx + 1] = diff[1];
            diff_out[3 * idx + 2] = diff[2];
        }
        ++idx;
    }
    return idx;
}

int orc_coll_constr(const orc_params* P, const double* p, const double* po, const double* vo,
                    int n, int k, const double* l, int N, const uint8_t* mask, double* Ain,
                    int ld, double* bin, double* prev_dist, int32_t* neigh_idx) {
    orc_model M;
    model_build(&M, P->h, P->K);
    int k_ctr = (P->variant == ORC_VARIANT_SOFT_BOUND2) ? k - 1 : k;
    int r = coll_rows(P, &M, p, po, vo, n, k, k_ctr, l, N, mask, Ain, ld, bin, prev_dist,
                      neigh_idx, NULL);
    model_free(&M);
    return r;
}

/* ------------------------------------------------------------------------- */
/* a16 stand-in: dense Goldfarb-Idnani dual active-set QP                     */
/* ------------------------------------------------------------------------- */
/*
 *   min 1/2 x'Hx + f'x   s.t.  A x <= b (m rows),  lb <= x <= ub
 * Internally every constraint is  c_i'x >= d_i :
 *   i in [0,m)        c = -A(i,:)',  d = -b_i
 *   i in [m,m+n)      c = +e_j,      d = lb_j
 *   i in [m+n,m+2n)   c = -e_j,      d = -ub_j
 * D. Goldfarb, A. Idnani, "A numerically stable dual method for solving strictly
 * convex quadratic programs", Math. Prog. 27 (1983) -- the method behind the
 * eigen-quadprog the C++ reference calls (dmpc.cpp:1068-1072).
 */
typedef struct {
    int n, m, lda;
    const double* A;
    const double* b;
    const double* lb;
    const double* ub;
} gi_cons;

static inline double cons_rhs(const gi_cons* C, int i) {
    if (i < C->m) return -C->b[i];
    if (i < C->m + C->n) return C->lb[i - C->m];
    return -C->ub[i - C->m - C->n];
}
/* c_i . v */
static inline double cons_dot(const gi_cons* C, int i, const double* v) {
    if (i < C->m) {
        double s = 0.0;
        const double* a = C->A + i;
        for (int j = 0; j < C->n; ++j) s -= a[(size_t)C->lda * j] * v[j];
        return s;
    }
    if (i < C->m + C->n) return v[i - C->m];
    return -v[i - C->m - C->n];
}
/* out = J' c_i  (J n x n col-major) */
static void cons_Jt(const gi_cons* C, int i, const double* J, double* out) {
    const int n = C->n;
    if (i < C->m) {
        const double* a = C->A + i;
        for (int col = 0; col < n; ++col) {
            const double* Jc = J + (size_t)n * col;
            double s = 0.0;
            for (int j = 0; j < n; ++j) s -= a[(size_t)C->lda * j] * Jc[j];
            out[col] = s;
        }
    } else if (i < C->m + n) {
        int j = i - C->m;
        for (int col = 0; col < n; ++col) out[col] = J[j + (size_t)n * col];
    } else {
        int j = i - C->m - n;
        for (int col = 0; col < n; ++col) out[col] = -J[j + (size_t)n * col];
    }
}
static inline double cons_norm(const gi_cons* C, int i) {
    if (i >= C->m) return 1.0;
    double s = 0.0;
    const double* a = C->A + i;
    for (int j = 0; j < C->n; ++j) s += a[(size_t)C->lda * j] * a[(size_t)C->lda * j];
    return sqrt(s);
}

static inline void givens(double a, double b, double* c, double* s, double* r) {
    double h = hypot(a, b);
    if (h == 0.0) {
        *c = 1.0;
        *s = 0.0;
        *r = 0.0;
    } else {
        *c = a / h;
        *s = b / h;
        *r = h;
    }
}
static inline void rot_cols(double* J, int n, int j0, int j1, double c, double s) {
    double* a = J + (size_t)n * j0;
    double* b = J + (size_t)n * j1;
    for (int i = 0; i < n; ++i) {
        double ta = a[i], tb = b[i];
        a[i] = c * ta + s * tb;
        b[i] = -s * ta + c * tb;
    }
}

int orc_qp_gi(int n, const double* H, const double* f, int m, const double* A, int lda,
              const double* b, const double* lb, const double* ub, double* x, double* lam,
              int32_t* iters_out, orc_diag* kkt) {
    const int mt = m + 2 * n;
    gi_cons C = {n, m, lda, A, b, lb, ub};
    int rc = 2;
    double* L = (double*)malloc(sizeof(double) * n * n);
    double* J = (double*)calloc((size_t)n * n, sizeof(double));
    double* R = (double*)calloc((size_t)n * n, sizeof(double));
    double* dv = (double*)malloc(sizeof(double) * n);
    double* z = (double*)malloc(sizeof(double) * n);
    double* r = (double*)malloc(sizeof(double) * n);
    double* u = (double*)calloc((size_t)n + 1, sizeof(double));
    double* cn = (double*)malloc(sizeof(double) * mt);
    int* act = (int*)malloc(sizeof(int) * (n + 1));
    uint8_t* isact = (uint8_t*)calloc((size_t)mt, 1);
    int q = 0, iters = 0;

    /* Cholesky H = L L' */
    memcpy(L, H, sizeof(double) * n * n);
    for (int j = 0; j < n; ++j) {
        double djj = L[j + (size_t)n * j];
        for (int k = 0; k < j; ++k) djj -= L[j + (size_t)n * k] * L[j + (size_t)n * k];
        if (!(djj > 0.0)) goto done;
        djj = sqrt(djj);
        L[j + (size_t)n * j] = djj;
        for (int i = j + 1; i < n; ++i) {
            double s = L[i + (size_t)n * j];
            for (int k = 0; k < j; ++k) s -= L[i + (size_t)n * k] * L[j + (size_t)n * k];
            L[i + (size_t)n * j] = s / djj;
        }
    }
    /* J = L^{-T}: column c of L^{-1} by forward substitution, stored as row c of J */
    for (int c = 0; c < n; ++c) {
        /* solve L y = e_c */
        for (int i = 0; i < n; ++i) {
            double s = (i == c) ? 1.0 : 0.0;
            for (int k = c; k < i; ++k) s -= L[i + (size_t)n * k] * z[k];
            z[i] = (i < c) ? 0.0 : s / L[i + (size_t)n * i];
        }
        /* y = L^{-1}(:,c) -> J(c,:) = y' */
        for (int i = 0; i < n; ++i) J[c + (size_t)n * i] = z[i];
    }
    /* unconstrained optimum x = -H^{-1} f = -J J' f */
    for (int i = 0; i < n; ++i) {
        double s = 0.0;
        for (int k = 0; k < n; ++k) s += J[k + (size_t)n * i] * f[k];
        dv[i] = s;
    }
    for (int i = 0; i < n; ++i) {
        double s = 0.0;
        for (int k = 0; k < n; ++k) s += J[i + (size_t)n * k] * dv[k];
        x[i] = -s;
    }
    for (int i = 0; i < mt; ++i) cn[i] = cons_norm(&C, i);

    const double feas_tol = 1e-11;
    const int max_iter = 20 * (n + mt) + 100;
    for (;;) {
        /* step 1: most violated constraint (normalised) */
        int p = -1;
        double worst = -feas_tol;
        for (int i = 0; i < mt; ++i) {
            if (isact[i] || cn[i] == 0.0) continue;
            double rhs = cons_rhs(&C, i);
            if (!isfinite(rhs)) continue;
            double s = (cons_dot(&C, i, x) - rhs) / cn[i];
            if (s < worst) {
                worst = s;
                p = i;
            }
        }
        if (p < 0) {
            rc = 0;
            break;
        }
        double sp = cons_dot(&C, p, x) - cons_rhs(&C, p);
        double uplus = 0.0;
        for (;;) { /* step 2 */
            if (++iters > max_iter) goto done;
            cons_Jt(&C, p, J, dv);
            double dn2 = 0.0, d2n2 = 0.0;
            for (int i = 0; i < n; ++i) dn2 += dv[i] * dv[i];
            for (int i = q; i < n; ++i) d2n2 += dv[i] * dv[i];
            /* z = J2 d2 */
            for (int i = 0; i < n; ++i) {
                double s = 0.0;
                for (int k = q; k < n; ++k) s += J[i + (size_t)n * k] * dv[k];
                z[i] = s;
            }
            /* r = R^{-1} d1 */
            for (int i = q - 1; i >= 0; --i) {
                double s = dv[i];
                for (int k = i + 1; k < q; ++k) s -= R[i + (size_t)n * k] * r[k];
                r[i] = s / R[i + (size_t)n * i];
            }
            int dependent = !(d2n2 > 1e-20 * dn2) || q == n;
            double t2 = dependent ? INFINITY : -sp / d2n2; /* z'c_p = ||d2||^2 */
            double t1 = INFINITY;
            int ldrop = -1;
            for (int k = 0; k < q; ++k)
                if (r[k] > 0.0) {
                    double t = u[k] / r[k];
                    if (t < t1) {
                        t1 = t;
                        ldrop = k;
                    }
                }
            double t = t1 < t2 ? t1 : t2;
            if (!isfinite(t)) {
                rc = 1; /* infeasible */
                goto done;
            }
            if (!dependent) {
                for (int i = 0; i < n; ++i) x[i] += t * z[i];
            }
            for (int k = 0; k < q; ++k) u[k] -= t * r[k];
            uplus += t;
            if (!dependent && t2 <= t1) {
                /* full step: add constraint p */
                for (int j = n - 1; j > q; --j) {
                    double c, s, h;
                    givens(dv[j - 1], dv[j], &c, &s, &h);
                    dv[j - 1] = h;
                    dv[j] = 0.0;
                    rot_cols(J, n, j - 1, j, c, s);
                }
                for (int i = 0; i <= q; ++i) R[i + (size_t)n * q] = dv[i];
                act[q] = p;
                u[q] = uplus;
                isact[p] = 1;
                ++q;
                break; /* back to step 1 */
            }
            /* partial (or pure dual) step: drop constraint at position ldrop */
            {
                int l0 = ldrop;
                isact[act[l0]] = 0;
                for (int k = l0; k < q - 1; ++k) {
                    act[k] = act[k + 1];
                    u[k] = u[k + 1];
                    for (int i = 0; i <= k + 1; ++i) R[i + (size_t)n * k] = R[i + (size_t)n * (k + 1)];
                }
                --q;
                for (int k = l0; k < q; ++k) {
                    double c, s, h;
                    givens(R[k + (size_t)n * k], R[k + 1 + (size_t)n * k], &c, &s, &h);
                    R[k + (size_t)n * k] = h;
                    R[k + 1 + (size_t)n * k] = 0.0;
                    for (int j = k + 1; j < q; ++j) {
                        double ta = R[k + (size_t)n * j], tb = R[k + 1 + (size_t)n * j];
                        R[k + (size_t)n * j] = c * ta + s * tb;
                        R[k + 1 + (size_t)n * j] = -s * ta + c * tb;
                    }
                    rot_cols(J, n, k, k + 1, c, s);
                }
                for (int i = 0; i <= q; ++i) R[i + (size_t)n * q] = 0.0;
                if (!dependent) sp = cons_dot(&C, p, x) - cons_rhs(&C, p);
            }
        }
    }
done:
    if (iters_out) *iters_out = iters;
    if (rc == 0) {
        /* KKT certificate, computed from x and the multipliers only */
        double* g = dv;
        double fn = 0.0;
        for (int i = 0; i < n; ++i) {
            double s = f[i];
            for (int k = 0; k < n; ++k) s += H[i + (size_t)n * k] * x[k];
            g[i] = s;
            if (fabs(f[i]) > fn) fn = fabs(f[i]);
        }
        double obj = 0.0;
        for (int i = 0; i < n; ++i) obj += 0.5 * x[i] * (g[i] + f[i]);
        if (lam) memset(lam, 0, sizeof(double) * mt);
        double comp = 0.0, dual = 0.0;
        for (int k = 0; k < q; ++k) {
            int i = act[k];
            if (lam) lam[i] = u[k];
            if (-u[k] > dual) dual = -u[k];
            double s = cons_dot(&C, i, x) - cons_rhs(&C, i);
            if (fabs(s * u[k]) > comp) comp = fabs(s * u[k]);
            /* g -= u_k c_i */
            if (i < m) {
                for (int j = 0; j < n; ++j) g[j] += u[k] * A[i + (size_t)lda * j];
            } else if (i < m + n) {
                g[i - m] -= u[k];
            } else {
                g[i - m - n] += u[k];
            }
        }
        double stat = 0.0, prim = 0.0;
        for (int i = 0; i < n; ++i)
            if (fabs(g[i]) > stat) stat = fabs(g[i]);
        for (int i = 0; i < mt; ++i) {
            double rhs = cons_rhs(&C, i);
            if (!isfinite(rhs)) continue;
            double s = cons_dot(&C, i, x) - rhs;
            if (-s > prim) prim = -s;
        }
        if (kkt) {
            kkt->kkt_stat = stat / (1.0 + fn);
            kkt->kkt_prim = prim;
            kkt->kkt_comp = comp;
            kkt->kkt_dual = dual;
            kkt->objective = obj;
        }
    }
    free(L);
    free(J);
    free(R);
    free(dv);
    free(z);
    free(r);
    free(u);
    free(cn);
    free(act);
    free(isact);
    return rc;
}

/* ------------------------------------------------------------------------- */
/* a8-a10 one agent's solve                                                  */
/* ------------------------------------------------------------------------- */

typedef struct {
    const orc_params* P;
    orc_model M;
} orc_ctx;

static int solve_agent_ctx(const orc_ctx* X, const double* po, const double* pf, const double* vo,
                           const double* ao, int n, const double* l, int N, const double* pmin,
                           const double* pmax, double* p, double* v, double* a, orc_diag* diag) {
    const orc_params* P = X->P;
    const orc_model* M = &X->M;
    const int K = P->K, n3 = 3 * K;
    const int variant = P->variant;
    const int soft = (variant == ORC_VARIANT_SOFT_BOUND || variant == ORC_VARIANT_SOFT_BOUND2);
    const double* prev_p = l + (size_t)3 * K * n; /* prev_p = l(:,:,n) */
    orc_diag dg;
    memset(&dg, 0, sizeof(dg));
    int status = 0;

    uint8_t* violation = (uint8_t*)malloc((size_t)N);
    uint8_t* viol_constr = (uint8_t*)malloc((size_t)N);
    const int max_rows = (variant == ORC_VARIANT_HARD) ? K * (N > 1 ? N - 1 : 1) : (N > 1 ? N - 1 : 1);
    double* Ac = (double*)malloc(sizeof(double) * (size_t)max_rows * n3); /* rows x 3K, ld = max_rows */
    double* bc = (double*)malloc(sizeof(double) * max_rows);
    double* pdist = (double*)malloc(sizeof(double) * max_rows);
    int nrows = 0;
    int any_violation = 0;

    if (variant == ORC_VARIANT_HARD) {
        /* solveHardDMPC.m:18-22: rows for every horizon step, stacked in k order */
        for (int k = 1; k <= K; ++k) {
            nrows += coll_rows(P, M, prev_p + 3 * (k - 1), po, vo, n, k, k, l, N, NULL, Ac + nrows,
                               max_rows, bc + nrows, pdist + nrows, NULL, NULL);
        }
        dg.kstar = nrows ? 1 : 0;
        /* quirk: the reference's Ain_coll is never empty for N >= 2, because
         * CollConstrHardDMPC returns N-1 (possibly all-zero) rows -> collision weights always. */
        any_violation = (N >= 2);
    } else {
        /* solveSoftDMPCbound.m:21-38 / solveSoftDMPCbound2.m:18-36 / solveHardDMPCOnDemand.m:18-27 */
        for (int k = 1; k <= K; ++k) {
            double md;
            int viol = orc_check_coll(P, prev_p + 3 * (k - 1), l, N, n, k, violation, viol_constr, &md);
            if (!viol) continue;
            if (soft && k == 1 && md < P->rmin - P->coll_tol) {
                dg.kstar = 1;
                dg.min_dist = md;
                status = ORC_ST_COLL;
                goto out;
            }
            if (variant == ORC_VARIANT_SOFT_BOUND2 && k == 1) continue; /* solveSoftDMPCbound2.m:29-31 */
            int k_ctr = (variant == ORC_VARIANT_SOFT_BOUND2) ? k - 1 : k;
            nrows = coll_rows(P, M, prev_p + 3 * (k - 1), po, vo, n, k, k_ctr, l, N, viol_constr, Ac,
                              max_rows, bc, pdist, NULL, NULL);
            dg.kstar = k;
            dg.min_dist = md;
            any_violation = 1;
            break;
        }
    }
    dg.nv = nrows;

    /* weights: solveSoftDMPCbound.m:43-58 */
    double q, s;
    {
        double dgoal = sqrt((po[0] - pf[0]) * (po[0] - pf[0]) + (po[1] - pf[1]) * (po[1] - pf[1]) +
                            (po[2] - pf[2]) * (po[2] - pf[2]));
        if (!any_violation && dgoal >= P->near_radius) {
            q = P->Q_far;
            s = P->S_free;
        } else if (!any_violation) {
            q = P->Q_near;
            s = P->S_free;
        } else {
            q = P->Q1;
            s = P->S1;
        }
    }

    /* dense QP, x = [a (3K); eps (nsl)]  (solveSoftDMPCbound.m:60-98) */
    const int nsl = soft ? nrows : 0;
    const int nx = n3 + nsl;
    const int m = nrows + 2 * n3;
    double* H = (double*)calloc((size_t)nx * nx, sizeof(double));
    double* f = (double*)calloc((size_t)nx, sizeof(double));
    double* Ad = (double*)calloc((size_t)m * nx, sizeof(double));
    double* bd = (double*)calloc((size_t)m, sizeof(double));
    double* lb = (double*)malloc(sizeof(double) * nx);
    double* ub = (double*)malloc(sizeof(double) * nx);
    double* x = (double*)calloc((size_t)nx, sizeof(double));
    double* lamv = (double*)calloc((size_t)m + 2 * nx, sizeof(double));

    /* H = 2 (A'QA + Delta'S Delta + R + EPS): Q = q on the last block (spd = 1), R = I, S = s I */
    for (int i = 0; i < K; ++i)
        for (int j = 0; j < K; ++j) {
            double hij = q * M->lam[(K - 1) * K + i] * M->lam[(K - 1) * K + j];
            /* (Delta'Delta)_{ij}: 2 on the diagonal except the last (1), -1 on the off-diagonals */
            if (i == j) hij += s * ((i == K - 1) ? 1.0 : 2.0) + 1.0;
            if (i == j + 1 || j == i + 1) hij -= s;
            for (int d = 0; d < 3; ++d) H[(3 * i + d) + (size_t)nx * (3 * j + d)] = 2.0 * hij;
        }
    for (int j = 0; j < nsl; ++j) H[(n3 + j) + (size_t)nx * (n3 + j)] = 2.0;
    /* f = -2 ( (pf_rep - A_initp x0)' Q A + ao_1 S Delta ) + f_eps */
    double tK = M->tt[K - 1];
    for (int j = 0; j < K; ++j)
        for (int d = 0; d < 3; ++d) {
            double e = pf[d] - (po[d] + tK * vo[d]);
            double g = q * e * M->lam[(K - 1) * K + j];
            if (j == 0) g += s * ao[d];
            f[3 * j + d] = -2.0 * g;
        }
    double term = P->term;
    double slb = P->slack_lb;
    /* rows: [Ain_coll diag(prev_dist); A 0; -A 0]  (solveSoftDMPCbound.m:33,97) */
    for (int r = 0; r < nrows; ++r) {
        for (int c2 = 0; c2 < n3; ++c2) Ad[r + (size_t)m * c2] = Ac[r + (size_t)max_rows * c2];
        if (soft) Ad[r + (size_t)m * (n3 + r)] = pdist[r];
        bd[r] = bc[r];
    }
    for (int k = 0; k < K; ++k)
        for (int d = 0; d < 3; ++d) {
            int r1 = nrows + 3 * k + d, r2 = nrows + n3 + 3 * k + d;
            for (int j = 0; j <= k; ++j) {
                Ad[r1 + (size_t)m * (3 * j + d)] = M->lam[k * K + j];
                Ad[r2 + (size_t)m * (3 * j + d)] = -M->lam[k * K + j];
            }
            double p0 = po[d] + M->tt[k] * vo[d];
            bd[r1] = pmax[d] - p0;
            bd[r2] = -pmin[d] + p0;
        }
    for (int i = 0; i < n3; ++i) {
        ub[i] = P->alim;
        lb[i] = -P->alim;
    }

    int tries = 0, solved = 0;
    while (!solved && tries < P->max_tries) {
        for (int j = 0; j < nsl; ++j) {
            f[n3 + j] = term;
            lb[n3 + j] = slb;
            ub[n3 + j] = 0.0;
        }
        int32_t it = 0;
        int rc = orc_qp_gi(nx, H, f, m, Ad, m, bd, lb, ub, x, lamv, &it, &dg);
        dg.qp_iters = it;
        if (rc == 0) {
            solved = 1;
            break;
        }
        if (rc == 2) {
            status |= ORC_ST_QPFAIL;
            break;
        }
        /* infeasible.  Soft variants with slack: double the slack bound and the penalty and
         * retry (solveSoftDMPCbound.m:148-153).  Otherwise the reference only loosens quadprog's
         * ConstraintTolerance (no effect on an exact solver) or returns success = 0
         * (solveHardDMPC.m:82-88): report infeasible. */
        if (!(soft && nsl > 0)) break;
        slb *= 2.0;
        term *= 2.0;
        ++tries;
    }
    dg.tries = tries;
    if (solved) {
        for (int i = 0; i < nrows; ++i) dg.n_act_row += lamv[i] > 0.0;
        for (int i = nrows; i < m; ++i) dg.n_act_pos += lamv[i] > 0.0;
        for (int i = 0; i < n3; ++i) dg.n_act_box += (lamv[m + i] > 0.0) || (lamv[m + nx + i] > 0.0);
        for (int i = 0; i < nsl; ++i) dg.n_act_eps += x[n3 + i] < -1e-12;
        memcpy(a, x, sizeof(double) * n3);
        prop_state_m(M, po, vo, a, p, v);
        status |= ORC_ST_SOLVED;
        if (!orc_is_inbounds(p, pmin, pmax, P->inb_tol)) status |= ORC_ST_OUTBOUND;
    } else if (!(status & ORC_ST_QPFAIL)) {
        status |= ORC_ST_INFEASIBLE;
    }
    free(H);
    free(f);
    free(Ad);
    free(bd);
    free(lb);
    free(ub);
    free(x);
    free(lamv);
out:
    free(violation);
    free(viol_constr);
    free(Ac);
    free(bc);
    free(pdist);
    if (diag) *diag = dg;
    return status;
}

int orc_solve_agent(const orc_params* P, const double* po, const double* pf, const double* vo,
                    const double* ao, int n, const double* l, int N, const double* pmin,
                    const double* pmax, double* p, double* v, double* a, orc_diag* diag) {
    orc_ctx X;
    X.P = P;
    model_build(&X.M, P->h, P->K);
    int st = solve_agent_ctx(&X, po, pf, vo, ao, n, l, N, pmin, pmax, p, v, a, diag);
    model_free(&X.M);
    return st;
}

/* ------------------------------------------------------------------------- */
/* a14 Jacobi step                                                           */
/* ------------------------------------------------------------------------- */

typedef struct {
    const orc_ctx* X;
    int N, n0, n1;
    const double *pk, *vk, *ak, *pf, *l_prev, *pmin, *pmax;
    double *l_new, *p1, *v1, *a1;
    int32_t* status;
    orc_diag* diags;
} step_job;

/* failure_rate.m:100-119 / dmpc_soft_bound.m:116-135: body of `for n = 1:N` for k > 1.
 * Jacobi: every agent reads l_prev only. */
static void* step_worker(void* arg) {
    step_job* J = (step_job*)arg;
    const int K = J->X->P->K;
    double* p = (double*)malloc(sizeof(double) * 9 * K);
    double* v = p + 3 * K;
    double* a = v + 3 * K;
    for (int n = J->n0; n < J->n1; ++n) {
        orc_diag dg;
        int st = solve_agent_ctx(J->X, J->pk + 3 * n, J->pf + 3 * n, J->vk + 3 * n, J->ak + 3 * n, n,
                                 J->l_prev, J->N, J->pmin, J->pmax, p, v, a, &dg);
        J->status[n] = st | (dg.tries << 8);
        if (J->diags) J->diags[n] = dg;
        if (st & ORC_ST_SOLVED) {
            memcpy(J->l_new + (size_t)3 * K * n, p, sizeof(double) * 3 * K);
            for (int d = 0; d < 3; ++d) {
                J->p1[3 * n + d] = p[d];
                J->v1[3 * n + d] = v[d];
                J->a1[3 * n + d] = a[d];
            }
        }
    }
    free(p);
    return NULL;
}

int orc_step(const orc_params* P, int N, int n0, int n1, const double* pk, const double* vk,
             const double* ak, const double* pf, const double* l_prev, const double* pmin,
             const double* pmax, double* l_new, double* p1, double* v1, double* a1,
             int32_t* status, orc_diag* diags, int nthreads) {
    orc_ctx X;
    X.P = P;
    model_build(&X.M, P->h, P->K);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > n1 - n0) nthreads = (n1 - n0) > 0 ? (n1 - n0) : 1;
    step_job* jobs = (step_job*)malloc(sizeof(step_job) * nthreads);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
    /* contiguous clusters like dmpc.cpp:1600-1625 */
    int per = (n1 - n0 + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; ++t) {
        step_job j = {&X, N, n0 + t * per, n0 + (t + 1) * per, pk, vk, ak, pf, l_prev, pmin, pmax,
                      l_new, p1, v1, a1, status, diags};
        if (j.n1 > n1) j.n1 = n1;
        if (j.n0 > n1) j.n0 = n1;
        jobs[t] = j;
    }
    if (nthreads == 1) {
        step_worker(&jobs[0]);
    } else {
        for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, step_worker, &jobs[t]);
        for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    }
    int first_fail = -1;
    for (int n = n0; n < n1; ++n) {
        int st = status[n] & 0xff;
        if (!(st & ORC_ST_SOLVED) || (st & ORC_ST_OUTBOUND)) {
            first_fail = n;
            break;
        }
    }
    free(jobs);
    free(th);
    model_free(&X.M);
    return first_fail;
}

/* ------------------------------------------------------------------------- */
/* scenario generation: randomTest.m:1-57 / randomExchange.m:1-56             */
/* (C++: gen_rand_pts dmpc.cpp:188-227, gen_rand_perm :229-265)               */
/* ------------------------------------------------------------------------- */
/* The reference draws from MATLAB's unseeded `rand`; its streams cannot be reproduced.  The random stream here
 * is the counter-based one the library documents (include/dmpc_b200.h, dmpcb200_gen_scenarios): splitmix64 of
 * (seed, scenario, set, draw index).  Everything else follows the reference line by line. */
static uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static uint64_t gen_key(uint64_t seed, int scen, int set) {
    return splitmix64(seed + 0x632BE59BD9B4E019ull * (uint64_t)(2 * scen + set + 1));
}
static double gen_u01(uint64_t key, uint64_t idx) {
    return (double)(splitmix64(key + idx) >> 11) * (1.0 / 9007199254740992.0);
}

/* one set of N points: randomTest.m:7-29 (mode 0: ellipsoidal distance) / randomExchange.m:7-29 (mode 1) */
static void gen_set(uint64_t key, int N, const double* pmin, const double* pmax, double rmin, double inv_c,
                    int max_iter, double* pts /*3 x N*/) {
    uint64_t t = 0;
    int n = 0, tries = 0;
    while (n < N) {
        double c[3];
        for (int x = 0; x < 3; ++x) c[x] = pmin[x] + (pmax[x] - pmin[x]) * gen_u01(key, 3 * t + (uint64_t)x);
        ++t;
        int ok = 1;
        for (int j = 0; j < n && ok; ++j) {
            double dx = pts[3 * j] - c[0], dy = pts[3 * j + 1] - c[1], ez = inv_c * (pts[3 * j + 2] - c[2]);
            double dist = sqrt(dx * dx + dy * dy + ez * ez);
            ok = dist > rmin;
        }
        if (ok) {
            pts[3 * n] = c[0]; pts[3 * n + 1] = c[1]; pts[3 * n + 2] = c[2];
            ++n;
            tries = 0;
        } else if (++tries > max_iter) { /* :23-25: start the set again */
            n = 0;
            tries = 0;
        }
    }
}

void orc_gen_scenario(uint64_t seed, int scen, int mode, int N, const double* pmin, const double* pmax,
                      double rmin, double c, double* po, double* pf) {
    const int max_iter = 200000;
    gen_set(gen_key(seed, scen, 0), N, pmin, pmax, rmin, mode == 0 ? 1.0 / c : 1.0, max_iter, po);
    if (mode == 0) {
        gen_set(gen_key(seed, scen, 1), N, pmin, pmax, rmin, 1.0 / c, max_iter, pf);
        return;
    }
    /* randomExchange.m:32-49 */
    int* array = (int*)malloc(sizeof(int) * (size_t)N);
    int* aux = (int*)malloc(sizeof(int) * (size_t)N);
    int* perm = (int*)malloc(sizeof(int) * (size_t)N);
    int left = N;
    uint64_t key = gen_key(seed, scen, 1);
    for (int i = 0; i < N; ++i) array[i] = i;
    for (int i = 0; i < N; ++i) {
        int na = 0;
        for (int e = 0; e < left; ++e)
            if (array[e] != i) aux[na++] = array[e]; /* array_aux(array_aux == i) = [] */
        int pick;
        if (i == N - 1) pick = array[0];
        else if (i == N - 2 && aux[na - 1] == N - 1) pick = N - 1;
        else pick = aux[(int)(gen_u01(key, (uint64_t)i) * (double)(N - i - 1))]; /* randi([1 N-i]) */
        perm[i] = pick;
        int w = 0;
        for (int e = 0; e < left; ++e)
            if (array[e] != pick) array[w++] = array[e];
        left = w;
    }
    for (int i = 0; i < N; ++i)
        for (int x = 0; x < 3; ++x) pf[3 * i + x] = po[3 * perm[i] + x];
    free(array);
    free(aux);
    free(perm);
}

// qp_core.cuh -- warp-cooperative exact solver for one agent's DMPC QP (device code).
//
// Problem (solveSoftDMPCbound.m:60-103 of the reference), variables
// a in R^{3K} (index i = 3k + x), one slack eps_j per collision row (soft variants):
//
//   min  sum_x [ 1/2 a_x' H_K a_x + f_x' a_x ]  +  sum_j ( eps_j^2 + term eps_j )
//   s.t. -alim <= a_i <= alim                                  (BOXL / BOXU)
//        pmin_x <= P_i <= pmax_x,  P = p0 + Lam a               (WSL / WSU, every horizon step)
//        d_j . P[kc_j] - dist_j eps_j >= rhs_j                  (ROW j)
//        lb <= eps_j <= 0                                       (SLB j / SUB j)
//
// The reference hands a dense (3K+nv)-variable QP to MATLAB quadprog.  Here the same optimum
// (the QP is strictly convex, so it is unique) is computed by a Goldfarb-Idnani dual active-set
// method written in Schur-complement form: with N the matrix of active constraint normals the
// method needs S = N' H^{-1} N only, and because H = kron(H_K, I_3) (+ 2I on the slacks) is one of
// three constant matrices, every entry of S and of H^{-1} n is a LOOKUP in the K x K tables
// G = H_K^{-1}, B = G Lam', C = Lam G Lam' (model_tables.h).  Neither H, nor a factor of it, nor a
// dense constraint row is ever formed.  We carry M = S^{-1} explicitly (q x q, shared memory):
//   add constraint    -> bordering update of M      (rank 1, no dependent chain)
//   drop constraint   -> Schur downdate + swap-remove
// so one iteration is a handful of warp-wide q x q sweeps; lanes own rows of M.  Slack upper
// bounds (eps_j <= 0) all start active (the slack's free optimum is -term/2 >> 0); they are kept
// IMPLICIT (not stored in M: they are decoupled from everything until row j is touched) and are
// materialised lazily, so q stays ~ (#active box) + 2 (#touched rows).
// A final "polish" re-synthesises x from the multipliers (x = x_unc + H^{-1} N u) and refines u
// with M as approximate inverse, removing any drift the explicit-inverse updates accumulated.
//
// One warp per agent; all branches are warp-uniform (decided on reduced values).  The same source
// compiles for the host with one "lane" (a test-only build): that build is a debugging aid of the
// test-suite only and is never part of the product library.
#pragma once

#include <math.h>
#include <stdint.h>
#ifdef DMPC_DEBUG
#include <stdio.h>
#endif

#if defined(__CUDACC__)
#define DMPC_HD __host__ __device__ __forceinline__
#define DMPC_D __device__ __forceinline__
#else
#define DMPC_HD inline
#define DMPC_D inline
#endif

namespace dmpc {

// ---- lane helpers -------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
DMPC_D int lane_id() { return (int)(threadIdx.x & 31u); }
constexpr int kLanes = 32;
DMPC_D void wsync() { __syncwarp(); }
DMPC_D double wsum(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
DMPC_D double wmax(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// (value, index) arg-min; ties -> smaller index; result identical in every lane
DMPC_D void wargmin(double& v, int& idx) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, v, o);
        int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov < v || (ov == v && oi < idx)) {
            v = ov;
            idx = oi;
        }
    }
}
DMPC_D unsigned wballot(bool p) { return __ballot_sync(0xffffffffu, p); }
DMPC_D int popc_below(unsigned m) { return __popc(m & ((1u << (threadIdx.x & 31u)) - 1u)); }
DMPC_D int popc_all(unsigned m) { return __popc(m); }
#else
inline int lane_id() { return 0; }
constexpr int kLanes = 1;
inline void wsync() {}
inline double wsum(double v) { return v; }
inline double wmax(double v) { return v; }
inline void wargmin(double&, int&) {}
inline unsigned wballot(bool p) { return p ? 1u : 0u; }
inline int popc_below(unsigned) { return 0; }
inline int popc_all(unsigned m) { return (int)(m & 1u); }
#endif

// constraint codes: type << 16 | index
enum { T_BOXL = 0, T_BOXU = 1, T_WSL = 2, T_WSU = 3, T_ROW = 4, T_SUB = 5, T_SLB = 6 };
DMPC_HD int mk_code(int type, int idx) { return (type << 16) | idx; }
DMPC_HD int code_type(int c) { return c >> 16; }
DMPC_HD int code_idx(int c) { return c & 0xffff; }

enum { QP_OK = 0, QP_INFEASIBLE = 1, QP_ITERCAP = 2, QP_OVERFLOW = 3 };

struct QpWs {
    int K, n3, QMAX;
    // constant tables of the agent's weight set (shared memory), row-major K x K
    const double *lam, *lnorm, *G, *B, *C;
    // problem data
    double alim, term, slb;
    double qw, sw;  // terminal-position and input-variation weights of the set in use
    double pmin[3], pmax[3];
    int soft;  // rows carry a slack variable
    const double* p0;    // n3: po + tt[k] vo
    const double* aunc;  // n3: unconstrained optimum of a
    // collision rows (SoA)
    int nv;
    const double *rd0, *rd1, *rd2, *rdist, *rrhs;
    const int* rkc;  // 0-based horizon index the row acts on
    double* rnorm;   // norm of the dense row (constraint selection is on normalised violation)
    // mutable per-agent state
    double *a, *P, *cbox, *cP, *z;
    double *eps, *zeps;
    int *mbox, *mws;                      // per i: bit0 lower active, bit1 upper active
    int *rslot, *ubslot, *lbslot, *rmat;  // per row: slot in the active list or -1; materialised?
    int* act;                             // active list: constraint codes
    double *u, *g, *r, *M;                // multipliers, scratch, M = (N'H^-1 N)^-1 (ld = QMAX)
    int* ralist;                          // scratch: slots of active rows
};

struct QpResult {
    int rc;
    int iters;
    int q;
};

// decoded candidate constraint p (uniform across the warp)
struct PInfo {
    int type, idx, k, x, j, kc;
    double sig, d0, d1, d2, dist, nph;
};

struct Qp {
    QpWs w;
    int q, nmat;

    // ---- H^{-1} n_p lookups ---------------------------------------------------------------
    // a-space entry i=(k,x) of H^{-1} n_p
    DMPC_D double hA(const PInfo& p, int k, int x) const {
        const int K = w.K;
        switch (p.type) {
            case T_BOXL:
            case T_BOXU: return (x == p.x) ? p.sig * w.G[k * K + p.k] : 0.0;
            case T_WSL:
            case T_WSU: return (x == p.x) ? p.sig * w.B[k * K + p.k] : 0.0;
            case T_ROW: {
                const double dx = (x == 0) ? p.d0 : ((x == 1) ? p.d1 : p.d2);
                return dx * w.B[k * K + p.kc];
            }
            default: return 0.0;
        }
    }
    // P-space image (Lam H^{-1} n_p) entry (k,x)
    DMPC_D double hP(const PInfo& p, int k, int x) const {
        const int K = w.K;
        switch (p.type) {
            case T_BOXL:
            case T_BOXU: return (x == p.x) ? p.sig * w.B[p.k * K + k] : 0.0;
            case T_WSL:
            case T_WSU: return (x == p.x) ? p.sig * w.C[k * K + p.k] : 0.0;
            case T_ROW: {
                const double dx = (x == 0) ? p.d0 : ((x == 1) ? p.d1 : p.d2);
                return dx * w.C[k * K + p.kc];
            }
            default: return 0.0;
        }
    }
    // slack-space entry j of H^{-1} n_p  (slack Hessian is 2)
    DMPC_D double hE(const PInfo& p, int j) const {
        if (j != p.j) return 0.0;
        switch (p.type) {
            case T_ROW: return w.soft ? -0.5 * p.dist : 0.0;
            case T_SUB: return -0.5;
            case T_SLB: return 0.5;
            default: return 0.0;
        }
    }

    DMPC_D PInfo decode(int code) const {
        PInfo p;
        p.type = code_type(code);
        p.idx = code_idx(code);
        p.k = p.x = 0;
        p.j = -1;
        p.kc = 0;
        p.sig = 1.0;
        p.d0 = p.d1 = p.d2 = p.dist = 0.0;
        const int K = w.K;
        if (p.type <= T_WSU) {
            p.k = p.idx / 3;
            p.x = p.idx - 3 * p.k;
            p.sig = (p.type == T_BOXL || p.type == T_WSL) ? 1.0 : -1.0;
            p.nph = (p.type <= T_BOXU) ? w.G[p.k * K + p.k] : w.C[p.k * K + p.k];
        } else {
            p.j = p.idx;
            if (p.type == T_ROW) {
                p.d0 = w.rd0[p.j];
                p.d1 = w.rd1[p.j];
                p.d2 = w.rd2[p.j];
                p.dist = w.rdist[p.j];
                p.kc = w.rkc[p.j];
                p.nph = (p.d0 * p.d0 + p.d1 * p.d1 + p.d2 * p.d2) * w.C[p.kc * K + p.kc] +
                        (w.soft ? 0.5 * p.dist * p.dist : 0.0);
            } else {
                p.nph = 0.5;
            }
        }
        return p;
    }

    // residual c'x - b of a constraint at the current x (>= 0 feasible)
    DMPC_D double resid(int code) const {
        const int t = code_type(code), i = code_idx(code);
        switch (t) {
            case T_BOXL: return w.a[i] + w.alim;
            case T_BOXU: return w.alim - w.a[i];
            case T_WSL: return w.P[i] - w.pmin[i % 3];
            case T_WSU: return w.pmax[i % 3] - w.P[i];
            case T_ROW: {
                const int kc = w.rkc[i];
                double s = w.rd0[i] * w.P[3 * kc] + w.rd1[i] * w.P[3 * kc + 1] +
                           w.rd2[i] * w.P[3 * kc + 2] - w.rrhs[i];
                if (w.soft) s -= w.rdist[i] * w.eps[i];
                return s;
            }
            case T_SUB: return -w.eps[i];
            default: return w.eps[i] - w.slb;
        }
    }

    // P = p0 + Lam a
    DMPC_D void update_P() {
        const int K = w.K;
        for (int i = lane_id(); i < w.n3; i += kLanes) {
            const int k = i / 3, x = i - 3 * k;
            double s = 0.0;
            for (int j = 0; j <= k; ++j) s = fma(w.lam[k * K + j], w.a[3 * j + x], s);
            w.P[i] = s + w.p0[i];
        }
        wsync();
    }

    // cbox / cP <- sum over the active set of coef_s * (normal of slot s), split by basis:
    //   cbox: coefficients on unit vectors e_i (box constraints)
    //   cP  : coefficients on rows of Lam (workspace constraints and collision rows)
    DMPC_D void accumulate(const double* coef) {
        for (int i = lane_id(); i < w.n3; i += kLanes) {
            w.cbox[i] = 0.0;
            w.cP[i] = 0.0;
        }
        wsync();
        // compact list of active row slots (for the gather below)
        int nra = 0;
        for (int base = 0; base < q; base += kLanes) {
            const int s = base + lane_id();
            bool isrow = false;
            if (s < q) {
                const int c = w.act[s];
                const int t = code_type(c), i = code_idx(c);
                if (t == T_BOXL) w.cbox[i] = coef[s];
                else if (t == T_BOXU) w.cbox[i] = -coef[s];
                else if (t == T_WSL) w.cP[i] = coef[s];
                else if (t == T_WSU) w.cP[i] = -coef[s];
                else if (t == T_ROW) isrow = true;
            }
            const unsigned m = wballot(isrow);
            if (isrow) w.ralist[nra + popc_below(m)] = s;
            nra += popc_all(m);
        }
        wsync();
        if (nra) {
            for (int i = lane_id(); i < w.n3; i += kLanes) {
                const int k = i / 3, x = i - 3 * k;
                double acc = 0.0;
                for (int e = 0; e < nra; ++e) {
                    const int s = w.ralist[e];
                    const int j = code_idx(w.act[s]);
                    if (w.rkc[j] == k) {
                        const double dx = (x == 0) ? w.rd0[j] : ((x == 1) ? w.rd1[j] : w.rd2[j]);
                        acc = fma(coef[s], dx, acc);
                    }
                }
                w.cP[i] += acc;
            }
            wsync();
        }
    }

    // out_i = base_i + sgn * (G cbox + B cP)_i     (a-space image of the accumulated normals)
    DMPC_D void apply_Hinv(double* out, const double* base, double sgn, const PInfo* hp) {
        const int K = w.K;
        for (int i = lane_id(); i < w.n3; i += kLanes) {
            const int k = i / 3, x = i - 3 * k;
            double s = 0.0;
            for (int j = 0; j < K; ++j) {
                s = fma(w.G[k * K + j], w.cbox[3 * j + x], s);
                s = fma(w.B[k * K + j], w.cP[3 * j + x], s);
            }
            const double b = hp ? hA(*hp, k, x) : base[i];
            out[i] = b + sgn * s;
        }
        wsync();
    }

    DMPC_D void set_active(int code, int slot) {
        const int t = code_type(code), i = code_idx(code);
        switch (t) {
            case T_BOXL: w.mbox[i] |= 1; break;
            case T_BOXU: w.mbox[i] |= 2; break;
            case T_WSL: w.mws[i] |= 1; break;
            case T_WSU: w.mws[i] |= 2; break;
            case T_ROW: w.rslot[i] = slot; break;
            case T_SUB: w.ubslot[i] = slot; break;
            default: w.lbslot[i] = slot; break;
        }
    }
    DMPC_D void clear_active(int code) {
        const int t = code_type(code), i = code_idx(code);
        switch (t) {
            case T_BOXL: w.mbox[i] &= ~1; break;
            case T_BOXU: w.mbox[i] &= ~2; break;
            case T_WSL: w.mws[i] &= ~1; break;
            case T_WSU: w.mws[i] &= ~2; break;
            case T_ROW: w.rslot[i] = -1; break;
            case T_SUB: w.ubslot[i] = -1; break;
            default: w.lbslot[i] = -1; break;
        }
    }

    // most violated (normalised) inactive constraint; returns code or -1
    DMPC_D int most_violated(double tol, double* sp_out) {
        double best = -tol;
        int bcode = 0x7fffffff;
        for (int i = lane_id(); i < w.n3; i += kLanes) {
            const int x = i % 3, k = i / 3;
            const int mb = w.mbox[i], mw = w.mws[i];
            const double ai = w.a[i], Pi = w.P[i];
            const double iln = 1.0 / w.lnorm[k];
            double s;
            if (!(mb & 1)) {
                s = ai + w.alim;
                if (s < best) { best = s; bcode = mk_code(T_BOXL, i); }
            }
            if (!(mb & 2)) {
                s = w.alim - ai;
                if (s < best) { best = s; bcode = mk_code(T_BOXU, i); }
            }
            if (!(mw & 1)) {
                s = Pi - w.pmin[x];
                if (s * iln < best) { best = s * iln; bcode = mk_code(T_WSL, i); }
            }
            if (!(mw & 2)) {
                s = w.pmax[x] - Pi;
                if (s * iln < best) { best = s * iln; bcode = mk_code(T_WSU, i); }
            }
        }
        for (int j = lane_id(); j < w.nv; j += kLanes) {
            if (w.rslot[j] < 0) {
                const double s = resid(mk_code(T_ROW, j));
                const double sn = s / w.rnorm[j];
                if (sn < best) { best = sn; bcode = mk_code(T_ROW, j); }
            }
            if (w.soft && w.rmat[j]) {
                if (w.ubslot[j] < 0) {
                    const double s = -w.eps[j];
                    if (s < best) { best = s; bcode = mk_code(T_SUB, j); }
                }
                if (w.lbslot[j] < 0) {
                    const double s = w.eps[j] - w.slb;
                    if (s < best) { best = s; bcode = mk_code(T_SLB, j); }
                }
            }
        }
        wargmin(best, bcode);
        if (bcode == 0x7fffffff) return -1;
        *sp_out = resid(bcode);  // recomputed uniformly by every lane
        return bcode;
    }

    // append slot with explicit M row/col = (zeros, diag)
    DMPC_D void append_isolated(int code, double uval, double mdiag) {
        const int Q = w.QMAX;
        for (int i = lane_id(); i < q; i += kLanes) {
            w.M[i + Q * q] = 0.0;
            w.M[q + Q * i] = 0.0;
        }
        if (lane_id() == 0) {
            w.M[q + Q * q] = mdiag;
            w.act[q] = code;
            w.u[q] = uval;
            set_active(code, q);
        }
        wsync();
        ++q;
    }

    // drop slot l: Schur downdate of M, then move the last slot into l
    DMPC_D void drop_slot(int l) {
        const int Q = w.QMAX;
        const double inv = 1.0 / w.M[l + Q * l];
        for (int i = lane_id(); i < q; i += kLanes) {
            if (i == l) continue;
            const double ci = w.M[i + Q * l] * inv;
            for (int j = 0; j < q; ++j) {
                if (j == l) continue;
                w.M[i + Q * j] = fma(-ci, w.M[j + Q * l], w.M[i + Q * j]);
            }
        }
        wsync();
        const int last = q - 1;
        if (lane_id() == 0) clear_active(w.act[l]);
        wsync();
        if (l != last) {
            for (int i = lane_id(); i < last; i += kLanes) {
                if (i == l) continue;
                const double v = w.M[i + Q * last];
                w.M[i + Q * l] = v;
                w.M[l + Q * i] = v;
            }
            if (lane_id() == 0) {
                w.M[l + Q * l] = w.M[last + Q * last];
                const int c = w.act[last];
                w.act[l] = c;
                w.u[l] = w.u[last];
                set_active(c, l);
            }
        }
        wsync();
        --q;
    }

    // x (a, eps, P) re-synthesised from the multipliers:  x = x_unc + H^{-1} N u
    DMPC_D void synth_from_u() {
        accumulate(w.u);
        apply_Hinv(w.a, w.aunc, 1.0, nullptr);
        if (w.soft) {
            for (int j = lane_id(); j < w.nv; j += kLanes) {
                if (!w.rmat[j]) continue;  // implicit upper bound: eps = 0
                double e = -0.5 * w.term;
                const int sr = w.rslot[j], su = w.ubslot[j], sl = w.lbslot[j];
                if (sr >= 0) e -= 0.5 * w.rdist[j] * w.u[sr];
                if (su >= 0) e -= 0.5 * w.u[su];
                if (sl >= 0) e += 0.5 * w.u[sl];
                w.eps[j] = e;
            }
        }
        wsync();
        update_P();
    }

    // g[s] = n_{act[s]}' H^{-1} n_p for s < cnt   (an entry of the Schur complement S: all lookups)
    DMPC_D void gvec(const PInfo& p, int cnt) {
        for (int s = lane_id(); s < cnt; s += kLanes) {
            const int c = w.act[s];
            const int t = code_type(c), i = code_idx(c);
            double gv;
            if (t <= T_BOXU) gv = ((t == T_BOXL) ? 1.0 : -1.0) * hA(p, i / 3, i % 3);
            else if (t <= T_WSU) gv = ((t == T_WSL) ? 1.0 : -1.0) * hP(p, i / 3, i % 3);
            else if (t == T_ROW) {
                const int kc = w.rkc[i];
                gv = w.rd0[i] * hP(p, kc, 0) + w.rd1[i] * hP(p, kc, 1) + w.rd2[i] * hP(p, kc, 2);
                if (w.soft) gv -= w.rdist[i] * hE(p, i);
            } else if (t == T_SUB) gv = -hE(p, i);
            else gv = hE(p, i);
            w.g[s] = gv;
        }
        wsync();
    }

    // r = M[0:cnt,0:cnt] g ; returns g'r and max|r| (uniform)
    DMPC_D void mat_vec(int cnt, double* gr_out, double* rmax_out) {
        const int Q = w.QMAX;
        double gr = 0.0, rmax = 0.0;
        for (int i = lane_id(); i < cnt; i += kLanes) {
            double s0 = 0.0, s1 = 0.0;
            int j = 0;
            for (; j + 1 < cnt; j += 2) {
                s0 = fma(w.M[i + Q * j], w.g[j], s0);
                s1 = fma(w.M[i + Q * (j + 1)], w.g[j + 1], s1);
            }
            if (j < cnt) s0 = fma(w.M[i + Q * j], w.g[j], s0);
            const double ri = s0 + s1;
            w.r[i] = ri;
            gr = fma(w.g[i], ri, gr);
            rmax = fmax(rmax, fabs(ri));
        }
        *gr_out = wsum(gr);
        *rmax_out = wmax(rmax);
        wsync();
    }

    // bordering update: M <- inverse of [[S, g],[g', nph]] given r = M g and delta = nph - g'r
    DMPC_D void border(int cnt, double delta) {
        const int Q = w.QMAX;
        const double id = 1.0 / delta;
        for (int i = lane_id(); i < cnt; i += kLanes) {
            const double ci = w.r[i] * id;
            for (int j = 0; j < cnt; ++j) w.M[i + Q * j] = fma(ci, w.r[j], w.M[i + Q * j]);
            w.M[i + Q * cnt] = -ci;
            w.M[cnt + Q * i] = -ci;
        }
        if (lane_id() == 0) w.M[cnt + Q * cnt] = id;
        wsync();
    }

    // rebuild M exactly from the active list (S entries are lookups): removes accumulated drift
    DMPC_D void refresh() {
        for (int s = 0; s < q; ++s) {
            const PInfo p = decode(w.act[s]);
            gvec(p, s);
            double gr, rmax;
            mat_vec(s, &gr, &rmax);
            double delta = p.nph - gr;
            if (!(delta > 1e-14 * p.nph)) delta = 1e-14 * p.nph;
            border(s, delta);
        }
    }

    // z'Hz for the a-part of z: H_K = 2 (q lamK lamK' + s Delta'Delta + I) -- a sum of squares
    DMPC_D double zHz() const {
        const int K = w.K;
        const double* lamK = w.lam + (K - 1) * K;
        double loc = 0.0, t0 = 0.0, t1 = 0.0, t2 = 0.0;
        for (int i = lane_id(); i < w.n3; i += kLanes) {
            const int k = i / 3, x = i - 3 * k;
            const double zi = w.z[i];
            const double dz = zi - ((k > 0) ? w.z[i - 3] : 0.0);
            loc = fma(zi, zi, loc);
            loc = fma(w.sw * dz, dz, loc);
            const double lz = lamK[k] * zi;
            if (x == 0) t0 += lz;
            else if (x == 1) t1 += lz;
            else t2 += lz;
        }
        loc = wsum(loc);
        t0 = wsum(t0);
        t1 = wsum(t1);
        t2 = wsum(t2);
        return 2.0 * (loc + w.qw * (t0 * t0 + t1 * t1 + t2 * t2));
    }

    // returns max residual of the active constraints after refinement
    DMPC_D double polish() {
        double mx = 0.0;
        for (int round = 0; round < 4; ++round) {
            synth_from_u();
            mx = 0.0;
            for (int s = lane_id(); s < q; s += kLanes) {
                const double rs = resid(w.act[s]);
                w.g[s] = rs;
                mx = fmax(mx, fabs(rs));
            }
            mx = wmax(mx);
            wsync();
#ifdef DMPC_DEBUG
            printf("  polish round %d q %d max resid %.3e\n", round, q, mx);
#endif
            if (!(mx > 1e-13)) break;
            double gr, rmax;
            mat_vec(q, &gr, &rmax);
            for (int i = lane_id(); i < q; i += kLanes) w.u[i] -= w.r[i];
            wsync();
        }
        return mx;
    }

    // ---- the solver ---------------------------------------------------------------------------
    // expects: w.a = aunc, eps = 0, masks cleared, rslot/ubslot/lbslot = -1, rmat = 0
    DMPC_D QpResult solve(int max_iter) {
        const int Q = w.QMAX;
        const double feas_tol = 1e-10;
        const double dep_tol = 1e-9;    // on delta = z'Hz relative to n_p'H^{-1}n_p
        const double ill_tol = 1e-5;    // adds below this mark M for an exact rebuild
        q = 0;
        nmat = 0;
        int iters = 0, npolish = 0;
        bool polished = false, dirty = false;
        QpResult res;
        update_P();
        for (;;) {
            double sp;
            const int pcode = most_violated(feas_tol, &sp);
            if (pcode < 0) {
                if (!polished && q > 0) {
                    if (dirty) { refresh(); dirty = false; }
                    double mx = polish();
                    if (mx > 1e-9) {
                        refresh();
                        mx = polish();
                    }
                    polished = true;
                    if (++npolish > 8) { res.rc = QP_ITERCAP; break; }
                    continue;
                }
                res.rc = QP_OK;
                break;
            }
            polished = false;
            PInfo p = decode(pcode);
            if (p.type == T_ROW && w.soft && !w.rmat[p.j]) {
                // materialise the (active) slack upper bound of this row: isolated so far,
                // S entry 1/2 -> M entry 2, multiplier -term - 2 eps = -term
                if (q + 2 > Q) { res.rc = QP_OVERFLOW; break; }
                if (lane_id() == 0) w.rmat[p.j] = 1;
                append_isolated(mk_code(T_SUB, p.j), -w.term - 2.0 * w.eps[p.j], 2.0);
                ++nmat;
            }
            if (q + 1 > Q) { res.rc = QP_OVERFLOW; break; }
            double up = 0.0;
            bool failed = false, added = false;
            while (!added) {
                if (++iters > max_iter) { res.rc = QP_ITERCAP; failed = true; break; }
                gvec(p, q);
                double gr, rmax;
                mat_vec(q, &gr, &rmax);
                // primal direction z = H^{-1}(n_p - N r); delta = n_p'z = z'Hz (no cancellation)
                accumulate(w.r);
                apply_Hinv(w.z, nullptr, -1.0, &p);
                double de = 0.0;
                if (w.soft) {
                    for (int j = lane_id(); j < w.nv; j += kLanes) {
                        if (!w.rmat[j]) continue;
                        double ze = hE(p, j);
                        const int sr = w.rslot[j], su = w.ubslot[j], sl = w.lbslot[j];
                        if (sr >= 0) ze += 0.5 * w.rdist[j] * w.r[sr];
                        if (su >= 0) ze += 0.5 * w.r[su];
                        if (sl >= 0) ze -= 0.5 * w.r[sl];
                        w.zeps[j] = ze;
                        de = fma(ze, ze, de);
                    }
                    de = 2.0 * wsum(de);
                }
                const double delta = zHz() + de;
#ifdef DMPC_DEBUG
                printf("it %d p type %d idx %d sp %.3e nph %.3e delta %.3e (schur %.3e) q %d\n", iters, p.type, p.idx, sp, p.nph, delta, p.nph - gr, q);
#endif
                const bool dependent = !(delta > dep_tol * p.nph) || (q >= w.n3 + nmat);
                // dual ratio test
                double t1 = INFINITY;
                int ldrop = 0x7fffffff;
                const double rthr = 1e-12 * rmax;
                for (int i = lane_id(); i < q; i += kLanes) {
                    const double ri = w.r[i];
                    if (ri > rthr) {
                        const double t = fmax(w.u[i], 0.0) / ri;
                        if (t < t1) { t1 = t; ldrop = i; }
                    }
                }
                wargmin(t1, ldrop);
                const double t2 = dependent ? INFINITY : (-sp / delta);
                const double t = (t1 < t2) ? t1 : t2;
                if (!(t < INFINITY)) { res.rc = QP_INFEASIBLE; failed = true; break; }
                if (!dependent) {
                    for (int i = lane_id(); i < w.n3; i += kLanes) w.a[i] = fma(t, w.z[i], w.a[i]);
                    if (w.soft)
                        for (int j = lane_id(); j < w.nv; j += kLanes)
                            if (w.rmat[j]) w.eps[j] = fma(t, w.zeps[j], w.eps[j]);
                    wsync();
                    update_P();
                }
                for (int i = lane_id(); i < q; i += kLanes) w.u[i] = fma(-t, w.r[i], w.u[i]);
                up += t;
                wsync();
                if (!dependent && t2 <= t1) {
                    // full step: constraint p becomes active
                    border(q, delta);
                    if (lane_id() == 0) {
                        w.act[q] = pcode;
                        w.u[q] = up;
                        set_active(pcode, q);
                    }
                    wsync();
                    ++q;
                    added = true;
                    if (delta < ill_tol * p.nph) dirty = true;
                } else {
                    drop_slot(ldrop);
                    if (dirty) { refresh(); dirty = false; }
                    if (!dependent) sp = resid(pcode);
                }
            }
            if (failed) break;
        }
        res.iters = iters;
        res.q = q;
        return res;
    }
};

}  // namespace dmpc

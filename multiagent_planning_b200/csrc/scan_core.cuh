// scan_core.cuh -- per-agent neighbour scan and collision-row construction (warp-cooperative).
//
// Reference: the horizon scan of solveSoftDMPCbound.m:21-38 (solveSoftDMPCbound2.m:18-36,
// solveHardDMPCOnDemand.m:18-27, solveHardDMPC.m:18-22) with CheckCollSoftDMPC.m:1-17 and
// CollConstrSoftDMPC.m:1-32 / CollConstrSoftDMPC2.m / CollConstrHardDMPC.m /
// CollConstrHardDMPCOnDemand.m.  The reference calls CheckColl once per horizon step (an O(N)
// interpreted loop each) and then re-walks all agents to build dense 3K-wide rows.  Here one pass
// over the neighbour horizons computes, for every neighbour, a K-bit "violates at step k" mask and
// a K-bit "near at step k" mask; the first violating step is a warp OR-reduction, and the rows
// are emitted from the near masks as the rank-1 data (d_j, dist_j, rhs_j, kc_j) only:
//     dense reference row  =  -(Lam[kc,:] (x) d_j'),   slack column dist_j,   b = -r_j
//     here                    d_j . P[kc] - dist_j eps_j >= rhs_j ,  rhs_j = r_j + d_j . p0[kc]
//                                                              = dist_j (rmin - dist_j) + d_j . p
// Distances use separately rounded multiply/add (no FMA) and IEEE sqrt/div so that the discrete
// decisions (dist < rmin, first violating k, neighbour set) are bit-identical to a plain C
// evaluation of the reference's formula.
#pragma once
#include <string.h>

#include "agent_solve.cuh"

namespace dmpc {

#if defined(__CUDA_ARCH__)
DMPC_D double mul_rn(double a, double b) { return __dmul_rn(a, b); }
DMPC_D double add_rn(double a, double b) { return __dadd_rn(a, b); }
DMPC_D unsigned wor(unsigned v) { return __reduce_or_sync(0xffffffffu, v); }
DMPC_D double wmin(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
#else
// host emulation: compiled with -ffp-contract=off
inline double mul_rn(double a, double b) { volatile double r = a * b; return r; }
inline double add_rn(double a, double b) { volatile double r = a + b; return r; }
inline unsigned wor(unsigned v) { return v; }
inline double wmin(double v) { return v; }
#endif

// norm(E1*(p-pj),2), E1 = diag(1,1,1/c)   (CheckCollSoftDMPC.m:10)
DMPC_D double ell_dist(double dx, double dy, double dz, double c) {
    const double ez = dz / c;
    return sqrt(add_rn(add_rn(mul_rn(dx, dx), mul_rn(dy, dy)), mul_rn(ez, ez)));
}
// the rounded sum of squares under the square root of ell_dist (same rounding sequence)
DMPC_D double ell_sq(double dx, double dy, double dz, double c) {
    const double ez = dz / c;
    return add_rn(add_rn(mul_rn(dx, dx), mul_rn(dy, dy)), mul_rn(ez, ez));
}

// neighbour threshold at 1-based step k (CheckCollSoftDMPC.m:12 ; C++ dmpc.cpp:418)
DMPC_HD double neigh_thr(const DevParams& P, int k1) {
    if (P.variant == VAR_HARD) return P.hard_radius;  // CollConstrHardDMPC.m:19
    if (P.neigh_mode == 1) return P.rmin * (1.0 + (double)(k1 - 1) / P.K);
    return P.rmin * P.neigh_factor;
}

// Exact squared thresholds.  sqrt is monotone and correctly rounded, so for a threshold r
//     fl(sqrt(s)) < r   <=>   s < T(r),   T(r) = min { t : fl(sqrt(t)) >= r }.
// With T the scan never takes a square root; with a cheap FMA estimate of s it never divides
// either unless the pair is close enough to matter -- and the decisions stay bit-identical to
// evaluating the reference's formula in plain C.
struct ScanThr {
    double T_viol;       // dist < rmin
    double T_coll;       // dist < rmin - coll_tol   (k = 1 only, solveSoftDMPCbound.m:25)
    double T_near[32];   // dist < neigh_thr(k)
    double inv_c;
    // high 32 bits of the exact thresholds, minus 1: decisions on the high word of the estimate
    // (scan_tile_hw).  he - h*_m (unsigned):  negative -> surely below;  0,1,2 -> ambiguous;  else surely above
    unsigned hv_m, hc_m, hn_m[32];
};
DMPC_HD unsigned hi_word(double v) {
#if defined(__CUDA_ARCH__)
    return (unsigned)__double2hiint(v);
#else
    unsigned long long b;
    memcpy(&b, &v, sizeof b);
    return (unsigned)(b >> 32);
#endif
}

inline double sq_threshold(double r) {
    if (!(r > 0.0)) return 0.0;  // dist < r is never true
    double t = r * r;
    while (sqrt(t) >= r) t = nextafter(t, 0.0);
    while (sqrt(t) < r) t = nextafter(t, INFINITY);
    return t;
}

inline ScanThr make_scan_thr(const DevParams& P) {
    ScanThr T;
    T.T_viol = sq_threshold(P.rmin);
    T.T_coll = sq_threshold(P.rmin - P.coll_tol);
    for (int k = 0; k < 32; ++k) {
        T.T_near[k] = (k < P.K) ? sq_threshold(neigh_thr(P, k + 1)) : 0.0;
    }
    T.inv_c = 1.0 / P.c;
    T.hv_m = hi_word(T.T_viol) - 1u;
    T.hc_m = hi_word(T.T_coll) - 1u;
    for (int k = 0; k < 32; ++k) T.hn_m[k] = hi_word(T.T_near[k]) - 1u;
    return T;
}

struct ScanAcc {
    unsigned vmask;  // per lane: bit k set if some neighbour of this lane violates at step k
    unsigned coll0;  // per lane: some neighbour is closer than rmin - coll_tol at step 1
};

// One (agent, neighbour) pair on this lane: pj = the neighbour's horizon (3K doubles), i = its agent index, own =
// this agent's previous horizon (3K doubles).  Accumulates the violation bits into acc and stores the K-bit near
// mask at nearmask[i].  Decisions are taken on the HIGH WORD
// of the FMA estimate of s with integer compares: for positive doubles  s < T  <=>  hi(s) < hi(T)  unless
// the high words are within 1 of each other -- only then (relative distance to a threshold < 2^-19, and
// the estimate is good to a few ulp) the pair goes through the reference's exact rounding sequence.  The
// fp64 pipe sees 7 operations per (neighbour, step); everything else is 32-bit integer work.
// KT: compile-time horizon (own[] then lives in registers after unrolling); 0 = run-time P.K.
template <int KT>
DMPC_D void scan_pair_hw(const DevParams& P, const ScanThr* __restrict__ thr, const double* __restrict__ own,
                         const double* __restrict__ pj, int i, unsigned* nearmask, ScanAcc& acc) {
    const int K = KT ? KT : P.K;
    unsigned nm = 0;
    {
        const double inv_c = thr->inv_c;
        const unsigned hv_m = thr->hv_m;
        unsigned vm = 0, amb = 0, c0 = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const double dx = own[3 * k] - pj[3 * k];
            const double dy = own[3 * k + 1] - pj[3 * k + 1];
            const double ez = (own[3 * k + 2] - pj[3 * k + 2]) * inv_c;
            const unsigned he = hi_word(fma(ez, ez, fma(dy, dy, dx * dx)));
            const unsigned xv = he - hv_m, xn = he - thr->hn_m[k];
            vm |= ((int)xv < 0 ? 1u : 0u) << k;
            nm |= ((int)xn < 0 ? 1u : 0u) << k;
            amb |= (xv <= 2u || xn <= 2u) ? 1u : 0u;
            if (k == 0) {  // rmin - coll_tol test of the first step (solveSoftDMPCbound.m:25)
                const unsigned xc = he - thr->hc_m;
                c0 = ((int)xc < 0) ? 1u : 0u;
                amb |= (xc <= 2u) ? 1u : 0u;
            }
        }
        if (amb) {
            // some estimate sits next to a threshold: redo this neighbour exactly
            vm = 0;
            nm = 0;
            c0 = 0;
            for (int k = 0; k < K; ++k) {
                const double s = ell_sq(own[3 * k] - pj[3 * k], own[3 * k + 1] - pj[3 * k + 1],
                                        own[3 * k + 2] - pj[3 * k + 2], P.c);
                if (s < thr->T_viol) vm |= 1u << k;
                if (s < thr->T_near[k]) nm |= 1u << k;
                if (k == 0 && s < thr->T_coll) c0 = 1u;
            }
        }
        acc.vmask |= vm;
        acc.coll0 |= c0;
    }
    nearmask[i] = nm;
}

// ONE neighbour per lane (cnt <= kLanes) starting at global index ibase whose horizons lie at `tile`
// (cnt x K x 3 doubles, same layout as l)
template <int KT>
DMPC_D void scan_tile_hw(const DevParams& P, const ScanThr* __restrict__ thr, const double* __restrict__ own, int n,
                         const double* __restrict__ tile, int ibase, int cnt, unsigned* nearmask, ScanAcc& acc) {
    const int K = KT ? KT : P.K;
    const int m = lane_id();
    if (m >= cnt) return;
    const int i = ibase + m;
    if (i != n) scan_pair_hw<KT>(P, thr, own, tile + (size_t)m * 3 * K, i, nearmask, acc);
    else nearmask[i] = 0u;
}

struct ScanOut {
    int kstar;     // 1-based, 0 = none
    int nv;        // rows written
    int flag;      // 0 / ST_COLL / ST_OVERFLOW
};

// Decide the first violating step and emit rows.  l = full horizon buffer (global).
// grow: d0[RMAX] d1[RMAX] d2[RMAX] dist[RMAX] rhs[RMAX]; gkc, gidx: RMAX ints.
// list: RMAX ints of scratch (shared memory in the kernel).  Two passes: (A) the near masks are compacted
// into the ordered list of (step, neighbour) pairs -- ballots and prefix counts only; (B) the rows of the
// list are computed side by side, one per lane (square root and division of all rows overlap).
// tiles_s != null: the neighbour horizons are resident in shared memory, tile t (32 neighbours) at
// tiles_s + ((t - tiles_rot) mod ntiles) * 32 * 3K.
DMPC_D ScanOut scan_finish(const DevParams& P, const double* __restrict__ own, int n,
                           const double* __restrict__ l, const unsigned* nearmask, ScanAcc acc,
                           int RMAX, double* grow, int* gkc, int* gidx, int* list,
                           const double* tiles_s = nullptr, int tiles_rot = 0, int ntiles = 0) {
    const int K = P.K, N = P.N;
    ScanOut o;
    o.kstar = 0;
    o.nv = 0;
    o.flag = 0;
    const unsigned vm = wor(acc.vmask);
    const unsigned coll0 = wor(acc.coll0);
    const bool soft = (P.variant == VAR_SOFT_BOUND || P.variant == VAR_SOFT_BOUND2);
    const double c2 = P.c * P.c;
    int kfirst = 0, klast = -1, kshift = 0;
    if (P.variant == VAR_HARD) {
        // solveHardDMPC.m:18-22: rows for every horizon step, k-major
        kfirst = 0;
        klast = K - 1;
        o.kstar = 0;
    } else {
        unsigned m = vm;
        if (soft && (m & 1u) && coll0) {
            // solveSoftDMPCbound.m:25-32: predicted collision at the very next step
            o.kstar = 1;
            o.flag = ST_COLL;
            return o;
        }
        if (P.variant == VAR_SOFT_BOUND2) {
            m &= ~1u;  // solveSoftDMPCbound2.m:29-31: k == 1 is skipped
            kshift = 1;  // CollConstrSoftDMPC2.m:8: k_ctr = k - 1
        }
        if (m == 0) return o;
        int ks = 0;
        while (!((m >> ks) & 1u)) ++ks;
        o.kstar = ks + 1;
        kfirst = klast = ks;
    }
    // (A) ordered compaction: rows in ascending neighbour order within a step (the reference's row order)
    int nv = 0;
    for (int k = kfirst; k <= klast; ++k) {
        for (int base = 0; base < N; base += kLanes) {
            const int i = base + lane_id();
            const bool hit = (i < N) && ((nearmask[i] >> k) & 1u);
            const unsigned bal = wballot(hit);
            if (hit) {
                const int slot = nv + popc_below(bal);
                if (slot < RMAX) list[slot] = (k << 26) | i;
            }
            nv += popc_all(bal);
        }
    }
    if (nv > RMAX) {
        o.flag = ST_QPFAIL | ST_OVERFLOW;
        nv = RMAX;
    }
    wsync();
    // (B) the rows
    for (int e = lane_id(); e < nv; e += kLanes) {
        const int code = list[e];
        const int k = code >> 26, i = code & 0x3ffffff;
        const double px = own[3 * k], py = own[3 * k + 1], pz = own[3 * k + 2];
        const double* pj;
        if (tiles_s) {
            int st = (i / 32) - tiles_rot;
            st += (st < 0) ? ntiles : 0;
            pj = tiles_s + ((size_t)st * 32 + (i & 31)) * 3 * K + 3 * k;
        } else {
            pj = l + 3 * ((size_t)k + (size_t)K * i);
        }
        const double dx = px - pj[0], dy = py - pj[1], dz = pz - pj[2];
        const double dist = ell_dist(dx, dy, dz, P.c);
        const double d0 = dx, d1 = dy, d2 = dz / c2;  // diff = E2 (p - pj)
        const double dp = d0 * px + d1 * py + d2 * pz;
        grow[e] = d0;
        grow[(size_t)RMAX + e] = d1;
        grow[2 * (size_t)RMAX + e] = d2;
        grow[3 * (size_t)RMAX + e] = dist;
        grow[4 * (size_t)RMAX + e] = dist * ((P.rmin - dist) + dp / dist);
        gkc[e] = k - kshift;
        if (gidx) gidx[e] = i;
    }
    o.nv = nv;
    if (P.variant == VAR_HARD) o.kstar = nv ? 1 : 0;
    return o;
}

}  // namespace dmpc

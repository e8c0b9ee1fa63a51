// agent_solve.cuh -- one agent's MPC step after the neighbour scan: weight selection, table
// staging, QP (qp_core.cuh) with the reference's infeasible-retry loop, state propagation.
//
// Reference: solveSoftDMPCbound.m:43-58 (weights), :60-98 (QP data), :102-155 (retry loop),
// :119-128 (propStatedmpc, is_inbounds); solveSoftDMPCbound2.m, solveHardDMPC.m,
// solveHardDMPCOnDemand.m for the variants; propStatedmpc.m:1-8; is_inbounds.m:1-6.
#pragma once
#include "qp_core.cuh"

namespace dmpc {

// status bits (same values as include/dmpc_b200.h)
enum { ST_SOLVED = 1, ST_COLL = 2, ST_INFEASIBLE = 4, ST_OUTBOUND = 8, ST_QPFAIL = 16, ST_OVERFLOW = 32 };
enum { VAR_SOFT_BOUND = 0, VAR_SOFT_BOUND2 = 1, VAR_HARD = 2, VAR_HARD_ONDEMAND = 3 };

// kernel-side copy of the parameters (plain data, passed by value)
struct DevParams {
    int K, variant, max_tries, neigh_mode, N;
    int ill_fallback = 1;  // infeasibility verdicts after an ill-conditioned add go to the generic solver (qp_warp.cuh)
    double h, rmin, c, alim, Q1, S1, term, Q_far, Q_near, S_free, near_radius, slack_lb,
        neigh_factor, coll_tol, inb_tol, hard_radius;
    double pmin[3], pmax[3];
};

struct AgentDiag {
    int kstar, nv, iters, nact;
};

DMPC_HD int round_up(int v, int m) { return (v + m - 1) / m * m; }

// bytes of per-agent scratch ("shared memory") for horizon K, active-set capacity QMAX and RCAP
// rows held on chip.  The constant tables are NOT part of it: they are staged once per thread
// block (all three weight sets, model_tables.h layout) and shared by the block's agents.
DMPC_HD size_t agent_smem_bytes(int K, int QMAX, int RCAP) {
    const int n3p = round_up(3 * K, 2);
    size_t nd = 8 * (size_t)n3p + 18 + 7 * (size_t)QMAX + (size_t)QMAX * QMAX + 9 * (size_t)RCAP;
    size_t ni = 3 * (size_t)QMAX + 4 * (size_t)n3p + 5 * (size_t)RCAP;
    return nd * sizeof(double) + round_up((int)ni, 4) * sizeof(int);  // multiple of 16 bytes
}

// layout of the table blob in global memory: see model_tables.h
DMPC_HD int tab_set_offset(int K, int wset) { return K * K + 4 * K + wset * 3 * K * K; }
// the blob of the register-resident solver (header copy padded to 4 doubles + interleaved T4 per set)
DMPC_HD int tab_fast_offset(int K) { return (K * K + 4 * K + 9 * K * K + 3) / 4 * 4; }
DMPC_HD int tab_fast_header(int K) { return (K * K + 4 * K + 3) / 4 * 4; }
DMPC_HD int tab_fast_size(int K) { return tab_fast_header(K) + 12 * K * K; }

struct AgentIO {
    // inputs
    const double *po, *pf, *vo, *ao;  // 3 each (global)
    int kstar;     // 1-based first violating step (0 none)
    int nv;        // rows produced by the scan
    int scanflag;  // ST_COLL / ST_OVERFLOW from the scan, else 0
    int RMAX;      // stride of the global row arrays
    const double* grow;  // global rows: d0[RMAX] d1[RMAX] d2[RMAX] dist[RMAX] rhs[RMAX]
    const int* gkc;      // global rows: kc[RMAX]
    double* gscr_d;      // global scratch (used when nv > RCAP): 4*RMAX doubles
    int* gscr_i;         // 4*RMAX ints
    // outputs (global), 3K each; v_hor / a_hor may be null
    double *out_p, *out_v, *out_a;
    double *p1, *v1, *a1;  // 3 each
    const double* l_prev_n;  // previous horizon of this agent (copied to out_p when unsolved)
    const double* bounds = nullptr;  // workspace box of the agent's scenario: pmin[3], pmax[3]; null = P.pmin / P.pmax
    // cross-step warm start of the fast path (qp_warp.cuh warm_start / warm_store); both optional
    const int* gidx = nullptr;  // global rows: neighbour's agent index of each row [RMAX]
    int* warm = nullptr;        // per agent kWarmStride ints: [0] = count, [1..] = the stored active set
    int dbg_n = -1;             // agent index (debug traces only)
    int start_tries = 0;        // generic solver: first try of the retry loop (slack bound and penalty doubled that often)
};
constexpr int kWarmStride = 68;

// All lanes of the warp call this with identical arguments.  Returns the status word.
// tab: the whole table blob (model_tables.h layout), normally resident in shared memory.
// KT: compile-time horizon length (0 = run-time P.K).
template <int KT>
DMPC_D int agent_solve(const DevParams& P, const double* __restrict__ tab, unsigned char* smem,
                       int QMAX, int RCAP, const AgentIO& io, AgentDiag* diag_out) {
    const int K = KT ? KT : P.K, n3 = 3 * K, n3p = round_up(n3, 2);
    const int lane = lane_id();
    AgentDiag dg;
    dg.kstar = io.kstar;
    dg.nv = io.nv;
    dg.iters = 0;
    dg.nact = 0;
    int status = 0;

    if (io.scanflag) {
        status = io.scanflag;
    } else {
        // ---- carve the scratch ---------------------------------------------------------------
        double* dptr = reinterpret_cast<double*>(smem);
        double* s_Punc = dptr; dptr += n3p;
        double* s_aunc = dptr; dptr += n3p;
        double* s_a = dptr; dptr += n3p;
        double* s_P = dptr; dptr += n3p;
        double* s_cbox = dptr; dptr += n3p;
        double* s_cP = dptr; dptr += n3p;
        double* s_z = dptr; dptr += n3p;
        double* s_Lz = dptr; dptr += n3p;
        double* s_x0 = dptr; dptr += 12;   // po, pf, vo, ao
        double* s_bnd = dptr; dptr += 6;   // pmin, pmax
        double* s_u = dptr; dptr += QMAX;
        double* s_g = dptr; dptr += QMAX;
        double* s_r = dptr; dptr += QMAX;
        double* s_sv0 = dptr; dptr += QMAX;
        double* s_sv1 = dptr; dptr += QMAX;
        double* s_sv2 = dptr; dptr += QMAX;
        double* s_se = dptr; dptr += QMAX;
        double* s_M = dptr; dptr += (size_t)QMAX * QMAX;
        double* s_rows = dptr; dptr += 5 * (size_t)RCAP;
        double* s_rnorm = dptr; dptr += RCAP;
        double* s_eps = dptr; dptr += RCAP;
        double* s_zeps = dptr; dptr += RCAP;
        double* s_rres = dptr; dptr += RCAP;
        int* iptr = reinterpret_cast<int*>(dptr);
        int* s_act = iptr; iptr += QMAX;
        int* s_ralist = iptr; iptr += QMAX;
        int* s_sinfo = iptr; iptr += QMAX;
        int* s_sbl = iptr; iptr += n3p;
        int* s_sbu = iptr; iptr += n3p;
        int* s_swl = iptr; iptr += n3p;
        int* s_swu = iptr; iptr += n3p;
        int* s_rslot = iptr; iptr += RCAP;
        int* s_ubslot = iptr; iptr += RCAP;
        int* s_lbslot = iptr; iptr += RCAP;
        int* s_rmat = iptr; iptr += RCAP;
        int* s_rkc = iptr; iptr += RCAP;

        // ---- weights (solveSoftDMPCbound.m:43-58) -------------------------------------------
        const bool soft = (P.variant == VAR_SOFT_BOUND || P.variant == VAR_SOFT_BOUND2);
        // solveHardDMPC quirk: Ain_coll is never empty for N >= 2 -> collision weights always
        const bool any_violation = (P.variant == VAR_HARD) ? (P.N >= 2) : (io.kstar > 0);
        for (int x = lane; x < 3; x += kLanes) {
            s_x0[x] = io.po[x];
            s_x0[3 + x] = io.pf[x];
            s_x0[6 + x] = io.vo[x];
            s_x0[9 + x] = io.ao[x];
            s_bnd[x] = io.bounds ? io.bounds[x] : P.pmin[x];
            s_bnd[3 + x] = io.bounds ? io.bounds[3 + x] : P.pmax[x];
        }
        wsync();
        const double *x_po = s_x0, *x_pf = s_x0 + 3, *x_vo = s_x0 + 6, *x_ao = s_x0 + 9;
        const double dgx = x_po[0] - x_pf[0], dgy = x_po[1] - x_pf[1], dgz = x_po[2] - x_pf[2];
        const double dgoal = sqrt(dgx * dgx + dgy * dgy + dgz * dgz);
        int wset;
        double qw, sw;
        if (!any_violation && dgoal >= P.near_radius) { wset = 0; qw = P.Q_far; sw = P.S_free; }
        else if (!any_violation) { wset = 1; qw = P.Q_near; sw = P.S_free; }
        else { wset = 2; qw = P.Q1; sw = P.S1; }

        // ---- tables of this agent's weight set (blob: lam, tt, lnorm, then G,B,C per set) --------
        const double* t_lam = tab;
        const double* t_tt = tab + K * K;
        const double* t_lnorm = tab + K * K + K;
        const double* t_ilnorm = tab + K * K + 2 * K;
        const double* t_G = tab + tab_set_offset(K, wset);
        const double* t_B = t_G + K * K;
        const double* t_C = t_B + K * K;

        // ---- rows: on chip when they fit, else in place in global memory --------------------
        const int nv = io.nv;
        const double *rd0, *rd1, *rd2, *rdist, *rrhs;
        const int* rkc;
        double *irnorm, *eps, *zeps, *rres;
        int *rslot, *ubslot, *lbslot, *rmat;
        if (nv <= RCAP) {
            for (int f = 0; f < 5; ++f)
                for (int j = lane; j < nv; j += kLanes) s_rows[f * RCAP + j] = io.grow[(size_t)f * io.RMAX + j];
            for (int j = lane; j < nv; j += kLanes) s_rkc[j] = io.gkc[j];
            rd0 = s_rows; rd1 = s_rows + RCAP; rd2 = s_rows + 2 * RCAP; rdist = s_rows + 3 * RCAP;
            rrhs = s_rows + 4 * RCAP;
            rkc = s_rkc; irnorm = s_rnorm; eps = s_eps; zeps = s_zeps; rres = s_rres;
            rslot = s_rslot; ubslot = s_ubslot; lbslot = s_lbslot; rmat = s_rmat;
        } else {
            rd0 = io.grow; rd1 = io.grow + io.RMAX; rd2 = io.grow + 2 * (size_t)io.RMAX;
            rdist = io.grow + 3 * (size_t)io.RMAX; rrhs = io.grow + 4 * (size_t)io.RMAX;
            rkc = io.gkc; irnorm = io.gscr_d; eps = io.gscr_d + io.RMAX; zeps = io.gscr_d + 2 * (size_t)io.RMAX;
            rres = io.gscr_d + 3 * (size_t)io.RMAX;
            rslot = io.gscr_i; ubslot = io.gscr_i + io.RMAX; lbslot = io.gscr_i + 2 * (size_t)io.RMAX;
            rmat = io.gscr_i + 3 * (size_t)io.RMAX;
        }
        wsync();
        for (int j = lane; j < nv; j += kLanes) {
            const double dd = rd0[j] * rd0[j] + rd1[j] * rd1[j] + rd2[j] * rd2[j];
            const double ln = t_lnorm[rkc[j]];
            irnorm[j] = 1.0 / sqrt(dd * ln * ln + (soft ? rdist[j] * rdist[j] : 0.0));
        }

        // ---- a_unc = -G f,  P_unc = A_initp [po;vo] + Lam a_unc ------------------------------------
        // f_x = -2 ( q e_x lamK + s ao_x e_0 ),  e = pf - (po + t_K vo)   (solveSoftDMPCbound.m:82-88)
        for (int i = lane; i < n3; i += kLanes) {
            const int k = i / 3, x = i - 3 * k;
            const double e = x_pf[x] - (x_po[x] + t_tt[K - 1] * x_vo[x]);
            s_aunc[i] = 2.0 * qw * e * t_B[k * K + (K - 1)] + 2.0 * sw * x_ao[x] * t_G[k * K];
        }
        wsync();
        for (int i = lane; i < n3; i += kLanes) {
            const int k = i / 3, x = i - 3 * k;
            double s = 0.0;
            for (int j = 0; j <= k; ++j) s = fma(t_lam[k * K + j], s_aunc[3 * j + x], s);
            s_Punc[i] = s + (x_po[x] + t_tt[k] * x_vo[x]);
        }
        wsync();

        Qp<KT> qp;
        qp.w.K = K; qp.w.n3 = n3; qp.w.QMAX = QMAX;
        qp.w.lam = t_lam; qp.w.ilnorm = t_ilnorm; qp.w.G = t_G; qp.w.B = t_B; qp.w.C = t_C;
        qp.w.alim = P.alim; qp.w.soft = soft ? 1 : 0; qp.w.qw = qw; qp.w.sw = sw;
        qp.w.bnd = s_bnd;
        qp.w.aunc = s_aunc; qp.w.Punc = s_Punc;
        qp.w.nv = nv;
        qp.w.rd0 = rd0; qp.w.rd1 = rd1; qp.w.rd2 = rd2; qp.w.rdist = rdist; qp.w.rrhs = rrhs;
        qp.w.rkc = rkc; qp.w.irnorm = irnorm;
        qp.w.a = s_a; qp.w.P = s_P; qp.w.cbox = s_cbox; qp.w.cP = s_cP; qp.w.z = s_z; qp.w.Lz = s_Lz;
        qp.w.eps = eps; qp.w.zeps = zeps; qp.w.rres = rres;
        qp.w.kc_all = (P.variant == VAR_HARD) ? -1 : (io.kstar > 0 ? io.kstar - 1 - (P.variant == VAR_SOFT_BOUND2 ? 1 : 0) : 0);
        qp.w.sbl = s_sbl; qp.w.sbu = s_sbu; qp.w.swl = s_swl; qp.w.swu = s_swu;
        qp.w.rslot = rslot; qp.w.ubslot = ubslot; qp.w.lbslot = lbslot; qp.w.rmat = rmat;
        qp.w.act = s_act; qp.w.u = s_u; qp.w.g = s_g; qp.w.r = s_r; qp.w.M = s_M;
        qp.w.ralist = s_ralist;
        qp.w.sv0 = s_sv0; qp.w.sv1 = s_sv1; qp.w.sv2 = s_sv2; qp.w.se = s_se; qp.w.sinfo = s_sinfo;

        // ---- retry loop (solveSoftDMPCbound.m:102-155) -------------------------------------------
        double term = P.term, slb = P.slack_lb;
        int tries = 0;
        // re-solve on behalf of the register-resident solver: the tries it has already found infeasible (exact 3-D
        // certificate or an unambiguous verdict) are not repeated
        for (; tries < io.start_tries && tries < P.max_tries; ++tries) {
            slb *= 2.0;
            term *= 2.0;
        }
        bool solved = false;
        const int max_iter = 40 * (n3 + nv) + 200;
        bool warm = false;
        for (;;) {
            qp.w.term = term;
            qp.w.slb = slb;
            if (warm) {
                // same constraint normals, new (term, slb): keep the active set and its inverse
                qp.warm_restart();
            } else {
                for (int i = lane; i < n3; i += kLanes) {
                    s_a[i] = s_aunc[i];
                    s_P[i] = s_Punc[i];
                    s_sbl[i] = -1; s_sbu[i] = -1; s_swl[i] = -1; s_swu[i] = -1;
                }
                for (int j = lane; j < nv; j += kLanes) {
                    eps[j] = 0.0;
                    rslot[j] = -1; ubslot[j] = -1; lbslot[j] = -1; rmat[j] = 0;
                }
                wsync();
                qp.reset();
                qp.rows_refresh();
            }
            const QpResult r = qp.solve(max_iter);
            dg.iters += r.iters;
            dg.nact = r.q;
            if (r.rc == QP_OK) { solved = true; break; }
            if (r.rc == QP_ITERCAP) { status |= ST_QPFAIL; break; }
            if (r.rc == QP_OVERFLOW) { status |= ST_QPFAIL | ST_OVERFLOW; break; }
            // infeasible: soft variants with slack double the slack bound and the penalty and retry;
            // otherwise the reference only loosens quadprog's tolerance or gives up.
            if (!(soft && nv > 0)) break;
            slb *= 2.0;
            term *= 2.0;
            if (++tries >= P.max_tries) break;
            warm = true;
        }
        status |= (tries & 0xff) << 8;
        if (solved) {
            status |= ST_SOLVED;
            // propStatedmpc.m: p = A_p a + A_initp [po;vo] (= P), v = A_v a + vo
            for (int i = lane; i < n3; i += kLanes) {
                const int k = i / 3, x = i - 3 * k;
                double sv = 0.0;
                for (int j = 0; j <= k; ++j) sv += P.h * s_a[3 * j + x];
                const double vv = sv + x_vo[x];
                io.out_p[i] = s_P[i];
                if (io.out_v) io.out_v[i] = vv;
                if (io.out_a) io.out_a[i] = s_a[i];
                if (k == 0) {
                    io.p1[x] = s_P[i];
                    io.v1[x] = vv;
                    io.a1[x] = s_a[i];
                }
            }
            // is_inbounds.m on the first predicted position
            bool inb = true;
#pragma unroll
            for (int x = 0; x < 3; ++x)
                inb = inb && (s_P[x] < s_bnd[3 + x] + P.inb_tol) && (s_P[x] > s_bnd[x] - P.inb_tol);
            if (!inb) status |= ST_OUTBOUND;
        } else if (!(status & ST_QPFAIL)) {
            status |= ST_INFEASIBLE;
        }
    }
    if (!(status & ST_SOLVED)) {
        // the reference returns empty p,v,a: the caller keeps the old horizon and state
        for (int i = lane; i < n3; i += kLanes) io.out_p[i] = io.l_prev_n[i];
        for (int x = lane; x < 3; x += kLanes) {
            io.p1[x] = io.po[x];
            io.v1[x] = io.vo[x];
            io.a1[x] = io.ao[x];
        }
    }
    if (diag_out && lane == 0) *diag_out = dg;
    return status;
}

}  // namespace dmpc

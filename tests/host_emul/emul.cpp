// emul.cpp -- TEST-ONLY host build of the device algorithm with one "lane".
//
// The CUDA kernels' per-agent code (multiagent_planning_b200/csrc/{scan_core,agent_solve,qp_core}.cuh)
// is written against a small lane abstraction; compiled by g++ it runs the identical algorithm
// sequentially.  This lets the CPU test-suite (-m "not gpu") check the ALGORITHM against the
// oracle where no GPU exists.  It is a debugging aid: it is built by tests/ into
// tests/host_emul/libemul.so, never linked into libdmpc_b200.so and never imported by the package.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/dmpc_b200.h"
#include "../../multiagent_planning_b200/csrc/model_tables.h"
#include "../../multiagent_planning_b200/csrc/scan_core.cuh"
#include "../../multiagent_planning_b200/csrc/qp_warp.cuh"

using namespace dmpc;

static DevParams to_dev(const dmpcb200_params* p, int N, const double* pmin, const double* pmax) {
    DevParams D;
    D.K = p->K; D.variant = p->variant; D.max_tries = p->max_tries; D.neigh_mode = p->neigh_mode;
    D.N = N; D.h = p->h; D.rmin = p->rmin; D.c = p->c; D.alim = p->alim; D.Q1 = p->Q1; D.S1 = p->S1;
    D.term = p->term; D.Q_far = p->Q_far; D.Q_near = p->Q_near; D.S_free = p->S_free;
    D.near_radius = p->near_radius; D.slack_lb = p->slack_lb; D.neigh_factor = p->neigh_factor;
    D.coll_tol = p->coll_tol; D.inb_tol = p->inb_tol; D.hard_radius = p->hard_radius;
    for (int x = 0; x < 3; ++x) { D.pmin[x] = pmin[x]; D.pmax[x] = pmax[x]; }
    return D;
}

extern "C" int emul_step(const dmpcb200_params* p, int N, int n0, int n1, const double* pk,
                         const double* vk, const double* ak, const double* pf, const double* l_prev,
                         const double* pmin, const double* pmax, int QMAX, int RCAP, int RMAX,
                         double* l_new, double* p1, double* v1, double* a1, double* v_hor,
                         double* a_hor, int32_t* status, int32_t* diag /*4 per agent*/,
                         int32_t* warm /*optional: kWarmStride per agent, kept by the caller across steps*/) {
    const int K = p->K;
    DevParams D = to_dev(p, N, pmin, pmax);
    std::vector<double> tab;
    const double qs[3][2] = {{p->Q_far, p->S_free}, {p->Q_near, p->S_free}, {p->Q1, p->S1}};
    build_tables(p->h, K, qs, tab);
    // QMAX < 0: the register-resident solver of qp_warp.cuh (capacity -QMAX) where its preconditions
    // hold, the generic solver elsewhere -- the dispatch of qp_kernel
    const bool fast = QMAX < 0;
    if (fast) QMAX = -QMAX;
    std::vector<unsigned char> smem(std::max(agent_smem_bytes(K, QMAX, RCAP), qw_smem_bytes()) + 64);
    const ScanThr thr = make_scan_thr(D);
    std::vector<unsigned> nearmask(N);
    std::vector<double> grow(5 * (size_t)RMAX), gscr_d(4 * (size_t)RMAX);
    std::vector<int> gkc(RMAX), gidx(RMAX), gscr_i(4 * (size_t)RMAX), list(RMAX);
    for (int n = n0; n < n1; ++n) {
        const double* own = l_prev + (size_t)3 * K * n;
        ScanAcc acc;
        acc.vmask = 0;
        acc.coll0 = 0;
        for (int i = 0; i < N; ++i)  // the kernel's tile function: one neighbour per lane, high-word decisions
            scan_tile_hw<0>(D, &thr, own, n, l_prev + (size_t)3 * K * i, i, 1, nearmask.data(), acc);
        ScanOut so = scan_finish(D, own, n, l_prev, nearmask.data(), acc, RMAX, grow.data(), gkc.data(),
                                 gidx.data(), list.data());
        AgentIO io;
        io.po = pk + 3 * n; io.pf = pf + 3 * n; io.vo = vk + 3 * n; io.ao = ak + 3 * n;
        io.kstar = so.kstar; io.nv = so.nv; io.scanflag = so.flag; io.RMAX = RMAX;
        io.grow = grow.data(); io.gkc = gkc.data(); io.gscr_d = gscr_d.data(); io.gscr_i = gscr_i.data();
        io.out_p = l_new + (size_t)3 * K * n;
        io.out_v = v_hor ? v_hor + (size_t)3 * K * n : nullptr;
        io.out_a = a_hor ? a_hor + (size_t)3 * K * n : nullptr;
        io.p1 = p1 + 3 * n; io.v1 = v1 + 3 * n; io.a1 = a1 + 3 * n;
        io.l_prev_n = own;
        io.gidx = gidx.data();
        io.dbg_n = n;
        io.warm = warm ? warm + (size_t)kWarmStride * n : nullptr;
        AgentDiag dg;
        if (fast && 3 * K <= kQW && so.nv <= kRowsFastMax && !so.flag)
            status[n] = (D.variant == VAR_HARD)
                            ? agent_solve_fast<0, kQW, true>(D, tab.data() + tables_fast_offset(K), smem.data(), QMAX, io, &dg)
                            : agent_solve_fast<0>(D, tab.data() + tables_fast_offset(K), smem.data(), QMAX, io, &dg);
        else
            status[n] = agent_solve<0>(D, tab.data(), smem.data(), QMAX, RCAP, io, &dg);
        if (fast && QMAX == kQW && (status[n] & ST_OVERFLOW) && !so.flag) {
            // the kernel's rescue path: the generic solver with a large capacity decides (active set beyond the
            // on-chip capacity, no free slot in the row working set, numerically inconsistent active set)
            const int it0 = dg.iters;
            io.start_tries = (status[n] >> 8) & 0xff;
            std::vector<unsigned char> big(agent_smem_bytes(K, 3 * kQW, RMAX) + 64);
            status[n] = agent_solve<0>(D, tab.data(), big.data(), 3 * kQW, RMAX, io, &dg);
            dg.iters += it0;
        }
        if (diag) { diag[4 * n] = dg.kstar; diag[4 * n + 1] = dg.nv; diag[4 * n + 2] = dg.iters; diag[4 * n + 3] = dg.nact; }
    }
    return 0;
}

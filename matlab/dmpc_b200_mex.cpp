// dmpc_b200_mex.cpp -- thin MEX gateway from MATLAB to libdmpc_b200.so (include/dmpc_b200.h).
//
// SOURCE ONLY in this repository: the build container has neither MATLAB nor mex.h, so this file is
// compiled where MATLAB exists (recipe in INTEGRATION.md):
//     mex -R2018a -I../include dmpc_b200_mex.cpp -L../multiagent_planning_b200 -ldmpc_b200
// Everything it calls is exercised through the same C-ABI by the Python ctypes binding in the
// test-suite.
//
// One persistent handle per (N, K, variant, weights) lives across calls (mexLock); the gateway is
// non-reentrant like every MEX function (MATLAB calls it on its single interpreter thread).
//
//   [l,pk,vk,ak]            = dmpc_b200_mex('init',  P, po, pf, pmin, pmax)     % initDMPC.m for all agents
//   [l_new,pk,vk,ak,status] = dmpc_b200_mex('step',  P, pk, vk, ak, l)          % the `for n = 1:N` body
//   [pk,vk,ak,steps,reached,fail_step,fail_agent] = dmpc_b200_mex('run', P, max_steps, stop_on_fail)
//   [p,v,a,status]          = dmpc_b200_mex('solve', P, po, pf, vo, ao, n, l, pmin, pmax)
//   [viol,min_dist,violc]   = dmpc_b200_mex('check', P, p, l, n, k)
//   [Ain,bin,prev_dist]     = dmpc_b200_mex('constr',P, p, po, vo, n, k, l, mask)
//   [pass,max_dist]         = dmpc_b200_mex('goal',  P, pk, pf, tol)            % ReachedGoal.m
//   [p,v]                   = dmpc_b200_mex('prop',  P, po, vo, a)
//   [pk,vk,ak,p,v,a,info]   = dmpc_b200_mex('post',  P, pk, vk, ak)            % failure_rate.m:134-195
//   [A,Av,A0,Delta]         = dmpc_b200_mex('mats',  h, K)
//   dmpc_b200_mex('close')
// P is a struct with the fields of dmpcb200_params (missing fields = reference defaults) plus N.
// MATLAB arrays are column-major fp64: l is 3 x K x N, states 3 x N -- exactly the library layout,
// so pointers are passed through (mxGetDoubles) with no repacking.
#include <cstring>
#include <string>

#include "dmpc_b200.h"
#include "mex.h"

static dmpcb200_t* g_h = nullptr;
static dmpcb200_params g_p;
static int g_N = 0;

static void close_handle() {
    if (g_h) dmpcb200_destroy(g_h);
    g_h = nullptr;
}

static void check(int rc, const char* what) {
    if (rc != 0) mexErrMsgIdAndTxt("dmpcb200:api", "%s: %s", what, dmpcb200_last_error());
}

static double field(const mxArray* s, const char* name, double dflt) {
    const mxArray* f = mxIsStruct(s) ? mxGetField(s, 0, name) : nullptr;
    return f ? mxGetScalar(f) : dflt;
}

static void get_params(const mxArray* s, dmpcb200_params* p, int* N) {
    const int variant = (int)field(s, "variant", 0);
    dmpcb200_default_params(p, variant);
    p->K = (int)field(s, "K", p->K);
    p->max_tries = (int)field(s, "max_tries", p->max_tries);
    p->neigh_mode = (int)field(s, "neigh_mode", p->neigh_mode);
#define F(n) p->n = field(s, #n, p->n)
    F(h); F(rmin); F(c); F(alim); F(Q1); F(S1); F(term); F(Q_far); F(Q_near); F(S_free); F(near_radius);
    F(slack_lb); F(neigh_factor); F(coll_tol); F(inb_tol); F(hard_radius); F(init_div); F(goal_tol);
#undef F
    *N = (int)field(s, "N", 0);
    if (*N < 1) mexErrMsgIdAndTxt("dmpcb200:arg", "P.N (number of agents) is required");
}

static void ensure_handle(const mxArray* s) {
    dmpcb200_params p;
    int N;
    get_params(s, &p, &N);
    if (g_h && N == g_N && std::memcmp(&p, &g_p, sizeof(p)) == 0) return;
    close_handle();
    check(dmpcb200_create(&p, N, 0, N, /*n_scenarios*/ 1, (int)field(s, "device", 0), 0, &g_h), "create");
    g_p = p;
    g_N = N;
    if (!mexIsLocked()) {
        mexLock();
        mexAtExit(close_handle);
    }
}

static mxArray* mat(int r, int c) { return mxCreateDoubleMatrix(r, c, mxREAL); }
static mxArray* cube(int a, int b, int c) {
    const mwSize d[3] = {(mwSize)a, (mwSize)b, (mwSize)c};
    return mxCreateNumericArray(3, d, mxDOUBLE_CLASS, mxREAL);
}
static mxArray* i32(int n) { return mxCreateNumericMatrix(n, 1, mxINT32_CLASS, mxREAL); }

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (nrhs < 1 || !mxIsChar(prhs[0])) mexErrMsgIdAndTxt("dmpcb200:arg", "first argument: command string");
    char cmd[32];
    mxGetString(prhs[0], cmd, sizeof(cmd));
    const std::string c(cmd);
    if (c == "close") {
        close_handle();
        if (mexIsLocked()) mexUnlock();
        return;
    }
    if (c == "mats") {
        const double h = mxGetScalar(prhs[1]);
        const int K = (int)mxGetScalar(prhs[2]);
        plhs[0] = mat(3 * K, 3 * K); plhs[1] = mat(3 * K, 3 * K); plhs[2] = mat(3 * K, 6); plhs[3] = mat(3 * K, 3 * K);
        check(dmpcb200_model_mats(h, K, mxGetDoubles(plhs[0]), mxGetDoubles(plhs[1]), mxGetDoubles(plhs[2]),
                                  mxGetDoubles(plhs[3])), "model_mats");
        return;
    }
    ensure_handle(prhs[1]);
    const int N = g_N, K = g_p.K;
    if (c == "init") {
        check(dmpcb200_set_goals(g_h, mxGetDoubles(prhs[3])), "set_goals");
        check(dmpcb200_set_bounds(g_h, mxGetDoubles(prhs[4]), mxGetDoubles(prhs[5])), "set_bounds");
        plhs[0] = cube(3, K, N); plhs[1] = mat(3, N); plhs[2] = mat(3, N); plhs[3] = mat(3, N);
        check(dmpcb200_init_horizons(g_h, mxGetDoubles(prhs[2]), mxGetDoubles(plhs[0]), mxGetDoubles(plhs[1]),
                                     mxGetDoubles(plhs[2]), mxGetDoubles(plhs[3])), "init_horizons");
    } else if (c == "step") {
        plhs[0] = mxDuplicateArray(prhs[5]);  // agents that fail keep their horizon
        plhs[1] = mxDuplicateArray(prhs[2]); plhs[2] = mxDuplicateArray(prhs[3]); plhs[3] = mxDuplicateArray(prhs[4]);
        plhs[4] = i32(N);
        int32_t ff = -1;
        check(dmpcb200_step(g_h, mxGetDoubles(prhs[2]), mxGetDoubles(prhs[3]), mxGetDoubles(prhs[4]),
                            mxGetDoubles(prhs[5]), mxGetDoubles(plhs[0]), mxGetDoubles(plhs[1]), mxGetDoubles(plhs[2]),
                            mxGetDoubles(plhs[3]), nullptr, nullptr, (int32_t*)mxGetData(plhs[4]), nullptr, &ff), "step");
        if (nlhs > 5) plhs[5] = mxCreateDoubleScalar(ff < 0 ? 0 : ff + 1);  // 1-based first failing agent, 0 none
    } else if (c == "run") {
        const int S = (int)mxGetScalar(prhs[2]);
        const int stop = nrhs > 3 ? (int)mxGetScalar(prhs[3]) : 0;
        plhs[0] = cube(3, S + 1, N); plhs[1] = cube(3, S + 1, N); plhs[2] = cube(3, S + 1, N);
        int32_t steps = 0, reached = 0, fs = -1, fa = -1;
        check(dmpcb200_run(g_h, S, stop, 0, mxGetDoubles(plhs[0]), mxGetDoubles(plhs[1]), mxGetDoubles(plhs[2]),
                           nullptr, &steps, &reached, &fs, &fa), "run");
        plhs[3] = mxCreateDoubleScalar(steps); plhs[4] = mxCreateDoubleScalar(reached);
        plhs[5] = mxCreateDoubleScalar(fs < 0 ? 0 : fs + 1); plhs[6] = mxCreateDoubleScalar(fa < 0 ? 0 : fa + 1);
    } else if (c == "solve") {
        check(dmpcb200_set_bounds(g_h, mxGetDoubles(prhs[8]), mxGetDoubles(prhs[9])), "set_bounds");
        plhs[0] = mat(3, K); plhs[1] = mat(3, K); plhs[2] = mat(3, K);
        int32_t st = 0;
        check(dmpcb200_solve_agent(g_h, mxGetDoubles(prhs[2]), mxGetDoubles(prhs[3]), mxGetDoubles(prhs[4]),
                                   mxGetDoubles(prhs[5]), (int)mxGetScalar(prhs[6]) - 1, mxGetDoubles(prhs[7]),
                                   mxGetDoubles(plhs[0]), mxGetDoubles(plhs[1]), mxGetDoubles(plhs[2]), &st, nullptr),
              "solve_agent");
        plhs[3] = mxCreateDoubleScalar(st);
    } else if (c == "check") {
        plhs[0] = mxCreateLogicalMatrix(N, 1); plhs[2] = mxCreateLogicalMatrix(N, 1);
        double md = 0; int32_t any = 0;
        check(dmpcb200_check_coll(g_h, mxGetDoubles(prhs[2]), mxGetDoubles(prhs[3]), (int)mxGetScalar(prhs[4]) - 1,
                                  (int)mxGetScalar(prhs[5]), (uint8_t*)mxGetLogicals(plhs[0]),
                                  (uint8_t*)mxGetLogicals(plhs[2]), &md, &any), "check_coll");
        plhs[1] = mxCreateDoubleScalar(md);
    } else if (c == "constr") {
        // rows: at most N-1 neighbours; n = 0 (no own agent, dec-iSCP CollConstr.m) allows N; optional 10th argument
        const int cap = nrhs > 9 ? (int)mxGetScalar(prhs[9]) : (N > 1 ? N - 1 : 1);
        mxArray* A = mat(cap, 3 * K); mxArray* b = mat(cap, 1); mxArray* pd = mat(cap, 1);
        int32_t nr = 0;
        const uint8_t* mask = (nrhs > 8 && !mxIsEmpty(prhs[8])) ? (const uint8_t*)mxGetLogicals(prhs[8]) : nullptr;
        check(dmpcb200_coll_constr(g_h, mxGetDoubles(prhs[2]), mxGetDoubles(prhs[3]), mxGetDoubles(prhs[4]),
                                   (int)mxGetScalar(prhs[5]) - 1, (int)mxGetScalar(prhs[6]), mxGetDoubles(prhs[7]), mask,
                                   cap, mxGetDoubles(A), mxGetDoubles(b), mxGetDoubles(pd), &nr), "coll_constr");
        // trim to nr rows (column-major: copy the leading nr rows of every column)
        plhs[0] = mat(nr, 3 * K); plhs[1] = mat(nr, 1); plhs[2] = mat(nr, 1);
        for (int j = 0; j < 3 * K; ++j)
            std::memcpy(mxGetDoubles(plhs[0]) + (size_t)nr * j, mxGetDoubles(A) + (size_t)cap * j, nr * sizeof(double));
        std::memcpy(mxGetDoubles(plhs[1]), mxGetDoubles(b), nr * sizeof(double));
        std::memcpy(mxGetDoubles(plhs[2]), mxGetDoubles(pd), nr * sizeof(double));
        mxDestroyArray(A); mxDestroyArray(b); mxDestroyArray(pd);
    } else if (c == "goal") {
        // [pass,max_dist] = dmpc_b200_mex('goal', P, pk, pf, tol)      % ReachedGoal.m:1-11, pk / pf are 3 x N
        double md = 0; int32_t pass = 0;
        check(dmpcb200_reached_goal(g_h, mxGetDoubles(prhs[2]), mxGetDoubles(prhs[3]), mxGetScalar(prhs[4]), &md, &pass),
              "reached_goal");
        plhs[0] = mxCreateDoubleScalar(pass);
        if (nlhs > 1) plhs[1] = mxCreateDoubleScalar(md);
    } else if (c == "prop") {
        const int B = (int)(mxGetNumberOfElements(prhs[4]) / (3 * K));
        plhs[0] = mat(3 * K, B); plhs[1] = mat(3 * K, B);
        check(dmpcb200_prop_state(g_h, B, mxGetDoubles(prhs[2]), mxGetDoubles(prhs[3]), mxGetDoubles(prhs[4]),
                                  mxGetDoubles(plhs[0]), mxGetDoubles(plhs[1])), "prop_state");
    } else if (c == "post") {
        // [pk,vk,ak,p,v,a,info] = dmpc_b200_mex('post', P, pk, vk, ak)   % failure_rate.m:134-195
        const int S = (int)(mxGetNumberOfElements(prhs[2]) / (3 * (size_t)N));
        plhs[0] = mxDuplicateArray(prhs[2]); plhs[1] = mxDuplicateArray(prhs[3]); plhs[2] = mxDuplicateArray(prhs[4]);
        dmpcb200_post res;
        // first pass for the size of the 100 Hz grid, second with the outputs allocated
        mxArray* tmp[3] = {mxDuplicateArray(prhs[2]), mxDuplicateArray(prhs[3]), mxDuplicateArray(prhs[4])};
        check(dmpcb200_postprocess(g_h, S, mxGetDoubles(tmp[0]), mxGetDoubles(tmp[1]), mxGetDoubles(tmp[2]), 2.0, 1.0,
                                   0.01, 0.05, nullptr, nullptr, nullptr, 0, nullptr, &res), "postprocess");
        for (auto* t : tmp) mxDestroyArray(t);
        plhs[3] = cube(3, res.nt, N); plhs[4] = cube(3, res.nt, N); plhs[5] = cube(3, res.nt, N);
        check(dmpcb200_postprocess(g_h, S, mxGetDoubles(plhs[0]), mxGetDoubles(plhs[1]), mxGetDoubles(plhs[2]), 2.0,
                                   1.0, 0.01, 0.05, mxGetDoubles(plhs[3]), mxGetDoubles(plhs[4]),
                                   mxGetDoubles(plhs[5]), res.nt, nullptr, &res), "postprocess");
        if (nlhs > 6) {
            plhs[6] = mat(1, 7);
            double* o = mxGetDoubles(plhs[6]);  // r_factor, h_scaled, T, violation, min_dist, totdist, traj_time
            o[0] = res.r_factor; o[1] = res.h_scaled; o[2] = res.T; o[3] = res.violation; o[4] = res.min_dist;
            o[5] = res.totdist; o[6] = res.traj_time;
        }
    } else {
        mexErrMsgIdAndTxt("dmpcb200:arg", "unknown command %s", cmd);
    }
}

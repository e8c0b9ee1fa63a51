// dmpc_b200.cu -- libdmpc_b200.so: handle, launches and the C-ABI of include/dmpc_b200.h.
//
// Built for sm_100a only.  There is no CPU path: every compute entry point needs a CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cmath>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/dmpc_b200.h"
#if defined(DMPC_SINGLE_TU)  // profiling build: one translation unit (one copy of the cycle counters)
#include "k_qp15.cu"
#include "k_qp20.cu"
#include "k_qpgen.cu"
#include "k_qphard.cu"
#include "k_scan.cu"
#endif
#include "launch.cuh"
#include "small_kernels.cuh"
#include "postprocess.cuh"
#include "model_tables.h"

using namespace dmpc;

static_assert(sizeof(AgentDiag) == sizeof(dmpcb200_diag), "diag layout");

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(DMPCB200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));    \
    } while (0)

template <typename T>
cudaError_t dalloc(T** p, size_t n) {
    cudaError_t e = cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T));
    if (e == cudaSuccess) e = cudaMemset(*p, 0, std::max<size_t>(n, 1) * sizeof(T));
    return e;
}

}  // namespace

struct dmpcb200_handle {
    dmpcb200_params prm;
    DevParams dp;
    int N = 0, n0 = 0, n1 = 0, NL = 0, Npad = 0, K = 0, device = 0;
    int n_static = 0;            // agents [N - n_static, N) are static obstacles (dmpcb200_set_static_obstacles)
    int S = 1;                   // scenarios batched in this handle (independent swarms of N agents)
    double* d_bounds = nullptr;  // S x 6: pmin, pmax of every scenario (dmpcb200_set_scenario)
    bool per_scen_bounds = false;
    std::vector<char> scen_set;  // which scenarios have been set
    int RMAX = 0, QMAX = 0, RCAP = 0, QBIG = 0, W = 4, n_rescue = 0;
    size_t rescue_bytes = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;  // host step: the small state copies run beside the horizon copy + scan
    cudaEvent_t ev_in = nullptr;
    // resident state
    double* d_tab = nullptr;
    // resident state lives in ONE arena so that the host step moves it with one copy per direction:
    //   side b: [ l (Npad*3K) | pk | vk | ak ]   then   [ status | diag | first_fail ]
    unsigned char* d_arena = nullptr;
    size_t side_bytes = 0, tailblk_bytes = 0;
    unsigned char* h_stage = nullptr;  // pinned + mapped: one input side + one output side + tail block
    unsigned char* d_stage = nullptr;  // device alias of h_stage (the QP kernel writes the outputs of a host step there)
    // caller-owned host arrays that are pinned (page-locked + mapped) are used in place: DMA straight from the
    // inputs, and the QP kernel writes a host step's outputs straight into the device alias of the outputs.
    struct Bound {
        const double *pk, *vk, *ak, *l_prev;
        double *l_new, *p1, *v1, *a1, *v_hor, *a_hor;
        int32_t* status;
        dmpcb200_diag* diag;
        bool in_pinned;          // resolved once at bind time (the arrays' lifetime is the caller's contract)
        double *w_l, *w_p, *w_v, *w_a;  // device aliases of l_new, p1, v1, a1 (all four or none)
    };
    std::vector<Bound> bound;  // dmpcb200_bind_step slots
    double* d_l[2] = {nullptr, nullptr};
    double* d_st[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};  // pk, vk, ak ping-pong
    double* d_pf = nullptr;
    double *d_vhor = nullptr, *d_ahor = nullptr;
    int* d_status = nullptr;
    AgentDiag* d_diag = nullptr;
    int cur = 0;  // index of the current l / state
    // scratch
    ScanRec* d_scan = nullptr;
    double *d_grow = nullptr, *d_gscr_d = nullptr;
    int *d_gkc = nullptr, *d_gidx = nullptr, *d_gscr_i = nullptr;
    unsigned char* d_rescue = nullptr;
    int* d_rescue_next = nullptr;
    unsigned* d_done = nullptr;  // CTA arrival counter of the fused tail
    // throughput layout of the QP kernel (launches of more than one wave of agents)
    RouteQ* d_rq = nullptr;
    int *d_qlight = nullptr, *d_qheavy = nullptr;
    int layout = 1;  // 1 classic persistent kernel (default), 0 two-role throughput kernel (DMPCB200_LAYOUT=throughput)
    Ctrl* d_ctrl = nullptr;
    double* d_goal = nullptr;  // 2 doubles
    int* d_fail = nullptr;
    // scratch of the per-agent drop-ins (solve_agent / check_coll / coll_constr): they never touch the
    // resident loop state.  [ l_in | l_out | pk vk ak | p1 v1 a1 | pf ], allocated at first use
    double* d_scr = nullptr;
    // helper scratch
    unsigned char* d_u8 = nullptr;  // 2N
    double* d_small = nullptr;      // misc
    int* d_ismall = nullptr;
    // trajectory record of dmpcb200_run
    double *d_traj[3] = {nullptr, nullptr, nullptr};
    int* d_hist = nullptr;
    int traj_S = 0, traj_scen = 1;
    bool hist_on = false;
    double batch_ms = 0;            // dmpcb200_run_batch: device time of the whole call
    long long batch_agent_steps = 0;  //   and the agent-steps it solved
    // closed-loop graph (two steps: even -> odd -> even)
    cudaGraphExec_t graph = nullptr;
    bool graph_record = false, graph_batch = false;
    // pinned host staging
    double* h_pin = nullptr;
    size_t h_pin_n = 0;
    // timing
    std::vector<cudaEvent_t> ev;
    double t_ms[3] = {0, 0, 0};
    double t_host_us[4] = {0, 0, 0, 0};  // host step: pack, submit, wait, unpack
    int64_t launches = 0;
    bool have_bounds = false, have_goals = false, have_init = false;
};

namespace {

int ensure_device(dmpcb200_t* h) {
    CK(cudaSetDevice(h->device));
    return 0;
}

StepArgs make_args(dmpcb200_t* h, int n0, int n1, const double* pk, const double* vk, const double* ak,
                   const double* l_prev, double* l_new, double* p1, double* v1, double* a1, double* v_hor,
                   double* a_hor, int* status, AgentDiag* diag, bool padded, Ctrl* ctrl) {
    StepArgs A;
    A.P = h->dp;
    A.thr = make_scan_thr(h->dp);
    A.n0 = n0;
    A.n1 = n1;
    A.n_scen = 1;
    A.lstride = 0;
    A.bounds = nullptr;
    A.RMAX = h->RMAX;
    A.QMAX = h->QMAX;
    A.RCAP = h->RCAP;
    A.QBIG = h->QBIG;
    A.n_rescue = h->n_rescue;
    A.tile_padded = padded ? 1 : 0;
    A.l_prev = l_prev;
    A.l_new = l_new;
    A.pk = pk;
    A.vk = vk;
    A.ak = ak;
    A.pf = h->d_pf;
    A.p1 = p1;
    A.v1 = v1;
    A.a1 = a1;
    A.v_hor = v_hor;
    A.a_hor = a_hor;
    A.status = status;
    A.diag = diag;
    A.tab = h->d_tab;
    A.scan = h->d_scan;
    A.grow = h->d_grow;
    A.gkc = h->d_gkc;
    A.gidx = h->d_gidx;
    A.gscr_d = h->d_gscr_d;
    A.gscr_i = h->d_gscr_i;
    A.rescue = h->d_rescue;
    A.rescue_bytes = h->rescue_bytes;
    A.rescue_next = h->d_rescue_next;
    A.ctrl = ctrl;
    A.fuse_tail = 0;
    A.done_cnt = h->d_done;
    A.work_cnt = h->d_done + 1;
    A.rq = nullptr;
    A.q_light = h->d_qlight;
    A.q_heavy = h->d_qheavy;
    return A;
}

// throughput layout: more agents than one wave of the one-agent-per-sub-partition layout, reference horizon,
// variant with a fast path.  Decided per launch (a per-agent drop-in on the same handle is a batch of one).
bool use_throughput_layout(const dmpcb200_t* h, const StepArgs& A) {
    if (h->layout == 1 || !h->d_rq || h->W != 4) return false;
    if (h->K != 15 && h->K != 20) return false;
    if (A.diag != h->d_diag) return false;  // the route prediction reads the handle's own record of the last step
    return (long long)(A.n1 - A.n0) * A.n_scen > 4LL * sm_count(current_device());
}

// 4 agents x 2 warps per CTA while one wave of CTAs covers the swarm, 8 agents beyond (or fewer when the
// per-agent near masks of a very large swarm do not fit next to a tile); the reference's horizons
// (15, 20) are compiled with the horizon as a constant (the own horizon then lives in registers)
cudaError_t launch_scan(dmpcb200_t* h, const StepArgs& A, cudaStream_t s) {
    const int nl = A.n1 - A.n0, K = h->K, N = h->N;  // (per scenario)
    if (const char* e = getenv("DMPCB200_SCAN_LAYOUT")) {  // experiment hook
        const int id = *e ? atoi(e) : -1;
        if (id >= 0 && id < SCAN_LAYOUTS) return launch_scan_layout((ScanLayout)id, A, nl, K, s);
    }
    // throughput mode (batched scenarios / more than one wave of agents) with every tile of a scenario resident:
    // 8 agents x 2 warps per CTA with the own horizon in shared memory (occupancy beats the register-resident
    // horizon: C5 scan 247 -> 150 us per step)
    if ((A.rq || A.n_scen > 1 || (long long)nl * A.n_scen > 4LL * sm_count(current_device())) &&
        scan_stages(K, N, 8, A.RMAX) >= (N + kTile - 1) / kTile)
        return launch_scan_layout(SCAN_8_2_0, A, nl, K, s);
    // single swarms with a compiled-in horizon: the register-tile layout (a warp per tile, the CTA's agents looped
    // over: 2.5x fewer shared-memory wavefronts per pair).  Measured (scan per step, N = 500 / 2000 / C4): 16.7 ->
    // 15.4 us, 120 -> 75 us (14 agents per CTA: one wave of 143 CTAs), 200 -> 126 us; bit-identical rows.
    const int n_sm = sm_count(current_device());
    if (nl <= 4 * n_sm || scan_stages(K, N, 8, A.RMAX) < 2) {
        if (scan_stages(K, N, 4, A.RMAX) < 1) return launch_scan_layout(SCAN_1_2_0, A, nl, K, s);
        if (K == 15) return launch_scan_layout(SCAN_RT_4_8_15, A, nl, K, s);
        if (K == 20) return launch_scan_layout(SCAN_RT_4_8_20, A, nl, K, s);
        return launch_scan_layout(SCAN_4_4_0, A, nl, K, s);
    }
    if (K == 15) {
        // one wave of CTAs when 14 agents per CTA cover the swarm and 8 tiles still fit beside their masks
        if ((nl + 13) / 14 <= n_sm && scan_stages(K, N, 14, A.RMAX) >= 8) return launch_scan_layout(SCAN_RT_14_8_15, A, nl, K, s);
        return launch_scan_layout(SCAN_RT_8_8_15, A, nl, K, s);
    }
    if (K == 20) return launch_scan_layout(SCAN_RT_8_8_20, A, nl, K, s);
    return launch_scan_layout(SCAN_8_2_0, A, nl, K, s);
}
// horizon lengths 15 and 20 (the reference's configurations) are compiled with the horizon as a
// compile-time constant (fully unrolled table products); anything else takes the generic kernel
cudaError_t launch_qp(dmpcb200_t* h, const StepArgs& A, cudaStream_t s) {
    const int nl = A.n1 - A.n0;
    if (A.rq) return h->K == 15 ? launch_qp2_15(A, nl, s) : launch_qp2_20(A, nl, s);
    const size_t smem = qp_smem_bytes(h->K, h->W, h->QMAX, h->RCAP);
    if (h->dp.variant == VAR_HARD) {
        if (h->K == 15 && h->W == 4) return launch_qp_hard_4_15(A, nl, smem, s);
        return h->W == 4 ? launch_qp_hard_4_0(A, nl, smem, s) : launch_qp_hard_3_0(A, nl, smem, s);
    }
    if (h->K == 15 && h->W == 4) return launch_qp_4_15(A, nl, smem, s);
    if (h->K == 20 && h->W == 4) return launch_qp_4_20(A, nl, smem, s);
    // W = 4 for every horizon up to 21, 3 beyond (the tables grow with K^2)
    if (h->W == 4) return launch_qp_4_0(A, nl, smem, s);
    return launch_qp_3_0(A, nl, smem, s);
}

TailArgs make_tail(dmpcb200_t* h, const double* p, int ld, const int* status, const double* p1, const double* v1,
                   const double* a1, bool record, Ctrl* ctrl) {
    TailArgs T;
    T.N = h->N - h->n_static;  // goal test, failure scan and records cover the commanded agents
    T.n0 = h->n0;
    T.n1 = h->n1;
    T.ld = ld;
    T.goal_tol = h->prm.goal_tol;
    T.p = p;
    T.pf = h->d_pf;
    T.status = status;
    T.p1 = p1;
    T.v1 = v1;
    T.a1 = a1;
    T.traj_p = record ? h->d_traj[0] : nullptr;
    T.traj_v = record ? h->d_traj[1] : nullptr;
    T.traj_a = record ? h->d_traj[2] : nullptr;
    T.status_hist = (record && h->hist_on) ? h->d_hist : nullptr;
    T.S = h->traj_S;
    T.goal_out = h->d_goal;
    T.fail_out = h->d_fail;
    T.rescue_next = h->d_rescue_next;
    T.ctrl = ctrl;
    T.copy_src = nullptr;
    T.copy_dst = nullptr;
    T.copy_bytes = 0;
    return T;
}

// one full resident step: state[cur] -> state[cur^1]
int launch_resident_step(dmpcb200_t* h, int cur, bool record, Ctrl* ctrl, cudaStream_t s, cudaEvent_t* evs) {
    const int nx = cur ^ 1;
    StepArgs A = make_args(h, h->n0, h->n1, h->d_st[cur][0], h->d_st[cur][1], h->d_st[cur][2], h->d_l[cur],
                           h->d_l[nx], h->d_st[nx][0], h->d_st[nx][1], h->d_st[nx][2], nullptr, nullptr,
                           h->d_status, h->d_diag, true, ctrl);
    A.T = make_tail(h, h->d_st[nx][0], 3, h->d_status, h->d_st[nx][0], h->d_st[nx][1], h->d_st[nx][2], record, ctrl);
    A.fuse_tail = 1;
    if (evs) CK(cudaEventRecord(evs[0], s));
    if (use_throughput_layout(h, A)) A.rq = h->d_rq;
    CK(launch_scan(h, A, s));
    if (evs) CK(cudaEventRecord(evs[1], s));
    CK(launch_qp(h, A, s));  // the tail of the step runs in the last CTA of the QP kernel
    if (evs) CK(cudaEventRecord(evs[2], s));
    if (evs) CK(cudaEventRecord(evs[3], s));
    return 0;
}

int ensure_events(dmpcb200_t* h, size_t n) {
    while (h->ev.size() < n) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        h->ev.push_back(e);
    }
    return 0;
}

int ensure_pin(dmpcb200_t* h, size_t n) {
    if (h->h_pin_n >= n) return 0;
    if (h->h_pin) cudaFreeHost(h->h_pin);
    h->h_pin = nullptr;
    h->h_pin_n = 0;
    CK(cudaMallocHost((void**)&h->h_pin, n * sizeof(double)));
    h->h_pin_n = n;
    return 0;
}

// device alias of a caller-owned host array if ALL of it is pinned and mapped, else nullptr.  Queried on every
// call (a cached answer would go stale when the caller frees the buffer and the address is reused by a
// pageable allocation); only dmpcb200_bind_step keeps the answer, where the arrays' lifetime is the caller's
// contract.  First and last byte must belong to the same mapping.
void* pinned_alias(const void* p, size_t bytes) {
    if (!p || !bytes) return nullptr;
    cudaPointerAttributes a0, a1;
    if (cudaPointerGetAttributes(&a0, p) != cudaSuccess || a0.type != cudaMemoryTypeHost || !a0.devicePointer) {
        cudaGetLastError();
        return nullptr;
    }
    const char* last = static_cast<const char*>(p) + bytes - 1;
    if (cudaPointerGetAttributes(&a1, last) != cudaSuccess || a1.type != cudaMemoryTypeHost || !a1.devicePointer ||
        static_cast<char*>(a1.devicePointer) - static_cast<char*>(a0.devicePointer) != (ptrdiff_t)(bytes - 1)) {
        cudaGetLastError();
        return nullptr;
    }
    return a0.devicePointer;
}

// scratch of the per-agent drop-ins: pointers into h->d_scr
struct Scratch {
    double *l_in, *l_out, *st_in[3], *st_out[3], *pf;
    int* status;
    AgentDiag* diag;
};
int get_scratch(dmpcb200_t* h, Scratch* S) {
    const size_t nl = (size_t)h->Npad * 3 * h->K, ns = 3 * (size_t)round_up(h->N, 2);
    const size_t nst = ((size_t)h->N + 1) / 2 + 1, ndg = 2 * (size_t)h->N;  // int[N], AgentDiag[N] in doubles
    if (!h->d_scr) CK(dalloc(&h->d_scr, 2 * nl + 7 * ns + nst + ndg));
    double* d = h->d_scr;
    S->l_in = d; d += nl;
    S->l_out = d; d += nl;
    for (int q = 0; q < 3; ++q) { S->st_in[q] = d; d += ns; }
    for (int q = 0; q < 3; ++q) { S->st_out[q] = d; d += ns; }
    S->pf = d; d += ns;
    S->diag = reinterpret_cast<AgentDiag*>(d); d += ndg;
    S->status = reinterpret_cast<int*>(d);
    return 0;
}

void drop_graph(dmpcb200_t* h) {
    if (h->graph) cudaGraphExecDestroy(h->graph);
    h->graph = nullptr;
}

int check_arch(int device) {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(DMPCB200_ERR_CUDA, std::string("device ") + prop.name +
                                           " is not sm_100: libdmpc_b200 is built for B200 (sm_100a) only");
    return 0;
}

}  // namespace

extern "C" {

int dmpcb200_abi_version(void) { return DMPCB200_ABI_VERSION; }
const char* dmpcb200_last_error(void) { return g_err.c_str(); }
int dmpcb200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

void dmpcb200_default_params(dmpcb200_params* p, int variant) {
    // test/failure_rate.m:7-27, dmpc_soft_bound.m:7-30, solveSoftDMPCbound.m, dmpc/cpp/dmpc.h:50-67
    std::memset(p, 0, sizeof(*p));
    p->K = 15;
    p->variant = variant;
    p->max_tries = 30;
    p->neigh_mode = 0;
    p->h = 0.2;
    p->rmin = 0.35;
    p->c = 2.0;
    p->alim = 1.0;
    p->Q1 = 1000.0;
    p->S1 = (variant == DMPCB200_HARD || variant == DMPCB200_HARD_ONDEMAND) ? 10.0 : 100.0;  // dmpc_hard.m
    p->term = -5e4;
    p->Q_far = 1000.0;
    p->Q_near = 10000.0;
    p->S_free = 10.0;
    p->near_radius = 1.0;
    p->slack_lb = (variant == DMPCB200_SOFT_BOUND2) ? -0.01 : -0.05;
    p->neigh_factor = 3.0;
    p->coll_tol = 0.05;
    p->inb_tol = 0.05;
    p->hard_radius = 1.0;
    p->init_div = 10.0;
    p->goal_tol = 0.01;
}

void dmpcb200_default_params_cpp(dmpcb200_params* p, int k_factor) {
    // the semantics of the C++ port, solveQPv2 (dmpc/cpp/dmpc.cpp:803-1287), on top of the MATLAB defaults:
    dmpcb200_default_params(p, k_factor == -1 ? DMPCB200_SOFT_BOUND2 : DMPCB200_SOFT_BOUND);  // k_ctr = k + k_factor (:516,:890)
    p->neigh_mode = 1;    // neighbour threshold rmin (1 + k / k_hor) (:418)
    p->slack_lb = -0.01;  // slack as rows  eps <= 0, -eps <= lim  with lim = 0.01 (:907-914, :1077)
    p->max_tries = 21;    // one solve + `while (status && tries < 20)` with lim and term doubled (:1079-1109)
    p->term = -1e6;       // `int term = -1000000` (:846)
    p->Q1 = 1000.0;       // collision weights are hard-coded (:940-945)
    p->S1 = 100.0;
    // struct Params defaults (dmpc.h:65-67)
    p->h = 0.2; p->K = 12; p->c = 1.5; p->rmin = 0.5; p->alim = 2.0; p->goal_tol = 0.05; p->coll_tol = 0.05;
}

int dmpcb200_model_mats(double h, int K, double* A_p, double* A_v, double* A_initp, double* Delta) {
    if (K < 1 || K > 32 || !(h > 0)) return fail(DMPCB200_ERR_ARG, "model_mats: need 1 <= K <= 32, h > 0");
    model_mats(h, K, A_p, A_v, A_initp, Delta);
    return 0;
}

int dmpcb200_create(const dmpcb200_params* p, int N, int n0, int n1, int n_scenarios, int device, int max_rows,
                    dmpcb200_t** out) {
    if (!p || !out) return fail(DMPCB200_ERR_ARG, "create: null argument");
    *out = nullptr;
    if (p->K < 1 || p->K > 32) return fail(DMPCB200_ERR_ARG, "create: horizon K must be in 1..32");
    // n0 == n1 is a valid (empty) block: a rank of a sharded run whose block is empty still owns a replica of
    // the horizon buffer and takes part in the exchange
    if (N < 1 || n0 < 0 || n1 > N || n0 > n1) return fail(DMPCB200_ERR_ARG, "create: bad agent range");
    if (n_scenarios < 1) return fail(DMPCB200_ERR_ARG, "create: n_scenarios must be >= 1");
    if (n_scenarios > 1 && (n0 != 0 || n1 != N))
        return fail(DMPCB200_ERR_ARG, "create: batched scenarios are sharded whole (n0 = 0, n1 = N)");
    if (p->variant < 0 || p->variant > 3) return fail(DMPCB200_ERR_ARG, "create: unknown variant");
    if (!(p->h > 0) || !(p->rmin > 0) || !(p->c > 0) || !(p->alim > 0))
        return fail(DMPCB200_ERR_ARG, "create: h, rmin, c, alim must be positive");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(DMPCB200_ERR_CUDA, "create: no CUDA device (libdmpc_b200 has no CPU path)");
    }
    if (device < 0 || device >= ndev) return fail(DMPCB200_ERR_ARG, "create: bad device index");
    if (int rc = check_arch(device)) return rc;
    CK(cudaSetDevice(device));

    dmpcb200_t* h = new dmpcb200_handle();
    h->prm = *p;
    h->N = N;
    h->n0 = n0;
    h->n1 = n1;
    h->NL = n1 - n0;
    h->K = p->K;
    h->device = device;
    h->Npad = round_up(N, kTile);
    h->S = n_scenarios;
    h->scen_set.assign(n_scenarios, 0);
    DevParams& D = h->dp;
    D.K = p->K; D.variant = p->variant; D.max_tries = p->max_tries; D.neigh_mode = p->neigh_mode; D.N = N;
    D.h = p->h; D.rmin = p->rmin; D.c = p->c; D.alim = p->alim; D.Q1 = p->Q1; D.S1 = p->S1; D.term = p->term;
    D.Q_far = p->Q_far; D.Q_near = p->Q_near; D.S_free = p->S_free; D.near_radius = p->near_radius;
    D.slack_lb = p->slack_lb; D.neigh_factor = p->neigh_factor; D.coll_tol = p->coll_tol;
    D.inb_tol = p->inb_tol; D.hard_radius = p->hard_radius;
    if (const char* e = getenv("DMPCB200_ILL_FALLBACK")) D.ill_fallback = atoi(e);  // experiment hook (default 1)
    for (int x = 0; x < 3; ++x) { D.pmin[x] = -1e30; D.pmax[x] = 1e30; }

    const int K = p->K, n3 = 3 * K;
    // capacities
    if (max_rows > 0) h->RMAX = max_rows;
    else if (p->variant == DMPCB200_HARD) h->RMAX = std::min(K * std::max(N - 1, 1), 1024);
    else h->RMAX = std::min(std::max(N - 1, 1), 256);
    h->RMAX = round_up(h->RMAX, 2);
    h->RCAP = 64;
    const int qwant = std::min(round_up(n3 + 16, 8), 64);  // on-chip active-set capacity (qp_warp.cuh: 64)
    const size_t smem_max = 227 * 1024;
    h->W = 0;
    for (int w : {4, 3})
        if (qp_smem_bytes(K, w, qwant, h->RCAP) <= smem_max) {
            h->W = w;
            break;
        }
    if (!h->W) {
        delete h;
        return fail(DMPCB200_ERR_ARG, "create: horizon too long for the on-chip QP workspace");
    }
    h->QMAX = qwant;
    if (const char* e = getenv("DMPCB200_QMAX")) {  // test hook: shrink the on-chip active-set capacity
        const int v = atoi(e);
        if (v >= 2 && v <= qwant) h->QMAX = round_up(v, 2);
    }
    h->QBIG = round_up(n3 + 2 * std::min(h->RMAX, 256) + 8, 8);
    h->n_rescue = 64;
    h->rescue_bytes = align_up(agent_smem_bytes(K, h->QBIG, h->RCAP), 256);

    auto bail = [&](cudaError_t e, const char* what) {
        std::string m = std::string("create: ") + what + ": " + cudaGetErrorString(e);
        dmpcb200_destroy(h);
        return fail(DMPCB200_ERR_CUDA, m);
    };
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "stream");
    if ((e = cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "stream");
    if ((e = cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "event");
    // tables
    std::vector<double> tab;
    const double qs[3][2] = {{p->Q_far, p->S_free}, {p->Q_near, p->S_free}, {p->Q1, p->S1}};
    build_tables(p->h, K, qs, tab);
    if ((e = dalloc(&h->d_tab, tab.size())) != cudaSuccess) return bail(e, "tables");
    if ((e = cudaMemcpy(h->d_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess)
        return bail(e, "tables copy");
    {
        const size_t nS = (size_t)h->S;
        const size_t lB = align_up(nS * h->Npad * n3 * sizeof(double), 256);
        const size_t sB = align_up(3 * nS * N * sizeof(double), 256);
        h->side_bytes = lB + 3 * sB;
        const size_t stB = align_up(nS * N * sizeof(int), 256), dgB = align_up(nS * N * sizeof(AgentDiag), 256);
        h->tailblk_bytes = stB + dgB + align_up(nS * sizeof(int), 256);
        const size_t total = 2 * h->side_bytes + h->tailblk_bytes;
        if ((e = cudaMalloc((void**)&h->d_arena, total)) != cudaSuccess) return bail(e, "state arena");
        if ((e = cudaMemset(h->d_arena, 0, total)) != cudaSuccess) return bail(e, "state arena");
        unsigned char* s0 = h->d_arena;
        unsigned char* s1 = h->d_arena + h->side_bytes;
        unsigned char* tb = h->d_arena + 2 * h->side_bytes;
        unsigned char* sides[2] = {s0, s1};
        for (int b = 0; b < 2; ++b) {
            h->d_l[b] = reinterpret_cast<double*>(sides[b]);
            for (int q = 0; q < 3; ++q) h->d_st[b][q] = reinterpret_cast<double*>(sides[b] + lB + q * sB);
        }
        h->d_status = reinterpret_cast<int*>(tb);
        h->d_diag = reinterpret_cast<AgentDiag*>(tb + stB);
        h->d_fail = reinterpret_cast<int*>(tb + stB + dgB);
        if ((e = cudaHostAlloc((void**)&h->h_stage, 2 * h->side_bytes + h->tailblk_bytes, cudaHostAllocMapped)) !=
            cudaSuccess)
            return bail(e, "pinned staging");
        if ((e = cudaHostGetDevicePointer((void**)&h->d_stage, h->h_stage, 0)) != cudaSuccess)
            return bail(e, "pinned staging (device alias)");
    }
    const size_t nS = (size_t)h->S;
    if ((e = dalloc(&h->d_pf, 3 * nS * N)) != cudaSuccess) return bail(e, "goals");
    if ((e = dalloc(&h->d_bounds, 6 * nS)) != cudaSuccess) return bail(e, "bounds");
    if ((e = dalloc(&h->d_vhor, nS * N * n3)) != cudaSuccess) return bail(e, "v_hor");
    if ((e = dalloc(&h->d_ahor, nS * N * n3)) != cudaSuccess) return bail(e, "a_hor");
    const size_t NL = (size_t)h->NL * nS;
    if ((e = dalloc(&h->d_scan, NL)) != cudaSuccess) return bail(e, "scan");
    if ((e = dalloc(&h->d_grow, NL * 5 * h->RMAX)) != cudaSuccess) return bail(e, "rows");
    if ((e = dalloc(&h->d_gkc, NL * h->RMAX)) != cudaSuccess) return bail(e, "rows kc");
    if ((e = dalloc(&h->d_gidx, NL * h->RMAX)) != cudaSuccess) return bail(e, "rows idx");
    if ((e = dalloc(&h->d_gscr_d, NL * 4 * h->RMAX)) != cudaSuccess) return bail(e, "row scratch");
    if ((e = dalloc(&h->d_gscr_i, NL * 4 * h->RMAX)) != cudaSuccess) return bail(e, "row scratch");
    if ((e = dalloc(&h->d_rescue, h->rescue_bytes * h->n_rescue)) != cudaSuccess) return bail(e, "rescue");
    if ((e = dalloc(&h->d_rescue_next, 1)) != cudaSuccess) return bail(e, "rescue counter");
    if ((e = dalloc(&h->d_done, 2)) != cudaSuccess) return bail(e, "done / work counters");
    {
        const char* lay = getenv("DMPCB200_LAYOUT");
        // The two-role kernel (8 agents per SM, light agents with a 32-capacity active set) is an OPT-IN experiment:
        // measured on one B200 it wins where light agents dominate (C5 late steps: QP 159 -> 123 us per step) and
        // loses where heavy agents do (C5 first 30 steps 711 -> 864 us, N = 2000 286 -> 466 us): net zero to
        // negative, so the classic persistent kernel stays the default.
        h->layout = (lay && std::string(lay) == "throughput") ? 0 : 1;
        int n_sm = 148;
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
        if (NL > 4 * (size_t)n_sm && h->layout == 0) {
            const size_t cap = NL + 2 * (size_t)kLightW * n_sm + 64;  // + the indices the draining warps overshoot by
            if ((e = dalloc(&h->d_rq, 1)) != cudaSuccess) return bail(e, "route queues");
            if ((e = dalloc(&h->d_qlight, NL)) != cudaSuccess) return bail(e, "route queues");
            if ((e = cudaMalloc((void**)&h->d_qheavy, cap * sizeof(int))) != cudaSuccess) return bail(e, "route queues");
            if ((e = cudaMemset(h->d_qheavy, 0xff, cap * sizeof(int))) != cudaSuccess) return bail(e, "route queues");
        }
    }
    if ((e = dalloc(&h->d_ctrl, nS)) != cudaSuccess) return bail(e, "ctrl");
    if ((e = dalloc(&h->d_goal, 2 * nS)) != cudaSuccess) return bail(e, "goal");
    if ((e = dalloc(&h->d_u8, 2 * (size_t)N)) != cudaSuccess) return bail(e, "u8");
    if ((e = dalloc(&h->d_small, 64)) != cudaSuccess) return bail(e, "small");
    if ((e = dalloc(&h->d_ismall, 16)) != cudaSuccess) return bail(e, "ismall");
    *out = h;
    return 0;
}

void dmpcb200_destroy(dmpcb200_t* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    drop_graph(h);
    for (auto e : h->ev) cudaEventDestroy(e);
    cudaFree(h->d_tab);
    cudaFree(h->d_arena);
    if (h->h_stage) cudaFreeHost(h->h_stage);
    cudaFree(h->d_pf); cudaFree(h->d_bounds); cudaFree(h->d_vhor); cudaFree(h->d_ahor);
    cudaFree(h->d_scan); cudaFree(h->d_grow); cudaFree(h->d_gkc); cudaFree(h->d_gidx);
    cudaFree(h->d_gscr_d); cudaFree(h->d_gscr_i); cudaFree(h->d_rescue); cudaFree(h->d_rescue_next);
    cudaFree(h->d_rq); cudaFree(h->d_qlight); cudaFree(h->d_qheavy);
    cudaFree(h->d_done); cudaFree(h->d_ctrl); cudaFree(h->d_goal); cudaFree(h->d_u8); cudaFree(h->d_small);
    cudaFree(h->d_ismall);
    cudaFree(h->d_scr);
    for (int s = 0; s < 3; ++s) cudaFree(h->d_traj[s]);
    cudaFree(h->d_hist);
    if (h->h_pin) cudaFreeHost(h->h_pin);
    if (h->ev_in) cudaEventDestroy(h->ev_in);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int dmpcb200_set_bounds(dmpcb200_t* h, const double* pmin, const double* pmax) {
    if (!h || !pmin || !pmax) return fail(DMPCB200_ERR_ARG, "set_bounds: null argument");
    for (int x = 0; x < 3; ++x) {
        if (!(pmin[x] < pmax[x])) return fail(DMPCB200_ERR_ARG, "set_bounds: need pmin < pmax");
        h->dp.pmin[x] = pmin[x];
        h->dp.pmax[x] = pmax[x];
    }
    h->have_bounds = true;
    drop_graph(h);
    return 0;
}

int dmpcb200_set_goals(dmpcb200_t* h, const double* pf) {
    if (!h || !pf) return fail(DMPCB200_ERR_ARG, "set_goals: null argument");
    if (h->S != 1) return fail(DMPCB200_ERR_STATE, "set_goals: single-scenario entry point on a batched handle (use set_scenario / run_batch / get_scenario)");
    if (int rc = ensure_device(h)) return rc;
    CK(cudaMemcpyAsync(h->d_pf, pf, 3 * (size_t)h->N * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->have_goals = true;
    return 0;
}

int dmpcb200_set_static_obstacles(dmpcb200_t* h, int n_cmd) {
    if (!h) return fail(DMPCB200_ERR_ARG, "set_static_obstacles: null handle");
    if (h->S != 1 || h->n0 != 0 || h->n1 + h->n_static != h->N)
        return fail(DMPCB200_ERR_STATE, "set_static_obstacles: needs a single-scenario handle that owns all agents");
    if (n_cmd < 1 || n_cmd > h->N) return fail(DMPCB200_ERR_ARG, "set_static_obstacles: need 1 <= n_cmd <= N");
    h->n_static = h->N - n_cmd;
    h->n1 = n_cmd;
    h->NL = n_cmd;
    h->have_init = false;
    drop_graph(h);
    return 0;
}

int dmpcb200_init_horizons(dmpcb200_t* h, const double* po, double* l, double* p1, double* v1, double* a1) {
    if (!h || !po) return fail(DMPCB200_ERR_ARG, "init_horizons: null argument");
    if (h->S != 1) return fail(DMPCB200_ERR_STATE, "init_horizons: single-scenario entry point on a batched handle (use set_scenario / run_batch / get_scenario)");
    if (!h->have_goals) return fail(DMPCB200_ERR_STATE, "init_horizons: call dmpcb200_set_goals first");
    if (int rc = ensure_device(h)) return rc;
    const int N = h->N, K = h->K;
    cudaStream_t s = h->stream;
    h->cur = 0;
    double* d_po = h->d_st[1][0];  // staging: the other state buffer
    CK(cudaMemcpyAsync(d_po, po, 3 * (size_t)N * sizeof(double), cudaMemcpyHostToDevice, s));
    if (h->n_static) {
        // un-commanded agents (dmpc.cpp:1633-1649) stay where they are: their goal is their start point, so
        // initDMPC.m gives them a constant horizon
        const size_t o = 3 * (size_t)h->n1;
        CK(cudaMemcpyAsync(h->d_pf + o, po + o, 3 * (size_t)h->n_static * sizeof(double), cudaMemcpyHostToDevice, s));
    }
    init_kernel<<<(N + 127) / 128, 128, 0, s>>>(N, K, h->prm.h, h->prm.init_div, d_po, h->d_pf, h->d_l[0],
                                                h->d_st[0][0], h->d_st[0][1], h->d_st[0][2]);
    CK(cudaGetLastError());
    if (h->n_static) {
        // ... in BOTH buffers of the ping-pong (no kernel ever writes their rows)
        const size_t oL = (size_t)h->n1 * 3 * K, bL = (size_t)h->n_static * 3 * K * sizeof(double);
        const size_t o3 = 3 * (size_t)h->n1, b3 = 3 * (size_t)h->n_static * sizeof(double);
        CK(cudaMemcpyAsync(h->d_l[1] + oL, h->d_l[0] + oL, bL, cudaMemcpyDeviceToDevice, s));
        for (int q = 1; q < 3; ++q) CK(cudaMemsetAsync(h->d_st[1][q] + o3, 0, b3, s));
        // (d_st[1][0] is the staging copy of po: already the obstacles' positions)
    }
    if (l) CK(cudaMemcpyAsync(l, h->d_l[0], (size_t)N * 3 * K * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (p1) CK(cudaMemcpyAsync(p1, h->d_st[0][0], 3 * (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (v1) CK(cudaMemcpyAsync(v1, h->d_st[0][1], 3 * (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (a1) CK(cudaMemcpyAsync(a1, h->d_st[0][2], 3 * (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    h->have_init = true;
    return 0;
}

int dmpcb200_step_dev(dmpcb200_t* h, const double* d_pk, const double* d_vk, const double* d_ak,
                      const double* d_l_prev, double* d_l_new, double* d_p1, double* d_v1, double* d_a1,
                      double* d_v_hor, double* d_a_hor, int32_t* d_status, dmpcb200_diag* d_diag, void* stream) {
    if (!h || !d_pk || !d_vk || !d_ak || !d_l_prev || !d_l_new || !d_p1 || !d_v1 || !d_a1 || !d_status)
        return fail(DMPCB200_ERR_ARG, "step_dev: null argument");
    if (h->S != 1) return fail(DMPCB200_ERR_STATE, "step_dev: single-scenario entry point on a batched handle (use set_scenario / run_batch / get_scenario)");
    if (!h->have_goals || !h->have_bounds)
        return fail(DMPCB200_ERR_STATE, "step_dev: set_goals and set_bounds first");
    if (((uintptr_t)d_l_prev & 15) != 0) return fail(DMPCB200_ERR_ARG, "step_dev: l_prev must be 16-byte aligned");
    if (int rc = ensure_device(h)) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const bool padded = (d_l_prev == h->d_l[0] || d_l_prev == h->d_l[1]);
    StepArgs A = make_args(h, h->n0, h->n1, d_pk, d_vk, d_ak, d_l_prev, d_l_new, d_p1, d_v1, d_a1, d_v_hor, d_a_hor,
                           d_status, d_diag ? reinterpret_cast<AgentDiag*>(d_diag) : h->d_diag, padded, nullptr);
    if (h->NL == 0) {  // empty block: nothing to solve
        h->launches = 0;
        return 0;
    }
    CK(cudaMemsetAsync(h->d_rescue_next, 0, sizeof(int), s));
    if (use_throughput_layout(h, A)) A.rq = h->d_rq;
    CK(launch_scan(h, A, s));
    CK(launch_qp(h, A, s));
    h->launches = 2;
    return 0;
}

int dmpcb200_goal_dev(dmpcb200_t* h, const double* d_p, int ld, double* d_out, void* stream) {
    if (!h || !d_p || !d_out || ld < 3) return fail(DMPCB200_ERR_ARG, "goal_dev: bad argument");
    if (int rc = ensure_device(h)) return rc;
    TailArgs T = make_tail(h, d_p, ld, nullptr, nullptr, nullptr, nullptr, false, nullptr);
    T.goal_out = d_out;
    T.fail_out = nullptr;
    T.rescue_next = nullptr;
    tail_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(T);
    CK(cudaGetLastError());
    return 0;
}

int dmpcb200_reached_goal(dmpcb200_t* h, const double* p, const double* pf, double tol, double* max_dist,
                          int32_t* pass) {
    if (!h || !p || !pf) return fail(DMPCB200_ERR_ARG, "reached_goal: null argument");
    if (int rc = ensure_device(h)) return rc;
    cudaStream_t s = h->stream;
    const size_t sN = 3 * (size_t)h->N * sizeof(double);
    const int nx = h->cur ^ 1;  // staging: the state buffers that the next step overwrites anyway
    CK(cudaMemcpyAsync(h->d_st[nx][0], p, sN, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(h->d_st[nx][1], pf, sN, cudaMemcpyHostToDevice, s));
    TailArgs T = make_tail(h, h->d_st[nx][0], 3, nullptr, nullptr, nullptr, nullptr, false, nullptr);
    T.pf = h->d_st[nx][1];
    T.goal_tol = tol;
    T.fail_out = nullptr;
    T.rescue_next = nullptr;
    tail_kernel<<<1, 256, 0, s>>>(T);
    CK(cudaGetLastError());
    double out[2] = {0, 0};
    CK(cudaMemcpyAsync(out, h->d_goal, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (max_dist) *max_dist = out[0];
    if (pass) *pass = out[1] != 0.0;
    h->launches = 1;
    return 0;
}

}  // extern "C"

namespace {
// pinned-ness of a host step's arrays: inputs all pinned? device aliases of the four outputs (all or none)
void resolve_pinned(dmpcb200_t* h, dmpcb200_handle::Bound& b) {
    const size_t sN = 3 * (size_t)h->N * sizeof(double), lB = (size_t)h->N * 3 * h->K * sizeof(double);
    b.in_pinned = pinned_alias(b.l_prev, lB) && pinned_alias(b.pk, sN) && pinned_alias(b.vk, sN) && pinned_alias(b.ak, sN);
    b.w_l = static_cast<double*>(pinned_alias(b.l_new, lB));
    b.w_p = b.w_l ? static_cast<double*>(pinned_alias(b.p1, sN)) : nullptr;
    b.w_v = b.w_p ? static_cast<double*>(pinned_alias(b.v1, sN)) : nullptr;
    b.w_a = b.w_v ? static_cast<double*>(pinned_alias(b.a1, sN)) : nullptr;
    if (!b.w_a) b.w_l = b.w_p = b.w_v = nullptr;
}

int step_impl(dmpcb200_t* h, const dmpcb200_handle::Bound& B, int32_t* first_fail) {
    const double *pk = B.pk, *vk = B.vk, *ak = B.ak, *l_prev = B.l_prev;
    double *l_new = B.l_new, *p1 = B.p1, *v1 = B.v1, *a1 = B.a1, *v_hor = B.v_hor, *a_hor = B.a_hor;
    int32_t* status = B.status;
    dmpcb200_diag* diag = B.diag;
    if (!h || !pk || !vk || !ak || !l_prev) return fail(DMPCB200_ERR_ARG, "step: null input");
    if (h->S != 1) return fail(DMPCB200_ERR_STATE, "step: single-scenario entry point on a batched handle (use set_scenario / run_batch / get_scenario)");
    if (!h->have_goals || !h->have_bounds) return fail(DMPCB200_ERR_STATE, "step: set_goals and set_bounds first");
    if (h->NL == 0) {
        if (first_fail) *first_fail = -1;
        return 0;
    }
    if (int rc = ensure_device(h)) return rc;
    const int N = h->N, K = h->K, n3 = 3 * K, n0 = h->n0, NL = h->NL;
    cudaStream_t s = h->stream;
    const size_t sN = 3 * (size_t)N * sizeof(double), lB = (size_t)N * n3 * sizeof(double);
    const int c = 0, nx = 1;
    // inputs: packed into the pinned staging block in the arena's layout, ONE host-to-device copy
    unsigned char* hs_in = h->h_stage;
    unsigned char* hs_out = h->h_stage + h->side_bytes;  // side 1 followed by the tail block, like the arena
    const size_t off_st[3] = {(size_t)((unsigned char*)h->d_st[c][0] - (unsigned char*)h->d_l[c]),
                              (size_t)((unsigned char*)h->d_st[c][1] - (unsigned char*)h->d_l[c]),
                              (size_t)((unsigned char*)h->d_st[c][2] - (unsigned char*)h->d_l[c])};
    const auto tp0 = std::chrono::steady_clock::now();
    const bool in_pinned = B.in_pinned;
    if (!in_pinned) {
        std::memcpy(hs_in, l_prev, lB);
        std::memcpy(hs_in + off_st[0], pk, sN);
        std::memcpy(hs_in + off_st[1], vk, sN);
        std::memcpy(hs_in + off_st[2], ak, sN);
    }
    const auto tp1 = std::chrono::steady_clock::now();
    if (int rc = ensure_events(h, 4)) return rc;
    // outputs: the QP kernel writes the new horizons and states STRAIGHT into the mapped pinned block (posted
    // writes over PCIe as each agent finishes, overlapped with the rest of the kernel); the last CTA adds
    // status | diag | first_fail after the tail.  No device-to-host copy is queued at all.
    unsigned char* ds_out = h->d_stage + h->side_bytes;
    // caller arrays that are pinned are written by the kernel directly (no staging, nothing to hand over)
    double *w_l = B.w_l, *w_p = B.w_p, *w_v = B.w_v, *w_a = B.w_a;
    const bool direct = w_l && w_p && w_v && w_a;
    // (all host-side preparation happens BEFORE the first device call: once the horizons are on their way the
    // scan kernel must already be in the queue -- the GPU used to idle ~10 us behind the copy while the host was
    // still enqueueing the small state copies and building the arguments)
    StepArgs A = make_args(h, h->n0, h->n1, h->d_st[c][0], h->d_st[c][1], h->d_st[c][2], h->d_l[c],
                           direct ? w_l : reinterpret_cast<double*>(ds_out),
                           direct ? w_p : reinterpret_cast<double*>(ds_out + off_st[0]),
                           direct ? w_v : reinterpret_cast<double*>(ds_out + off_st[1]),
                           direct ? w_a : reinterpret_cast<double*>(ds_out + off_st[2]),
                           (v_hor ? h->d_vhor : nullptr), (a_hor ? h->d_ahor : nullptr), h->d_status, h->d_diag, true,
                           nullptr);
    A.T = make_tail(h, nullptr, 3, h->d_status, nullptr, nullptr, nullptr, false, nullptr);
    A.T.copy_src = reinterpret_cast<const unsigned char*>(h->d_status);
    A.T.copy_dst = ds_out + h->side_bytes;
    A.T.copy_bytes = h->tailblk_bytes;
    A.fuse_tail = 1;  // first failing agent + rescue-slot reset in the last CTA of the QP kernel
    if (use_throughput_layout(h, A)) A.rq = h->d_rq;
    if (in_pinned) {
        // pinned caller arrays: DMA straight from them -- the horizons first (the scan kernel needs only those),
        // the scan right behind them, the states on a second stream beside the scan
        CK(cudaMemcpyAsync(h->d_l[c], l_prev, lB, cudaMemcpyHostToDevice, s));
        CK(cudaEventRecord(h->ev[0], s));
        CK(launch_scan(h, A, s));
        CK(cudaMemcpyAsync(h->d_st[c][0], pk, sN, cudaMemcpyHostToDevice, h->stream2));
        CK(cudaMemcpyAsync(h->d_st[c][1], vk, sN, cudaMemcpyHostToDevice, h->stream2));
        CK(cudaMemcpyAsync(h->d_st[c][2], ak, sN, cudaMemcpyHostToDevice, h->stream2));
        CK(cudaEventRecord(h->ev_in, h->stream2));
    } else {
        CK(cudaMemcpyAsync(h->d_l[c], hs_in, h->side_bytes, cudaMemcpyHostToDevice, s));
        CK(cudaEventRecord(h->ev[0], s));
        CK(launch_scan(h, A, s));
    }
    CK(cudaEventRecord(h->ev[1], s));
    if (in_pinned) CK(cudaStreamWaitEvent(s, h->ev_in, 0));  // the states have arrived
    CK(launch_qp(h, A, s));
    CK(cudaEventRecord(h->ev[2], s));
    CK(cudaEventRecord(h->ev[3], s));
    const size_t o3 = 3 * (size_t)n0, b3 = 3 * (size_t)NL * sizeof(double);
    const size_t oL = (size_t)n0 * n3, bL = (size_t)NL * n3 * sizeof(double);
    if (v_hor) CK(cudaMemcpyAsync(v_hor + oL, h->d_vhor + oL, bL, cudaMemcpyDeviceToHost, s));
    if (a_hor) CK(cudaMemcpyAsync(a_hor + oL, h->d_ahor + oL, bL, cudaMemcpyDeviceToHost, s));
    const auto tp2 = std::chrono::steady_clock::now();
    CK(cudaStreamSynchronize(s));
    const auto tp3 = std::chrono::steady_clock::now();
    {
        const unsigned char* tb = hs_out + h->side_bytes;
        const size_t stB = (size_t)((unsigned char*)h->d_diag - (unsigned char*)h->d_status);
        const size_t dgB = (size_t)((unsigned char*)h->d_fail - (unsigned char*)h->d_diag);
        if (!direct) {
            if (l_new) std::memcpy(l_new + oL, reinterpret_cast<const double*>(hs_out) + oL, bL);
            if (p1) std::memcpy(p1 + o3, reinterpret_cast<const double*>(hs_out + off_st[0]) + o3, b3);
            if (v1) std::memcpy(v1 + o3, reinterpret_cast<const double*>(hs_out + off_st[1]) + o3, b3);
            if (a1) std::memcpy(a1 + o3, reinterpret_cast<const double*>(hs_out + off_st[2]) + o3, b3);
        }
        if (status) std::memcpy(status + n0, reinterpret_cast<const int*>(tb) + n0, (size_t)NL * sizeof(int));
        if (diag) std::memcpy(diag + n0, reinterpret_cast<const AgentDiag*>(tb + stB) + n0, (size_t)NL * sizeof(AgentDiag));
        if (first_fail) *first_fail = *reinterpret_cast<const int*>(tb + stB + dgB);
    }
    {
        const auto tp4 = std::chrono::steady_clock::now();
        auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
        h->t_host_us[0] = us(tp0, tp1);
        h->t_host_us[1] = us(tp1, tp2);
        h->t_host_us[2] = us(tp2, tp3);
        h->t_host_us[3] = us(tp3, tp4);
    }
    float a = 0, b = 0, w = 0;
    cudaEventElapsedTime(&a, h->ev[0], h->ev[1]);
    cudaEventElapsedTime(&b, h->ev[1], h->ev[2]);
    cudaEventElapsedTime(&w, h->ev[0], h->ev[3]);
    h->t_ms[0] = a;
    h->t_ms[1] = b;
    h->t_ms[2] = w;
    h->launches = 2;
    h->cur = 0;
    return 0;
}
}  // namespace

extern "C" {

int dmpcb200_step(dmpcb200_t* h, const double* pk, const double* vk, const double* ak, const double* l_prev,
                  double* l_new, double* p1, double* v1, double* a1, double* v_hor, double* a_hor,
                  int32_t* status, dmpcb200_diag* diag, int32_t* first_fail) {
    if (!h || !pk || !vk || !ak || !l_prev) return fail(DMPCB200_ERR_ARG, "step: null input");
    dmpcb200_handle::Bound b = {pk, vk, ak, l_prev, l_new, p1, v1, a1, v_hor, a_hor, status, diag,
                                false, nullptr, nullptr, nullptr, nullptr};
    resolve_pinned(h, b);  // queried afresh on every call: nothing about the caller's arrays is remembered
    return step_impl(h, b, first_fail);
}

int dmpcb200_bind_step(dmpcb200_t* h, const double* pk, const double* vk, const double* ak, const double* l_prev,
                       double* l_new, double* p1, double* v1, double* a1, double* v_hor, double* a_hor,
                       int32_t* status, dmpcb200_diag* diag, int32_t* slot) {
    if (!h || !pk || !vk || !ak || !l_prev || !slot) return fail(DMPCB200_ERR_ARG, "bind_step: null argument");
    if (h->bound.size() >= 64) return fail(DMPCB200_ERR_STATE, "bind_step: too many bindings on this handle (64)");
    dmpcb200_handle::Bound b = {pk, vk, ak, l_prev, l_new, p1, v1, a1, v_hor, a_hor, status, diag,
                                false, nullptr, nullptr, nullptr, nullptr};
    if (int rc = ensure_device(h)) return rc;
    resolve_pinned(h, b);
    h->bound.push_back(b);
    *slot = (int32_t)h->bound.size() - 1;
    return 0;
}

int dmpcb200_step_bound(dmpcb200_t* h, int32_t slot, int32_t* first_fail) {
    if (!h || slot < 0 || (size_t)slot >= h->bound.size()) return fail(DMPCB200_ERR_ARG, "step_bound: bad slot");
    return step_impl(h, h->bound[slot], first_fail);
}

int dmpcb200_run(dmpcb200_t* h, int max_steps, int stop_on_fail, int mode, double* traj_p, double* traj_v,
                 double* traj_a, int32_t* status_hist, int32_t* steps_done, int32_t* reached,
                 int32_t* first_fail_step, int32_t* first_fail_agent) {
    if (!h || max_steps < 1) return fail(DMPCB200_ERR_ARG, "run: bad argument");
    if (h->S != 1) return fail(DMPCB200_ERR_STATE, "run: single-scenario entry point on a batched handle (use set_scenario / run_batch / get_scenario)");
    if (h->n0 != 0 || h->n1 + h->n_static != h->N)
        return fail(DMPCB200_ERR_STATE, "run: needs a handle that owns all agents");
    if (!h->have_init || !h->have_bounds) return fail(DMPCB200_ERR_STATE, "run: set_bounds and init_horizons first");
    if (int rc = ensure_device(h)) return rc;
    const int N = h->N;
    cudaStream_t s = h->stream;
    const bool record = traj_p || traj_v || traj_a || status_hist;
    const bool timed = (mode & 1) != 0;     // per-kernel CUDA events, no graph
    const bool no_graph = (mode & 2) != 0;  // plain launches
    if (record) {
        if (h->traj_S < max_steps || (status_hist && !h->d_hist)) {
            drop_graph(h);
            for (int i = 0; i < 3; ++i) {
                cudaFree(h->d_traj[i]);
                h->d_traj[i] = nullptr;
            }
            cudaFree(h->d_hist);
            h->d_hist = nullptr;
            h->traj_S = max_steps;
            for (int i = 0; i < 3; ++i) CK(dalloc(&h->d_traj[i], 3 * (size_t)(max_steps + 1) * N));
            CK(dalloc(&h->d_hist, (size_t)max_steps * N));
        }
        if (h->hist_on != (status_hist != nullptr)) drop_graph(h);
        h->hist_on = status_hist != nullptr;
        // column 0 = current state
        for (int i = 0; i < 3; ++i)
            CK(cudaMemcpy2DAsync(h->d_traj[i], 3 * (size_t)(h->traj_S + 1) * sizeof(double), h->d_st[h->cur][i],
                                 3 * sizeof(double), 3 * sizeof(double), N, cudaMemcpyDeviceToDevice, s));
    }
    Ctrl c;
    std::memset(&c, 0, sizeof(c));
    c.fail_step = -1;
    c.fail_agent = -1;
    c.stop_on_fail = stop_on_fail;
    c.max_steps = max_steps;
    CK(cudaMemcpyAsync(h->d_ctrl, &c, sizeof(c), cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(h->d_rescue_next, 0, sizeof(int), s));
    if (int rc = ensure_events(h, timed ? 4 * (size_t)max_steps + 2 : 2)) return rc;
    Ctrl hc = c;
    const int start_cur = h->cur;
    int issued = 0;
    if (!timed && !no_graph) {
        if (!h->graph || h->graph_record != record || h->graph_batch) {
            drop_graph(h);
            h->graph_batch = false;
            cudaGraph_t g;
            CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            int rc = launch_resident_step(h, 0, record, h->d_ctrl, s, nullptr);
            if (!rc) rc = launch_resident_step(h, 1, record, h->d_ctrl, s, nullptr);
            cudaError_t e = cudaStreamEndCapture(s, &g);
            if (rc) return rc;
            CK(e);
            CK(cudaGraphInstantiate(&h->graph, g, 0));
            cudaGraphDestroy(g);
            h->graph_record = record;
        }
    }
    CK(cudaEventRecord(h->ev[0], s));
    if (!timed && !no_graph && start_cur == 0) {
        const int check_every = 8;  // graph launches (= 16 steps) between looks at the control word
        int since = 0;
        while (issued < max_steps) {
            CK(cudaGraphLaunch(h->graph, s));
            issued += 2;
            if (++since == check_every && issued < max_steps) {
                since = 0;
                CK(cudaMemcpyAsync(&hc, h->d_ctrl, sizeof(hc), cudaMemcpyDeviceToHost, s));
                CK(cudaStreamSynchronize(s));
                if (hc.done) break;
            }
        }
    } else {
        int cur = start_cur;
        while (issued < max_steps) {
            cudaEvent_t* evs = timed ? &h->ev[2 + 4 * (size_t)issued] : nullptr;
            if (int rc = launch_resident_step(h, cur, record, h->d_ctrl, s, evs)) return rc;
            cur ^= 1;
            ++issued;
            if ((issued & 15) == 0 && issued < max_steps) {
                CK(cudaMemcpyAsync(&hc, h->d_ctrl, sizeof(hc), cudaMemcpyDeviceToHost, s));
                CK(cudaStreamSynchronize(s));
                if (hc.done) break;
            }
        }
    }
    CK(cudaEventRecord(h->ev[1], s));
    CK(cudaMemcpyAsync(&hc, h->d_ctrl, sizeof(hc), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const int steps = hc.step;
    h->cur = start_cur ^ (steps & 1);
    float whole = 0;
    cudaEventElapsedTime(&whole, h->ev[0], h->ev[1]);
    h->t_ms[0] = h->t_ms[1] = 0;
    h->t_ms[2] = steps ? whole / steps : 0;
    if (timed && steps) {
        double a = 0, b = 0, w = 0;
        for (int i = 0; i < steps; ++i) {
            float x;
            cudaEvent_t* evs = &h->ev[2 + 4 * (size_t)i];
            cudaEventElapsedTime(&x, evs[0], evs[1]);
            a += x;
            cudaEventElapsedTime(&x, evs[1], evs[2]);
            b += x;
            cudaEventElapsedTime(&x, evs[0], evs[3]);
            w += x;
        }
        h->t_ms[0] = a / steps;
        h->t_ms[1] = b / steps;
        h->t_ms[2] = w / steps;
    }
    h->launches = 2 * (int64_t)steps;
    if (record) {
        const size_t cols = (size_t)(steps + 1);
        double* outs[3] = {traj_p, traj_v, traj_a};
        for (int i = 0; i < 3; ++i)
            if (outs[i])
                // device 3 x (S+1) x N -> host 3 x (max_steps+1) x N, first steps+1 columns
                CK(cudaMemcpy2DAsync(outs[i], 3 * (size_t)(max_steps + 1) * sizeof(double), h->d_traj[i],
                                     3 * (size_t)(h->traj_S + 1) * sizeof(double), 3 * cols * sizeof(double), N,
                                     cudaMemcpyDeviceToHost, s));
        if (status_hist)
            CK(cudaMemcpyAsync(status_hist, h->d_hist, (size_t)steps * N * sizeof(int), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    }
    if (steps_done) *steps_done = steps;
    if (reached) *reached = hc.reached;
    if (first_fail_step) *first_fail_step = hc.fail_step;
    if (first_fail_agent) *first_fail_agent = hc.fail_agent;
    return 0;
}

/* ---- scenario batching: test/failure_rate.m:61-133 (trial loops) as one launch per kernel ------------- */

int dmpcb200_set_scenario(dmpcb200_t* h, int s, const double* po, const double* pf, const double* pmin,
                          const double* pmax) {
    if (!h || !po || !pf || !pmin || !pmax) return fail(DMPCB200_ERR_ARG, "set_scenario: null argument");
    if (s < 0 || s >= h->S) return fail(DMPCB200_ERR_ARG, "set_scenario: scenario index out of range");
    for (int x = 0; x < 3; ++x)
        if (!(pmin[x] < pmax[x])) return fail(DMPCB200_ERR_ARG, "set_scenario: need pmin < pmax");
    if (int rc = ensure_device(h)) return rc;
    const int N = h->N, K = h->K;
    cudaStream_t st = h->stream;
    const size_t o3 = 3 * (size_t)s * N, b3 = 3 * (size_t)N * sizeof(double);
    double bb[6] = {pmin[0], pmin[1], pmin[2], pmax[0], pmax[1], pmax[2]};
    CK(cudaMemcpyAsync(h->d_bounds + 6 * s, bb, sizeof(bb), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->d_pf + o3, pf, b3, cudaMemcpyHostToDevice, st));
    double* d_po = h->d_st[1][0] + o3;  // staging: the other side's position block of this scenario
    CK(cudaMemcpyAsync(d_po, po, b3, cudaMemcpyHostToDevice, st));
    // initDMPC.m for the scenario's agents into side 0
    init_kernel<<<(N + 127) / 128, 128, 0, st>>>(N, K, h->prm.h, h->prm.init_div, d_po, h->d_pf + o3,
                                                 h->d_l[0] + (size_t)s * h->Npad * 3 * K, h->d_st[0][0] + o3,
                                                 h->d_st[0][1] + o3, h->d_st[0][2] + o3);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    if (h->S == 1) {  // the single-swarm entry points see the same scenario
        for (int x = 0; x < 3; ++x) { h->dp.pmin[x] = pmin[x]; h->dp.pmax[x] = pmax[x]; }
        h->have_bounds = h->have_goals = h->have_init = true;
        drop_graph(h);
    }
    h->per_scen_bounds = true;
    h->scen_set[s] = 1;
    h->cur = 0;
    return 0;
}

int dmpcb200_gen_scenarios(dmpcb200_t* h, uint64_t seed, int mode, double rmin_init, const double* pmin,
                           const double* pmax, double* po_out, double* pf_out) {
    if (!h || !pmin || !pmax) return fail(DMPCB200_ERR_ARG, "gen_scenarios: null argument");
    if (mode != 0 && mode != 1) return fail(DMPCB200_ERR_ARG, "gen_scenarios: mode 0 (randomTest) or 1 (randomExchange)");
    if (!(rmin_init > 0)) return fail(DMPCB200_ERR_ARG, "gen_scenarios: rmin_init must be positive");
    if (h->n0 != 0 || h->n1 + h->n_static != h->N || h->n_static)
        return fail(DMPCB200_ERR_STATE, "gen_scenarios: needs a handle that owns all agents");
    for (int x = 0; x < 3; ++x)
        if (!(pmin[x] < pmax[x])) return fail(DMPCB200_ERR_ARG, "gen_scenarios: need pmin < pmax");
    const int N = h->N, K = h->K, S = h->S;
    if (N > kGenMaxN) return fail(DMPCB200_ERR_ARG, "gen_scenarios: at most 4096 agents per scenario");
    if (int rc = ensure_device(h)) return rc;
    cudaStream_t st = h->stream;
    double* d_po = h->d_st[1][0];  // staging: the other side's position block
    const size_t smem = 3 * (size_t)N * sizeof(double);
    static bool attr_set[64] = {false};
    const int dev = current_device();
    if (smem > 48 * 1024 && !attr_set[dev]) {
        CK(cudaFuncSetAttribute(gen_scenarios_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * kGenMaxN * (int)sizeof(double)));
        attr_set[dev] = true;
    }
    gen_scenarios_kernel<<<S, 256, smem, st>>>(N, mode, (unsigned long long)seed, rmin_init, 1.0 / h->prm.c, pmin[0], pmin[1],
                                              pmin[2], pmax[0], pmax[1], pmax[2], 200000, d_po, h->d_pf);
    CK(cudaGetLastError());
    std::vector<double> bb(6 * (size_t)S);
    for (int s = 0; s < S; ++s)
        for (int x = 0; x < 3; ++x) { bb[6 * s + x] = pmin[x]; bb[6 * s + 3 + x] = pmax[x]; }
    CK(cudaMemcpyAsync(h->d_bounds, bb.data(), bb.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    for (int s = 0; s < S; ++s) {
        const size_t o3 = 3 * (size_t)s * N;
        init_kernel<<<(N + 127) / 128, 128, 0, st>>>(N, K, h->prm.h, h->prm.init_div, d_po + o3, h->d_pf + o3,
                                                     h->d_l[0] + (size_t)s * h->Npad * 3 * K, h->d_st[0][0] + o3,
                                                     h->d_st[0][1] + o3, h->d_st[0][2] + o3);
    }
    CK(cudaGetLastError());
    if (po_out) CK(cudaMemcpyAsync(po_out, d_po, 3 * (size_t)N * S * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (pf_out) CK(cudaMemcpyAsync(pf_out, h->d_pf, 3 * (size_t)N * S * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int x = 0; x < 3; ++x) { h->dp.pmin[x] = pmin[x]; h->dp.pmax[x] = pmax[x]; }
    h->have_bounds = h->have_goals = h->have_init = true;
    h->per_scen_bounds = true;
    std::fill(h->scen_set.begin(), h->scen_set.end(), 1);
    h->cur = 0;
    drop_graph(h);
    return 0;
}

int dmpcb200_run_batch(dmpcb200_t* h, int max_steps, int stop_on_fail, int mode, double* traj_p, double* traj_v,
                       double* traj_a, int32_t* steps_done, int32_t* reached, int32_t* first_fail_step,
                       int32_t* first_fail_agent, double* goal_dist) {
    if (!h || max_steps < 1) return fail(DMPCB200_ERR_ARG, "run_batch: bad argument");
    for (int s = 0; s < h->S; ++s)
        if (!h->scen_set[s]) return fail(DMPCB200_ERR_STATE, "run_batch: dmpcb200_set_scenario every scenario first");
    if (int rc = ensure_device(h)) return rc;
    const int N = h->N, S = h->S;
    cudaStream_t st = h->stream;
    const bool record = traj_p || traj_v || traj_a;
    const bool no_graph = (mode & 3) != 0;
    const bool timed = (mode & 1) != 0;  // per-kernel CUDA events (plain launches)
    if (record) {
        if (h->traj_S < max_steps || h->traj_scen < S) {
            drop_graph(h);
            for (int i = 0; i < 3; ++i) { cudaFree(h->d_traj[i]); h->d_traj[i] = nullptr; }
            h->traj_S = max_steps;
            h->traj_scen = S;
            for (int i = 0; i < 3; ++i) CK(dalloc(&h->d_traj[i], 3 * (size_t)(max_steps + 1) * N * S));
        }
        // column 0 = current state of every agent of every scenario (agents are contiguous: S*N of them)
        for (int i = 0; i < 3; ++i)
            CK(cudaMemcpy2DAsync(h->d_traj[i], 3 * (size_t)(h->traj_S + 1) * sizeof(double), h->d_st[h->cur][i],
                                 3 * sizeof(double), 3 * sizeof(double), (size_t)N * S, cudaMemcpyDeviceToDevice, st));
    }
    std::vector<Ctrl> hc(S);
    for (auto& c : hc) {
        std::memset(&c, 0, sizeof(c));
        c.fail_step = -1;
        c.fail_agent = -1;
        c.stop_on_fail = stop_on_fail;
        c.max_steps = max_steps;
    }
    CK(cudaMemcpyAsync(h->d_ctrl, hc.data(), S * sizeof(Ctrl), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(h->d_rescue_next, 0, sizeof(int), st));
    if (int rc = ensure_events(h, timed ? 4 * (size_t)max_steps + 6 : 2)) return rc;
    int ev_step = 0;
    auto launch_step = [&](int cur) -> int {
        const int nx = cur ^ 1;
        cudaEvent_t* evs = timed ? &h->ev[2 + 4 * (size_t)ev_step++] : nullptr;
        StepArgs A = make_args(h, 0, N, h->d_st[cur][0], h->d_st[cur][1], h->d_st[cur][2], h->d_l[cur], h->d_l[nx],
                               h->d_st[nx][0], h->d_st[nx][1], h->d_st[nx][2], nullptr, nullptr, h->d_status,
                               h->d_diag, true, h->d_ctrl);
        A.n_scen = S;
        A.lstride = (size_t)h->Npad * 3 * h->K;
        A.bounds = h->d_bounds;
        if (evs) CK(cudaEventRecord(evs[0], st));
        if (use_throughput_layout(h, A)) A.rq = h->d_rq;
        CK(launch_scan(h, A, st));
        if (evs) CK(cudaEventRecord(evs[1], st));
        CK(launch_qp(h, A, st));
        if (evs) CK(cudaEventRecord(evs[2], st));
        TailArgs T = make_tail(h, h->d_st[nx][0], 3, h->d_status, h->d_st[nx][0], h->d_st[nx][1], h->d_st[nx][2],
                               record, h->d_ctrl);
        T.n0 = 0;
        T.n1 = N;
        T.status_hist = nullptr;
        tail_batch_kernel<<<S, 128, 0, st>>>(T);
        CK(cudaGetLastError());
        if (evs) CK(cudaEventRecord(evs[3], st));
        return 0;
    };
    const int start_cur = h->cur;
    if (!no_graph && (!h->graph || h->graph_record != record || !h->graph_batch)) {
        drop_graph(h);
        cudaGraph_t g;
        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        int rc = launch_step(0);
        if (!rc) rc = launch_step(1);
        cudaError_t e = cudaStreamEndCapture(st, &g);
        if (rc) return rc;
        CK(e);
        CK(cudaGraphInstantiate(&h->graph, g, 0));
        cudaGraphDestroy(g);
        h->graph_record = record;
        h->graph_batch = true;
    }
    CK(cudaEventRecord(h->ev[0], st));
    int issued = 0, cur = start_cur;
    auto all_done = [&]() -> int {
        CK(cudaMemcpyAsync(hc.data(), h->d_ctrl, S * sizeof(Ctrl), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (const auto& c : hc)
            if (!c.done) return 0;
        return 1;
    };
    while (issued < max_steps) {
        if (!no_graph && cur == 0) {
            CK(cudaGraphLaunch(h->graph, st));
            issued += 2;
        } else {
            if (int rc = launch_step(cur)) return rc;
            cur ^= 1;
            issued += 1;
        }
        if ((issued & 15) == 0 && issued < max_steps) {
            const int d = all_done();
            if (d < 0) return d;
            if (d) break;
        }
    }
    CK(cudaEventRecord(h->ev[1], st));
    {
        const int d = all_done();
        if (d < 0) return d;
    }
    // every scenario that is still running has done `issued` steps (capped by max_steps through its own control
    // word); the buffers have been swapped `issued` times
    h->cur = start_cur ^ (issued & 1);
    float whole = 0;
    cudaEventElapsedTime(&whole, h->ev[0], h->ev[1]);
    long long agent_steps = 0;
    int max_s = 0;
    for (int s = 0; s < S; ++s) {
        agent_steps += (long long)hc[s].step * N;
        max_s = std::max(max_s, hc[s].step);
        if (steps_done) steps_done[s] = hc[s].step;
        if (reached) reached[s] = hc[s].reached;
        if (first_fail_step) first_fail_step[s] = hc[s].fail_step;
        if (first_fail_agent) first_fail_agent[s] = hc[s].fail_agent;
        if (goal_dist) goal_dist[s] = hc[s].goal_dist;
    }
    h->t_ms[0] = h->t_ms[1] = 0;
    h->t_ms[2] = issued ? whole / issued : 0;
    if (timed && ev_step) {
        double a = 0, b = 0;
        for (int i = 0; i < ev_step; ++i) {
            float x;
            cudaEvent_t* evs = &h->ev[2 + 4 * (size_t)i];
            cudaEventElapsedTime(&x, evs[0], evs[1]);
            a += x;
            cudaEventElapsedTime(&x, evs[1], evs[2]);
            b += x;
        }
        h->t_ms[0] = a / ev_step;
        h->t_ms[1] = b / ev_step;
    }
    h->batch_ms = whole;
    h->batch_agent_steps = agent_steps;
    h->launches = 3 * (int64_t)issued;
    if (record) {
        // device: per scenario 3 x (traj_S+1) x N  ->  host: per scenario 3 x (max_steps+1) x N
        const size_t cols = (size_t)std::min(max_s, max_steps) + 1;
        double* outs[3] = {traj_p, traj_v, traj_a};
        for (int i = 0; i < 3; ++i)
            if (outs[i])
                CK(cudaMemcpy2DAsync(outs[i], 3 * (size_t)(max_steps + 1) * sizeof(double), h->d_traj[i],
                                     3 * (size_t)(h->traj_S + 1) * sizeof(double), 3 * cols * sizeof(double),
                                     (size_t)N * S, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    return 0;
}

int dmpcb200_get_scenario(dmpcb200_t* h, int s, double* l, double* pk, double* vk, double* ak, int32_t* status,
                          dmpcb200_diag* diag) {
    if (!h) return fail(DMPCB200_ERR_ARG, "get_scenario: null handle");
    if (s < 0 || s >= h->S) return fail(DMPCB200_ERR_ARG, "get_scenario: scenario index out of range");
    if (int rc = ensure_device(h)) return rc;
    const int N = h->N, n3 = 3 * h->K, c = h->cur;
    cudaStream_t st = h->stream;
    const size_t o3 = 3 * (size_t)s * N, b3 = 3 * (size_t)N * sizeof(double);
    if (l) CK(cudaMemcpyAsync(l, h->d_l[c] + (size_t)s * h->Npad * n3, (size_t)N * n3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (pk) CK(cudaMemcpyAsync(pk, h->d_st[c][0] + o3, b3, cudaMemcpyDeviceToHost, st));
    if (vk) CK(cudaMemcpyAsync(vk, h->d_st[c][1] + o3, b3, cudaMemcpyDeviceToHost, st));
    if (ak) CK(cudaMemcpyAsync(ak, h->d_st[c][2] + o3, b3, cudaMemcpyDeviceToHost, st));
    if (status) CK(cudaMemcpyAsync(status, h->d_status + (size_t)s * N, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (diag) CK(cudaMemcpyAsync(diag, h->d_diag + (size_t)s * N, (size_t)N * sizeof(AgentDiag), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

int dmpcb200_last_batch_timing(dmpcb200_t* h, double* total_ms, int64_t* agent_steps) {
    if (!h) return fail(DMPCB200_ERR_ARG, "last_batch_timing: null handle");
    if (total_ms) *total_ms = h->batch_ms;
    if (agent_steps) *agent_steps = h->batch_agent_steps;
    return 0;
}

int dmpcb200_get_state(dmpcb200_t* h, double* l, double* pk, double* vk, double* ak, int32_t* status,
                       dmpcb200_diag* diag) {
    if (!h) return fail(DMPCB200_ERR_ARG, "get_state: null handle");
    if (h->S != 1) return fail(DMPCB200_ERR_STATE, "get_state: single-scenario entry point on a batched handle (use set_scenario / run_batch / get_scenario)");
    if (int rc = ensure_device(h)) return rc;
    const int N = h->N, n3 = 3 * h->K;
    cudaStream_t s = h->stream;
    const size_t sN = 3 * (size_t)N * sizeof(double);
    if (l) CK(cudaMemcpyAsync(l, h->d_l[h->cur], (size_t)N * n3 * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (pk) CK(cudaMemcpyAsync(pk, h->d_st[h->cur][0], sN, cudaMemcpyDeviceToHost, s));
    if (vk) CK(cudaMemcpyAsync(vk, h->d_st[h->cur][1], sN, cudaMemcpyDeviceToHost, s));
    if (ak) CK(cudaMemcpyAsync(ak, h->d_st[h->cur][2], sN, cudaMemcpyDeviceToHost, s));
    if (status) CK(cudaMemcpyAsync(status, h->d_status, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (diag) CK(cudaMemcpyAsync(diag, h->d_diag, (size_t)N * sizeof(AgentDiag), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return 0;
}

int dmpcb200_set_state(dmpcb200_t* h, const double* l, const double* pk, const double* vk, const double* ak) {
    if (!h || !l || !pk || !vk || !ak) return fail(DMPCB200_ERR_ARG, "set_state: null argument");
    if (h->S != 1) return fail(DMPCB200_ERR_STATE, "set_state: single-scenario entry point on a batched handle (use set_scenario / run_batch / get_scenario)");
    if (!h->have_goals) return fail(DMPCB200_ERR_STATE, "set_state: call dmpcb200_set_goals first");
    if (int rc = ensure_device(h)) return rc;
    const int N = h->N, n3 = 3 * h->K;
    cudaStream_t s = h->stream;
    const size_t sN = 3 * (size_t)N * sizeof(double);
    h->cur = 0;
    CK(cudaMemcpyAsync(h->d_l[0], l, (size_t)N * n3 * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(h->d_st[0][0], pk, sN, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(h->d_st[0][1], vk, sN, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(h->d_st[0][2], ak, sN, cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));
    h->have_init = true;
    return 0;
}

int dmpcb200_solve_agent(dmpcb200_t* h, const double* po, const double* pf, const double* vo, const double* ao,
                         int n, const double* l, double* p, double* v, double* a, int32_t* status,
                         dmpcb200_diag* diag) {
    if (!h || !po || !pf || !vo || !ao || !l) return fail(DMPCB200_ERR_ARG, "solve_agent: null argument");
    if (n < 0 || n >= h->N) return fail(DMPCB200_ERR_ARG, "solve_agent: agent index out of range");
    if (!h->have_bounds) return fail(DMPCB200_ERR_STATE, "solve_agent: call dmpcb200_set_bounds first");
    if (int rc = ensure_device(h)) return rc;
    const int N = h->N, K = h->K, n3 = 3 * K;
    cudaStream_t s = h->stream;
    const size_t b3 = 3 * sizeof(double);
    // a batch of one on the handle's drop-in scratch: the resident loop state (l, pk, vk, ak, goals) of
    // dmpcb200_run / dmpcb200_step is not touched
    Scratch S;
    if (int rc = get_scratch(h, &S)) return rc;
    CK(cudaMemcpyAsync(S.st_in[0] + 3 * n, po, b3, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(S.st_in[1] + 3 * n, vo, b3, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(S.st_in[2] + 3 * n, ao, b3, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(S.pf + 3 * n, pf, b3, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(S.l_in, l, (size_t)N * n3 * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(h->d_rescue_next, 0, sizeof(int), s));
    // scratch slot 0 of the row buffers (local index = n - n0 with n0 = n)
    StepArgs A = make_args(h, n, n + 1, S.st_in[0], S.st_in[1], S.st_in[2], S.l_in, S.l_out, S.st_out[0],
                           S.st_out[1], S.st_out[2], h->d_vhor, h->d_ahor, S.status, S.diag, true, nullptr);
    A.pf = S.pf;
    CK(launch_scan(h, A, s));
    CK(launch_qp(h, A, s));
    int st = 0;
    CK(cudaMemcpyAsync(&st, S.status + n, sizeof(int), cudaMemcpyDeviceToHost, s));
    if (p) CK(cudaMemcpyAsync(p, S.l_out + (size_t)n * n3, n3 * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (v) CK(cudaMemcpyAsync(v, h->d_vhor + (size_t)n * n3, n3 * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (a) CK(cudaMemcpyAsync(a, h->d_ahor + (size_t)n * n3, n3 * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (diag) CK(cudaMemcpyAsync(diag, S.diag + n, sizeof(AgentDiag), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (status) *status = st;
    h->launches = 2;
    return 0;
}

int dmpcb200_check_coll(dmpcb200_t* h, const double* p3, const double* l, int n, int k, uint8_t* violation,
                        uint8_t* viol_constr, double* min_dist, int32_t* any_violation) {
    if (!h || !p3 || !l) return fail(DMPCB200_ERR_ARG, "check_coll: null argument");
    if (k < 1 || k > h->K || n < 0 || n >= h->N) return fail(DMPCB200_ERR_ARG, "check_coll: k or n out of range");
    if (int rc = ensure_device(h)) return rc;
    const int N = h->N, n3 = 3 * h->K;
    cudaStream_t s = h->stream;
    Scratch S;
    if (int rc = get_scratch(h, &S)) return rc;
    CK(cudaMemcpyAsync(S.l_in, l, (size_t)N * n3 * sizeof(double), cudaMemcpyHostToDevice, s));
    check_coll_kernel<<<1, 256, 0, s>>>(h->dp, p3[0], p3[1], p3[2], S.l_in, n, k, h->d_u8, h->d_u8 + N,
                                        h->d_small, h->d_ismall);
    CK(cudaGetLastError());
    double md = 0;
    int any = 0;
    if (violation) CK(cudaMemcpyAsync(violation, h->d_u8, N, cudaMemcpyDeviceToHost, s));
    if (viol_constr) CK(cudaMemcpyAsync(viol_constr, h->d_u8 + N, N, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&md, h->d_small, sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&any, h->d_ismall, sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (min_dist) *min_dist = md;
    if (any_violation) *any_violation = any;
    h->launches = 1;
    return 0;
}

int dmpcb200_coll_constr(dmpcb200_t* h, const double* p3, const double* po, const double* vo, int n, int k,
                         const double* l, const uint8_t* mask, int cap, double* Ain, double* bin,
                         double* prev_dist, int32_t* nrows) {
    if (!h || !p3 || !po || !vo || !l || !nrows) return fail(DMPCB200_ERR_ARG, "coll_constr: null argument");
    // n = -1: no own agent (dec-iSCP/CollConstr.m:1-23, where l holds the obstacles only)
    if (k < 1 || k > h->K || n < -1 || n >= h->N || cap < 1)
        return fail(DMPCB200_ERR_ARG, "coll_constr: k, n or cap out of range");
    if (h->prm.variant != DMPCB200_HARD && !mask)
        return fail(DMPCB200_ERR_ARG, "coll_constr: mask required for this variant");
    if (h->prm.variant == DMPCB200_SOFT_BOUND2 && k < 2)
        return fail(DMPCB200_ERR_ARG, "coll_constr: bound2 has no constraint step for k = 1");
    if (int rc = ensure_device(h)) return rc;
    const int N = h->N, n3 = 3 * h->K;
    cudaStream_t s = h->stream;
    double *d_A = nullptr, *d_b = nullptr;
    Scratch S;
    if (int rc = get_scratch(h, &S)) return rc;
    CK(cudaMemcpyAsync(S.l_in, l, (size_t)N * n3 * sizeof(double), cudaMemcpyHostToDevice, s));
    if (mask) CK(cudaMemcpyAsync(h->d_u8, mask, N, cudaMemcpyHostToDevice, s));
    CK(cudaMallocAsync((void**)&d_A, (size_t)cap * n3 * sizeof(double), s));
    CK(cudaMallocAsync((void**)&d_b, 2 * (size_t)cap * sizeof(double), s));
    CK(cudaMemsetAsync(d_A, 0, (size_t)cap * n3 * sizeof(double), s));
    coll_constr_kernel<<<1, 32, 0, s>>>(h->dp, h->d_tab, p3[0], p3[1], p3[2], po[0], po[1], po[2], vo[0], vo[1],
                                        vo[2], n, k, S.l_in, h->d_u8, cap, d_A, d_b, d_b + cap, h->d_ismall);
    CK(cudaGetLastError());
    int nr = 0;
    CK(cudaMemcpyAsync(&nr, h->d_ismall, sizeof(int), cudaMemcpyDeviceToHost, s));
    if (Ain) CK(cudaMemcpyAsync(Ain, d_A, (size_t)cap * n3 * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (bin) CK(cudaMemcpyAsync(bin, d_b, (size_t)cap * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (prev_dist) CK(cudaMemcpyAsync(prev_dist, d_b + cap, (size_t)cap * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaFreeAsync(d_A, s));
    CK(cudaFreeAsync(d_b, s));
    CK(cudaStreamSynchronize(s));
    *nrows = nr;
    h->launches = 1;
    if (nr > cap) return fail(DMPCB200_ERR_ARG, "coll_constr: more rows than cap");
    return 0;
}

int dmpcb200_prop_state(dmpcb200_t* h, int B, const double* po, const double* vo, const double* a, double* p,
                        double* v) {
    if (!h || B < 1 || !po || !vo || !a) return fail(DMPCB200_ERR_ARG, "prop_state: bad argument");
    if (int rc = ensure_device(h)) return rc;
    const int n3 = 3 * h->K;
    cudaStream_t s = h->stream;
    double* d = nullptr;
    const size_t nb = (size_t)B * n3;
    CK(cudaMallocAsync((void**)&d, (3 * nb + 6 * (size_t)B) * sizeof(double), s));
    double *d_a = d, *d_p = d + nb, *d_v = d + 2 * nb, *d_po = d + 3 * nb, *d_vo = d_po + 3 * (size_t)B;
    CK(cudaMemcpyAsync(d_a, a, nb * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_po, po, 3 * (size_t)B * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_vo, vo, 3 * (size_t)B * sizeof(double), cudaMemcpyHostToDevice, s));
    prop_state_kernel<<<(int)((nb + 127) / 128), 128, 0, s>>>(B, h->K, h->d_tab, h->prm.h, d_po, d_vo, d_a, d_p, d_v);
    CK(cudaGetLastError());
    if (p) CK(cudaMemcpyAsync(p, d_p, nb * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (v) CK(cudaMemcpyAsync(v, d_v, nb * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaFreeAsync(d, s));
    CK(cudaStreamSynchronize(s));
    h->launches = 1;
    return 0;
}

}  // extern "C"

namespace {
int postprocess_impl(dmpcb200_t* h, const double* d_goals, int S, double* pk, double* vk, double* ak, double vmax,
                     double amax, double Ts, double goal_radius, double* p, double* v, double* a, int nt_cap,
                     int32_t* time_index, dmpcb200_post* res) {
    if (!h || !pk || !vk || !ak || !res) return fail(DMPCB200_ERR_ARG, "postprocess: null argument");
    if (S < 4) return fail(DMPCB200_ERR_ARG, "postprocess: needs at least 4 trajectory columns (not-a-knot spline)");
    if (!(vmax > 0) || !(amax > 0) || !(Ts > 0)) return fail(DMPCB200_ERR_ARG, "postprocess: vmax, amax, Ts must be positive");
    if (!h->have_goals && !h->per_scen_bounds) return fail(DMPCB200_ERR_STATE, "postprocess: set_goals first");
    if (int rc = ensure_device(h)) return rc;
    const int N = h->N;
    cudaStream_t s = h->stream;
    const size_t nS = (size_t)N * S * 3, bS = nS * sizeof(double);
    const bool want_all = p && v && a;
    const int nq = want_all ? 3 : 1;  // positions only unless the caller wants v and a as well
    double *d_tr = nullptr, *d_sl = nullptr, *d_out = nullptr, *d_dist = nullptr;
    int* d_tidx = nullptr;
    unsigned long long* d_bits = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_tr); cudaFree(d_sl); cudaFree(d_out); cudaFree(d_dist); cudaFree(d_tidx); cudaFree(d_bits);
    };
    auto bail = [&](cudaError_t e, const char* what) {
        cleanup();
        return fail(DMPCB200_ERR_CUDA, std::string("postprocess: ") + what + ": " + cudaGetErrorString(e));
    };
    cudaError_t e;
    if ((e = cudaMalloc((void**)&d_tr, 3 * bS)) != cudaSuccess) return bail(e, "trajectory");
    if ((e = cudaMalloc((void**)&d_sl, 2 * (size_t)nq * nS * sizeof(double))) != cudaSuccess) return bail(e, "slopes");
    if ((e = cudaMalloc((void**)&d_dist, (size_t)N * sizeof(double))) != cudaSuccess) return bail(e, "dist");
    if ((e = cudaMalloc((void**)&d_tidx, (size_t)N * sizeof(int))) != cudaSuccess) return bail(e, "time index");
    if ((e = cudaMalloc((void**)&d_bits, 2 * sizeof(unsigned long long))) != cudaSuccess) return bail(e, "reductions");
    double *d_pk = d_tr, *d_vk = d_tr + nS, *d_ak = d_tr + 2 * nS;
    if (int rc = ensure_events(h, 4)) { cleanup(); return rc; }
    if ((e = cudaMemcpyAsync(d_pk, pk, bS, cudaMemcpyHostToDevice, s)) != cudaSuccess) return bail(e, "copy");
    if ((e = cudaMemcpyAsync(d_vk, vk, bS, cudaMemcpyHostToDevice, s)) != cudaSuccess) return bail(e, "copy");
    if ((e = cudaMemcpyAsync(d_ak, ak, bS, cudaMemcpyHostToDevice, s)) != cudaSuccess) return bail(e, "copy");
    // the two atomicMin reductions run on the bit patterns of non-negative doubles: seed with +inf, so that a
    // trajectory without motion (all v = a = 0) leaves r_factor = inf and is reported as such
    static const unsigned long long kInfBits[2] = {0x7ff0000000000000ull, 0x7ff0000000000000ull};
    if ((e = cudaMemcpyAsync(d_bits, kInfBits, sizeof(kInfBits), cudaMemcpyHostToDevice, s)) != cudaSuccess) return bail(e, "seed");
    cudaEventRecord(h->ev[0], s);
    // 1. r_factor  (failure_rate.m:141-144)
    pp_rfactor_kernel<<<std::min((N * S + 255) / 256, 592), 256, 0, s>>>(N * S, d_vk, d_ak, vmax, amax, d_bits);
    unsigned long long bits = 0;
    if ((e = cudaMemcpyAsync(&bits, d_bits, sizeof(bits), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return bail(e, "copy");
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return bail(e, "r_factor");
    double r_factor;
    std::memcpy(&r_factor, &bits, sizeof(double));
    if (!(r_factor > 0) || !std::isfinite(r_factor)) {
        cleanup();
        return fail(DMPCB200_ERR_ARG, "postprocess: the trajectory has no motion (r_factor is not finite)");
    }
    const double hs = h->prm.h / std::sqrt(r_factor);  // :145
    const double T = (double)(S - 1) * hs;             // :148
    const int nt = (int)std::floor(T / Ts + 1e-9) + 1; // length(0:Ts:T)
    // 2. time scaling (:156-162)
    pp_rescale_kernel<<<(3 * N + 127) / 128, 128, 0, s>>>(N, S, r_factor, hs, d_pk, d_vk, d_ak);
    // 3. spline slopes and 100 Hz evaluation (:164-168)
    if ((e = cudaMalloc((void**)&d_out, (size_t)nq * N * nt * 3 * sizeof(double))) != cudaSuccess) return bail(e, "interpolation");
    double* d_p = d_out;
    double* d_v = want_all ? d_out + (size_t)N * nt * 3 : nullptr;
    double* d_a = want_all ? d_out + 2 * (size_t)N * nt * 3 : nullptr;
    pp_slopes_kernel<<<(nq * 3 * N + 63) / 64, 64, 0, s>>>(N, S, nq, hs, d_pk, d_vk, d_ak, d_sl, d_sl + (size_t)nq * nS);
    {
        const long long tot = (long long)nq * N * nt;
        pp_eval_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(N, S, nt, nq, hs, Ts, d_pk, d_vk, d_ak, d_sl, d_p, d_v, d_a);
    }
    // 4. pairwise check (:170-181), distance and trajectory time (:183-194)
    {
        const int tiles = (N + kPairTile - 1) / kPairTile;
        int ex = 0;
        const bool pow2 = std::frexp(h->prm.c, &ex) == 0.5;  // dz / c is then an exact scaling
        if (pow2)
            pp_pairs_kernel<true><<<dim3(tiles, tiles), dim3(kPairTile, kPairTile), 0, s>>>(N, nt, h->prm.c, 1.0 / h->prm.c,
                                                                                            d_p, d_bits + 1);
        else
            pp_pairs_kernel<false><<<dim3(tiles, tiles), dim3(kPairTile, kPairTile), 0, s>>>(N, nt, h->prm.c, 1.0 / h->prm.c,
                                                                                             d_p, d_bits + 1);
        pp_stats_kernel<<<(N + 3) / 4, 128, 0, s>>>(N, nt, goal_radius, d_p, d_goals, d_dist, d_tidx);
    }
    cudaEventRecord(h->ev[1], s);
    if ((e = cudaGetLastError()) != cudaSuccess) return bail(e, "launch");
    std::vector<double> dist(N);
    std::vector<int> tidx(N);
    if ((e = cudaMemcpyAsync(pk, d_pk, bS, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return bail(e, "copy");
    if ((e = cudaMemcpyAsync(vk, d_vk, bS, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return bail(e, "copy");
    if ((e = cudaMemcpyAsync(ak, d_ak, bS, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return bail(e, "copy");
    if ((e = cudaMemcpyAsync(dist.data(), d_dist, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return bail(e, "copy");
    if ((e = cudaMemcpyAsync(tidx.data(), d_tidx, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return bail(e, "copy");
    if ((e = cudaMemcpyAsync(&bits, d_bits + 1, sizeof(bits), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return bail(e, "copy");
    const size_t bO = (size_t)N * nt * 3 * sizeof(double);
    if (nt <= nt_cap) {
        if (p && (e = cudaMemcpyAsync(p, d_p, bO, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return bail(e, "copy");
        if (want_all && (e = cudaMemcpyAsync(v, d_v, bO, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return bail(e, "copy");
        if (want_all && (e = cudaMemcpyAsync(a, d_a, bO, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return bail(e, "copy");
    }
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return bail(e, "sync");
    double min_sq;
    std::memcpy(&min_sq, &bits, sizeof(double));
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
    res->r_factor = r_factor;
    res->h_scaled = hs;
    res->T = T;
    res->nt = nt;
    res->min_dist = N > 1 ? std::sqrt(min_sq) : INFINITY;
    res->violation = res->min_dist < h->prm.rmin - h->prm.coll_tol ? 1 : 0;
    double tot = 0.0;
    int tmax = 0;
    for (int n = 0; n < N; ++n) {
        tot += dist[n];
        tmax = std::max(tmax, tidx[n]);
        if (time_index) time_index[n] = tidx[n];
    }
    res->totdist = tot;
    res->traj_time = tmax * Ts;
    res->device_ms = ms;
    h->launches = 6;
    cleanup();
    return 0;
}
}  // namespace

extern "C" {

int dmpcb200_postprocess(dmpcb200_t* h, int S, double* pk, double* vk, double* ak, double vmax, double amax,
                         double Ts, double goal_radius, double* p, double* v, double* a, int nt_cap,
                         int32_t* time_index, dmpcb200_post* res) {
    if (!h) return fail(DMPCB200_ERR_ARG, "postprocess: null handle");
    return postprocess_impl(h, h->d_pf, S, pk, vk, ak, vmax, amax, Ts, goal_radius, p, v, a, nt_cap, time_index, res);
}

int dmpcb200_postprocess_scenario(dmpcb200_t* h, int scen, int S, double* pk, double* vk, double* ak, double vmax,
                                  double amax, double Ts, double goal_radius, double* p, double* v, double* a,
                                  int nt_cap, int32_t* time_index, dmpcb200_post* res) {
    if (!h) return fail(DMPCB200_ERR_ARG, "postprocess_scenario: null handle");
    if (scen < 0 || scen >= h->S) return fail(DMPCB200_ERR_ARG, "postprocess_scenario: scenario index out of range");
    return postprocess_impl(h, h->d_pf + 3 * (size_t)scen * h->N, S, pk, vk, ak, vmax, amax, Ts, goal_radius, p, v, a,
                            nt_cap, time_index, res);
}

/* host-side phases of the last dmpcb200_step in microseconds: pack, submit, wait, unpack */
int dmpcb200_last_host_timing(dmpcb200_t* h, double* us4) {
    if (!h || !us4) return fail(DMPCB200_ERR_ARG, "last_host_timing: null argument");
    for (int i = 0; i < 4; ++i) us4[i] = h->t_host_us[i];
    return 0;
}

int dmpcb200_last_timing(dmpcb200_t* h, double* ms, int64_t* launches) {
    if (!h) return fail(DMPCB200_ERR_ARG, "last_timing: null handle");
    if (ms)
        for (int i = 0; i < 3; ++i) ms[i] = h->t_ms[i];
    if (launches) *launches = h->launches;
    return 0;
}

void* dmpcb200_device_ptr(dmpcb200_t* h, int which) {
    if (!h) return nullptr;
    switch (which) {
        case 0: return h->d_l[h->cur];
        case 1: return h->d_l[h->cur ^ 1];
        case 2: return h->d_st[h->cur][0];
        case 3: return h->d_st[h->cur][1];
        case 4: return h->d_st[h->cur][2];
        case 5: return h->d_pf;
        case 6: return h->d_status;
        case 7: return h->d_goal;
        case 8: return h->d_st[h->cur ^ 1][0];
        case 9: return h->d_st[h->cur ^ 1][1];
        case 10: return h->d_st[h->cur ^ 1][2];
        case 11: return h->d_diag;
        default: return nullptr;
    }
}

int dmpcb200_swap_horizons(dmpcb200_t* h) {
    if (!h) return fail(DMPCB200_ERR_ARG, "swap_horizons: null handle");
    h->cur ^= 1;
    return 0;
}

/* profiling builds (-DDMPC_PROF) only: cycles[0..15], counts[16..31] of the QP phases; reset after read */
int dmpcb200_prof_read(uint64_t* out32) {
#if defined(DMPC_PROF) || defined(DMPC_PROF_SCAN)
    unsigned long long z[32] = {0};
    if (cudaMemcpyFromSymbol(out32, g_prof, sizeof(z)) != cudaSuccess) return DMPCB200_ERR_CUDA;
    if (cudaMemcpyToSymbol(g_prof, z, sizeof(z)) != cudaSuccess) return DMPCB200_ERR_CUDA;
    return 0;
#else
    (void)out32;
    return DMPCB200_ERR_STATE;
#endif
}

int dmpcb200_config(dmpcb200_t* h, int32_t* out8) {
    if (!h || !out8) return fail(DMPCB200_ERR_ARG, "config: null argument");
    out8[0] = h->W;
    out8[1] = h->QMAX;
    out8[2] = h->RCAP;
    out8[3] = h->RMAX;
    out8[4] = h->QBIG;
    out8[5] = h->n_rescue;
    out8[6] = (int32_t)qp_smem_bytes(h->K, h->W, h->QMAX, h->RCAP);
    {
        const int w = (h->NL <= 4 * 148) ? 4 : 8;
        out8[7] = (int32_t)scan_smem_bytes(h->K, w, std::max(scan_stages(h->K, h->N, w, h->RMAX), 0), h->Npad, h->RMAX);
    }
    return 0;
}

}  // extern "C"

// dmpc_kernels.cuh -- the sm_100a kernels of the DMPC step.
//
//   scan_kernel   (K1)  O(N^2 K) neighbour scan + collision-row build.  One warp per agent, W agents
//                       per CTA; the neighbour-prediction buffer l (3 x K x N fp64, agent-major)
//                       is streamed through shared memory in tiles of 32 agents by TMA bulk copies
//                       (cp.async.bulk + mbarrier, double buffered) and every tile is shared by
//                       the CTA's W agents.  Lane m of a warp owns neighbour m of the tile.
//   qp_kernel     (K2)  batched per-agent QP.  One warp per agent; the constant tables of all
//                       three weight sets are staged once per CTA by one TMA bulk copy; the
//                       agent's rows, multipliers and Schur inverse live in shared memory
//                       (qp_core.cuh).  Agents whose active set outgrows the on-chip capacity
//                       re-solve in a global-memory rescue slot.
//   tail_kernel   (K3)  ReachedGoal.m reduction, first failing agent, trajectory record, loop
//                       control word (so the closed loop needs no host synchronisation).
//   small kernels       initDMPC.m for all agents, the per-agent helper drop-ins.
//
// Reference: test/failure_rate.m:99-127 (loop body), solveSoftDMPCbound.m, CheckCollSoftDMPC.m,
// CollConstrSoftDMPC.m, ReachedGoal.m, initDMPC.m, propStatedmpc.m (dmpc/matlab).
#pragma once
#include <cuda_runtime.h>

#include "scan_core.cuh"
#include "qp_warp.cuh"

namespace dmpc {

constexpr int kTile = 32;  // neighbours per shared-memory tile (= one per lane)
// throughput layout of K2 (qp2_kernel, below): agents per SM in the light / heavy phases, and the size of last
// step's active set from which the scan kernel queues an agent as heavy
constexpr int kLightW = 8, kHeavyW = 2, kHeavyMax = 4;
constexpr int kRouteNact = 28;

struct ScanRec {
    int kstar, nv, flag, pad;
};

// loop control word (device resident)
struct Ctrl {
    int step;        // MPC steps completed so far
    int done;        // set when the loop must stop: kernels become no-ops
    int reached;     // ReachedGoal
    int fail_step;   // first step at which some agent failed, -1
    int fail_agent;  // lowest failing agent of that step, -1
    int stop_on_fail;
    int max_steps;
    int rescue_used;  // global rescue slots handed out (monotone, diagnostics)
    double goal_dist;
    int infeasible;   // stop_on_fail = 2: the loop ended on an infeasible QP
    int pad;
};

struct TailArgs {
    int N, n0, n1, ld;  // p has leading dimension ld (3: packed, 3K: first column of a horizon)
    double goal_tol;
    const double* p;
    const double* pf;
    const int* status;
    const double *p1, *v1, *a1;        // recorded into the trajectory (3 x N)
    double *traj_p, *traj_v, *traj_a;  // optional 3 x (S+1) x N
    int* status_hist;                  // optional S x N
    int S;
    double* goal_out;  // [0] max distance, [1] reached (0/1)
    int* fail_out;     // first failing agent or -1
    int* rescue_next;
    Ctrl* ctrl;  // optional
    // optional: after the tail, copy a block (status | diag | first_fail) to mapped host memory
    const unsigned char* copy_src;
    unsigned char* copy_dst;
    size_t copy_bytes;  // multiple of 16
};

// route queues of the throughput layout of K2 (qp2_kernel): filled by the scan kernel (and by light agents that
// outgrow their capacity), consumed and reset by the QP kernel
struct RouteQ {
    unsigned n_light, n_heavy;        // entries pushed (n_heavy still grows while the QP kernel runs)
    unsigned head_light, head_heavy;  // next entry to take
    unsigned light_done;              // CTAs that have finished their light phase
    unsigned done_ctas;               // CTAs that have finished (the last one runs the tail and resets the queues)
};

struct StepArgs {
    DevParams P;
    ScanThr thr;  // exact squared distance thresholds (scan_core.cuh)
    int n0, n1;  // agents solved by this launch (of every scenario)
    // scenario batching (test/failure_rate.m trial loops as ONE launch): n_scen independent swarms of P.N agents.
    // Scenario s keeps its horizons at l + s * lstride (doubles; a whole number of 32-agent tiles) and the
    // per-agent arrays (states, goals, status, first columns, v/a horizons) at agent index s * P.N + n.
    // n_scen = 1, lstride = 0 is the single-swarm layout.
    int n_scen;
    size_t lstride;
    const double* bounds;  // optional: per scenario pmin[3], pmax[3] (null: P.pmin / P.pmax for all)
    int RMAX, QMAX, RCAP, QBIG, n_rescue;
    int tile_padded;  // l_prev is readable up to a multiple of kTile agents
    const double* l_prev;
    double* l_new;
    const double *pk, *vk, *ak, *pf;
    double *p1, *v1, *a1;
    double *v_hor, *a_hor;  // optional
    int* status;
    AgentDiag* diag;  // optional
    const double* tab;
    ScanRec* scan;  // per local agent
    double* grow;   // per local agent 5*RMAX
    int* gkc;       // per local agent RMAX
    int* gidx;      // per local agent RMAX (neighbour index of each row)
    double* gscr_d;  // per local agent 4*RMAX
    int* gscr_i;     // per local agent 4*RMAX
    unsigned char* rescue;  // n_rescue slots of rescue_bytes
    size_t rescue_bytes;
    int* rescue_next;  // slot allocator (reset by the tail kernel)
    Ctrl* ctrl;        // optional
    // K3 fused into K2: the last CTA of the QP kernel to finish runs the step's tail (optional)
    int fuse_tail;
    unsigned* done_cnt;  // CTAs that finished (reset by the last one)
    unsigned* work_cnt;  // agent queue of the persistent QP grid (reset by the last CTA)
    RouteQ* rq;          // throughput layout (null: classic layout): route queues
    int *q_light, *q_heavy;  // agent indices; q_heavy entries are -1 when empty (self-cleaning)
    TailArgs T;
};

// ---- PTX helpers: mbarrier + TMA bulk copy ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
    } while (!ok);
}

DMPC_HD size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- K1 -------------------------------------------------------------------------------------
constexpr int kScanMaxStages = 24;
// ring depth: as many 32-agent tiles as fit in ~200 KB of shared memory (N = 500, K = 15: the whole
// neighbour buffer, 16 tiles, is in flight at once; the TMA round trip is paid once, not per tile)
// shared memory of K1 besides the tile ring: own horizons, barriers, partial results, and per agent the
// near masks of all neighbours (Npad words) and the compaction list of scan_finish (RMAX words)
DMPC_HD size_t scan_fixed_bytes(int K, int W, int Npad, int RMAX) {
    return (size_t)W * round_up(3 * K, 2) * sizeof(double) + kScanMaxStages * sizeof(uint64_t) +
           64 * 2 * sizeof(unsigned) + (size_t)W * ((size_t)Npad + RMAX) * sizeof(unsigned);
}
DMPC_HD int scan_stages(int K, int N, int W, int RMAX) {
    const size_t tile_bytes = (size_t)kTile * 3 * K * sizeof(double);
    const int ntiles = (N + kTile - 1) / kTile;
    const size_t fixed = scan_fixed_bytes(K, W, ntiles * kTile, RMAX);
    const size_t budget = 224u * 1024u;
    int s = fixed + tile_bytes <= budget ? (int)((budget - fixed) / tile_bytes) : 0;
    if (s > ntiles) s = ntiles;
    if (s > kScanMaxStages) s = kScanMaxStages;
    return s;  // 0: does not fit (the caller picks a smaller W)
}
DMPC_HD size_t scan_smem_bytes(int K, int W, int stages, int Npad, int RMAX) {
    return (size_t)stages * kTile * 3 * K * sizeof(double) + scan_fixed_bytes(K, W, Npad, RMAX);
}

#if defined(DMPC_PROF_SCAN)
#define SCAN_PROF(i)                                                                           \
    do {                                                                                       \
        if ((threadIdx.x & 31) == 0) {                                                         \
            const unsigned long long d = (unsigned long long)(clock64() - scan_t0);            \
            atomicMax(&g_prof[i], d);                                                          \
            atomicAdd(&g_prof[8 + (i)], d);                                                    \
            atomicAdd(&g_prof[16 + (i)], 1ull);                                                \
        }                                                                                      \
    } while (0)
#else
#define SCAN_PROF(i)
#endif

// OWNREG: the agent's own horizon lives in registers (needs KT > 0 and ~2 x 3KT registers: layouts of up to 8
// warps); false: it is read from shared memory (broadcast reads) -- the 16-warp layouts, where occupancy beats it
template <int W, int S, int KT, bool OWNREG = (KT > 0)>
__global__ void __launch_bounds__(W * S * 32) scan_kernel(const __grid_constant__ StepArgs A, int stages) {
    // programmatic dependent launch: the QP kernel of this step may start its prologue (barrier, table
    // TMA) while this grid is still running; it waits (griddepcontrol.wait) before it reads our output
    asm volatile("griddepcontrol.launch_dependents;");
    // CTA -> (scenario, block of W agents of it): the W agents of a CTA share the scenario's tiles
    const int nl_s = A.n1 - A.n0;
    const int cps = (nl_s + W - 1) / W;
    const int scen = (A.n_scen > 1) ? (int)blockIdx.x / cps : 0;
    const int blk = (int)blockIdx.x - scen * cps;
    const double* const l_prev = A.l_prev + (size_t)scen * A.lstride;
#if defined(DMPC_PROF_SCAN)
    const long long scan_t0 = clock64();
    long long scan_wait = 0;
#endif
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int K = KT ? KT : A.P.K, n3 = 3 * K, n3p = round_up(n3, 2), N = A.P.N;
    const int tile_d = kTile * n3;
    const uint32_t tile_bytes = (uint32_t)(tile_d * sizeof(double));  // 32*3K*8: multiple of 16
    double* tiles = reinterpret_cast<double*>(smem_raw);
    double* own_all = tiles + (size_t)stages * tile_d;
    uint64_t* bars = reinterpret_cast<uint64_t*>(own_all + W * n3p);
    unsigned* s_acc = reinterpret_cast<unsigned*>(bars + kScanMaxStages);
    unsigned* nm_all = s_acc + 64 * 2;
    const int Npad = round_up(N, kTile);
    int* list_all = reinterpret_cast<int*>(nm_all + (size_t)W * Npad);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ag = warp / S, sub = warp - ag * S;
    const int li_s = blk * W + ag;                  // local index within the scenario's block of agents
    const int n = A.n0 + li_s;
    const bool valid = n < A.n1;
    const int li = scen * nl_s + li_s;              // index into the per-agent scratch (scan records, rows)
    if (A.ctrl && A.ctrl[scen].done) {
        // the scenario's loop has ended: its agents only have their state carried over by the QP kernel
        // (a single swarm whose loop has ended: the QP kernel returns at once and takes nothing from a queue)
        if (A.rq && A.n_scen > 1 && valid && sub == 0 && lane == 0) A.q_light[atomicAdd(&A.rq->n_light, 1u)] = li;
        return;
    }
    double* own = own_all + ag * n3p;

    const int ntma = A.tile_padded ? (N + kTile - 1) / kTile : N / kTile;
    // every CTA walks the tiles in its own rotation: the CTAs do not all pull the same L2 lines at once
    const int rot = ntma ? (int)((unsigned)blk % (unsigned)ntma) : 0;
    if (threadIdx.x == 0) {
        for (int b = 0; b < stages; ++b) mbar_init(&bars[b], 1);
        mbar_fence_init();
        for (int t = 0; t < stages && t < ntma; ++t) {
            int tau = t + rot;
            tau -= (tau >= ntma) ? ntma : 0;
            mbar_expect_tx(&bars[t], tile_bytes);
            tma_bulk_g2s(tiles + (size_t)t * tile_d, l_prev + (size_t)tau * tile_d, tile_bytes, &bars[t]);
        }
    }
    if (valid && sub == 0)
        for (int i = lane; i < n3; i += 32) own[i] = l_prev[(size_t)n * n3 + i];
    __syncthreads();  // barrier init + own horizons visible
    SCAN_PROF(0);

    double ow[OWNREG ? 3 * KT : 1];
    if (OWNREG) {
#pragma unroll
        for (int i = 0; i < (OWNREG ? 3 * KT : 1); ++i) ow[i] = own[i];
    }
    const double* ownp = OWNREG ? ow : own;

    unsigned* nm = nm_all + (size_t)ag * Npad;
    ScanAcc acc;
    acc.vmask = 0;
    acc.coll0 = 0;
    const bool refill = ntma > stages;
    for (int t0 = 0; t0 < ntma; t0 += S) {
        const int t = t0 + sub;
        if (t < ntma) {
            const int b = t % stages;
#if defined(DMPC_PROF_SCAN)
            const long long w0 = clock64();
#endif
            mbar_wait(&bars[b], (uint32_t)((t / stages) & 1));
#if defined(DMPC_PROF_SCAN)
            scan_wait += clock64() - w0;
#endif
            int tau = t + rot;
            tau -= (tau >= ntma) ? ntma : 0;
            const int base = tau * kTile;
            const int cnt = (N - base < kTile) ? (N - base) : kTile;
            if (valid) scan_tile_hw<KT>(A.P, &A.thr, ownp, n, tiles + (size_t)b * tile_d, base, cnt, nm, acc);
        }
        if (refill) {
            __syncthreads();  // every warp is done with the stages of this round
            if (threadIdx.x == 0) {
                for (int i = 0; i < S; ++i) {
                    const int tt = t0 + i;
                    if (tt < ntma && tt + stages < ntma) {
                        const int b = tt % stages;
                        int tau = tt + stages + rot;
                        tau -= (tau >= ntma) ? ntma : 0;
                        mbar_expect_tx(&bars[b], tile_bytes);
                        tma_bulk_g2s(tiles + (size_t)b * tile_d, l_prev + (size_t)tau * tile_d, tile_bytes,
                                     &bars[b]);
                    }
                }
            }
        }
    }
    SCAN_PROF(1);
#if defined(DMPC_PROF_SCAN)
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&g_prof[4], (unsigned long long)scan_wait);
        atomicAdd(&g_prof[12], (unsigned long long)scan_wait);
        atomicAdd(&g_prof[20], 1ull);
    }
#endif
    const int rem_base = ntma * kTile;
    if (rem_base < N) {
        // caller-owned buffer without tile padding: the ragged last tile is loaded by the threads
        const int cnt = N - rem_base;
        __syncthreads();
        for (int i = threadIdx.x; i < cnt * n3; i += W * S * 32) tiles[i] = l_prev[(size_t)rem_base * n3 + i];
        __syncthreads();
        if (valid && sub == 0) scan_tile_hw<KT>(A.P, &A.thr, ownp, n, tiles, rem_base, cnt, nm, acc);
    }
    // combine the S partial results of an agent (the barrier also publishes the siblings' near masks)
    {
        const unsigned vm = wor(acc.vmask), c0 = wor(acc.coll0);
        if (lane == 0) {
            s_acc[2 * warp] = vm;
            s_acc[2 * warp + 1] = c0;
        }
    }
    __syncthreads();
    SCAN_PROF(2);
    if (!valid || sub != 0) return;
#pragma unroll
    for (int i = 1; i < S; ++i) {
        acc.vmask |= s_acc[2 * (warp + i)];
        acc.coll0 |= s_acc[2 * (warp + i) + 1];
    }
    // neighbour positions from the resident tiles when the whole buffer is in shared memory
    const bool resident = !refill && rem_base >= N;
    const ScanOut so = scan_finish(A.P, own, n, l_prev, nm, acc, A.RMAX, A.grow + (size_t)li * 5 * A.RMAX,
                                   A.gkc + (size_t)li * A.RMAX, A.gidx ? A.gidx + (size_t)li * A.RMAX : nullptr,
                                   list_all + (size_t)ag * A.RMAX, resident ? tiles : nullptr, rot, ntma);
    if (lane == 0) {
        ScanRec r;
        r.kstar = so.kstar;
        r.nv = so.nv;
        r.flag = so.flag;
        r.pad = 0;
        A.scan[li] = r;
        if (A.rq) {
            // route of the throughput layout: heavy = needs the 64-capacity workspace for sure (solveHardDMPC,
            // more than 64 rows) or probably (its active set of the previous MPC step was large)
            const int ng = scen * N + n;
            const bool heavy = !so.flag && (A.P.variant == VAR_HARD || so.nv > kQW ||
                                            (A.diag && A.diag[ng].nact >= kRouteNact));
            if (heavy) A.q_heavy[atomicAdd(&A.rq->n_heavy, 1u)] = li;
            else A.q_light[atomicAdd(&A.rq->n_light, 1u)] = li;
        }
    }
    SCAN_PROF(3);
}

// ---- K1, register-tile layout -----------------------------------------------------------------------------
// The tile loop of scan_kernel is bound by shared-memory wavefronts: every warp (one agent) re-reads the whole
// tile, 2 wavefronts per 64-bit read and lane (profiles/r2f_scan_experiments.txt: the SM runs at ~1.65 warp
// instructions per cycle whatever the occupancy, whatever the number of pair evaluations).  Here the roles are
// swapped: a warp takes a TILE, every lane holds one neighbour's whole horizon in REGISTERS (3 KT doubles) and
// the warp loops over the CTA's W agents, whose horizons are read from shared memory as 128-bit BROADCASTS
// (one wavefront per two doubles): ~2.5x fewer wavefronts per pair.  Tiles are drawn from a shared counter (NW
// warps, any number of ring slots: the warp that has finished a tile re-arms its slot itself with the tile
// `stages` further on -- no CTA-wide barrier in the loop).  Same per-pair arithmetic and decisions as
// scan_pair_hw (high-word compares, exact redo next to a threshold): bit-identical masks.
// Needs a compile-time horizon (KT = 15, 20).  smem: like scan_kernel (+ own horizons padded to 128-bit rows).
template <int KT>
DMPC_D void scan_pair_rt(const DevParams& P, const ScanThr* __restrict__ thr, const double* __restrict__ own_s,
                         const double (&pj)[3 * KT], const double* __restrict__ pj_mem, unsigned& vm_out, unsigned& nm_out,
                         unsigned& c0_out) {
    const double inv_c = thr->inv_c;
    const unsigned hv_m = thr->hv_m;
    unsigned vm = 0, nm = 0, amb = 0, c0 = 0;
    constexpr int n3 = 3 * KT;
    double ow[n3 + 1];
#pragma unroll
    for (int w = 0; w < n3; w += 2) {  // rows of own_all are padded to an even length, 16-byte aligned
        const double2 v = *reinterpret_cast<const double2*>(own_s + w);
        ow[w] = v.x;
        ow[w + 1] = v.y;  // (w + 1 == n3: the padding, never used)
    }
#pragma unroll
    for (int k = 0; k < KT; ++k) {
        const double dx = ow[3 * k] - pj[3 * k];
        const double dy = ow[3 * k + 1] - pj[3 * k + 1];
        const double ez = (ow[3 * k + 2] - pj[3 * k + 2]) * inv_c;
        const unsigned he = hi_word(fma(ez, ez, fma(dy, dy, dx * dx)));
        const unsigned xv = he - hv_m, xn = he - thr->hn_m[k];
        vm |= ((int)xv < 0 ? 1u : 0u) << k;
        nm |= ((int)xn < 0 ? 1u : 0u) << k;
        amb |= (xv <= 2u || xn <= 2u) ? 1u : 0u;
        if (k == 0) {
            const unsigned xc = he - thr->hc_m;
            c0 = ((int)xc < 0) ? 1u : 0u;
            amb |= (xc <= 2u) ? 1u : 0u;
        }
    }
    if (amb) {
        // some estimate sits next to a threshold: redo this pair exactly (operands from shared memory)
        vm = 0;
        nm = 0;
        c0 = 0;
        for (int k = 0; k < KT; ++k) {
            const double s = ell_sq(own_s[3 * k] - pj_mem[3 * k], own_s[3 * k + 1] - pj_mem[3 * k + 1],
                                    own_s[3 * k + 2] - pj_mem[3 * k + 2], P.c);
            if (s < thr->T_viol) vm |= 1u << k;
            if (s < thr->T_near[k]) nm |= 1u << k;
            if (k == 0 && s < thr->T_coll) c0 = 1u;
        }
    }
    vm_out = vm;
    nm_out = nm;
    c0_out = c0;
}

template <int W, int NW, int KT>
__global__ void __launch_bounds__(NW * 32) scan_rt_kernel(const __grid_constant__ StepArgs A, int stages) {
    static_assert(KT > 0, "register-tile scan: compile-time horizon");
    asm volatile("griddepcontrol.launch_dependents;");
    const int nl_s = A.n1 - A.n0;
    const int cps = (nl_s + W - 1) / W;
    const int scen = (A.n_scen > 1) ? (int)blockIdx.x / cps : 0;
    const int blk = (int)blockIdx.x - scen * cps;
    const double* const l_prev = A.l_prev + (size_t)scen * A.lstride;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int K = KT, n3 = 3 * K, n3p = (n3 + 1) / 2 * 2;
    const int N = A.P.N;
    constexpr int tile_d = kTile * n3;
    constexpr uint32_t tile_bytes = (uint32_t)(tile_d * sizeof(double));
    double* tiles = reinterpret_cast<double*>(smem_raw);
    double* own_all = tiles + (size_t)stages * tile_d;
    uint64_t* bars = reinterpret_cast<uint64_t*>(own_all + W * n3p);
    unsigned* s_acc = reinterpret_cast<unsigned*>(bars + kScanMaxStages);  // [2a] violation mask, [2a+1] coll0; [127] tile counter
    unsigned* nm_all = s_acc + 64 * 2;
    const int Npad = round_up(N, kTile);
    int* list_all = reinterpret_cast<int*>(nm_all + (size_t)W * Npad);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int a0 = A.n0 + blk * W;  // first agent of this CTA
    const int na = (A.n1 - a0 < W) ? (A.n1 - a0) : W;
    if (A.ctrl && A.ctrl[scen].done) {
        if (A.rq && A.n_scen > 1 && lane == 0)
            for (int a = warp; a < na; a += NW) A.q_light[atomicAdd(&A.rq->n_light, 1u)] = scen * nl_s + blk * W + a;
        return;
    }
    const int ntma = A.tile_padded ? (N + kTile - 1) / kTile : N / kTile;
    const int rot = ntma ? (int)((unsigned)blk % (unsigned)ntma) : 0;
    if (threadIdx.x == 0) {
        for (int b = 0; b < stages; ++b) mbar_init(&bars[b], 1);
        mbar_fence_init();
        for (int t = 0; t < stages && t < ntma; ++t) {
            int tau = t + rot;
            tau -= (tau >= ntma) ? ntma : 0;
            mbar_expect_tx(&bars[t], tile_bytes);
            tma_bulk_g2s(tiles + (size_t)t * tile_d, l_prev + (size_t)tau * tile_d, tile_bytes, &bars[t]);
        }
        s_acc[127] = 0u;
    }
    for (int i = threadIdx.x; i < 2 * W; i += NW * 32) s_acc[i] = 0u;
    for (int i = threadIdx.x; i < W * n3p; i += NW * 32) {
        const int a = i / n3p, c = i - a * n3p;
        own_all[i] = (a < na && c < n3) ? l_prev[(size_t)(a0 + a) * n3 + c] : 0.0;
    }
    __syncthreads();
    for (;;) {
        int t = 0;
        if (lane == 0) t = (int)atomicAdd(&s_acc[127], 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= ntma) break;
        const int b = t % stages;
        mbar_wait(&bars[b], (uint32_t)((t / stages) & 1));
        int tau = t + rot;
        tau -= (tau >= ntma) ? ntma : 0;
        const int i = tau * kTile + lane;  // this lane's neighbour
        const double* pjm = tiles + (size_t)b * tile_d + (size_t)lane * n3;
        double pj[n3];
#pragma unroll
        for (int w = 0; w < n3; ++w) pj[w] = pjm[w];
        for (int a = 0; a < na; ++a) {
            unsigned vm, nm, c0;
            scan_pair_rt<KT>(A.P, &A.thr, own_all + a * n3p, pj, pjm, vm, nm, c0);
            const bool on = i < N && i != a0 + a;
            nm_all[(size_t)a * Npad + i] = on ? nm : 0u;
            vm = __reduce_or_sync(0xffffffffu, on ? vm : 0u);
            c0 = __reduce_or_sync(0xffffffffu, on ? c0 : 0u);
            if (lane == 0 && (vm | c0)) {
                atomicOr(&s_acc[2 * a], vm);
                atomicOr(&s_acc[2 * a + 1], c0);
            }
        }
        __syncwarp();  // every lane has its copy of the tile: the slot can take the tile `stages` further on
        if (lane == 0 && t + stages < ntma) {
            int tn = t + stages + rot;
            tn -= (tn >= ntma) ? ntma : 0;
            mbar_expect_tx(&bars[b], tile_bytes);
            tma_bulk_g2s(tiles + (size_t)b * tile_d, l_prev + (size_t)tn * tile_d, tile_bytes, &bars[b]);
        }
    }
    const int rem_base = ntma * kTile;
    if (rem_base < N) {
        // caller-owned buffer without tile padding: the ragged last tile goes through the per-agent function
        const int cnt = N - rem_base;
        __syncthreads();
        for (int i = threadIdx.x; i < cnt * n3; i += NW * 32) tiles[i] = l_prev[(size_t)rem_base * n3 + i];
        __syncthreads();
        for (int a = warp; a < na; a += NW) {
            ScanAcc acc;
            acc.vmask = 0;
            acc.coll0 = 0;
            scan_tile_hw<KT>(A.P, &A.thr, own_all + a * n3p, a0 + a, tiles, rem_base, cnt, nm_all + (size_t)a * Npad, acc);
            const unsigned vm = wor(acc.vmask), c0 = wor(acc.coll0);
            if (lane == 0) {
                atomicOr(&s_acc[2 * a], vm);
                atomicOr(&s_acc[2 * a + 1], c0);
            }
        }
    }
    __syncthreads();  // masks and violation bits of all tiles are in place
    for (int a = warp; a < na; a += NW) {
    const int n = a0 + a, li = scen * nl_s + blk * W + a;
    ScanAcc acc;
    acc.vmask = s_acc[2 * a];
    acc.coll0 = s_acc[2 * a + 1];
    const bool resident = stages >= ntma && rem_base >= N;
    const ScanOut so = scan_finish(A.P, own_all + a * n3p, n, l_prev, nm_all + (size_t)a * Npad, acc, A.RMAX,
                                   A.grow + (size_t)li * 5 * A.RMAX, A.gkc + (size_t)li * A.RMAX,
                                   A.gidx ? A.gidx + (size_t)li * A.RMAX : nullptr, list_all + (size_t)a * A.RMAX,
                                   resident ? tiles : nullptr, rot, ntma);
    if (lane == 0) {
        ScanRec r;
        r.kstar = so.kstar;
        r.nv = so.nv;
        r.flag = so.flag;
        r.pad = 0;
        A.scan[li] = r;
        if (A.rq) {
            const int ng = scen * N + n;
            const bool heavy = !so.flag && (A.P.variant == VAR_HARD || so.nv > kQW ||
                                            (A.diag && A.diag[ng].nact >= kRouteNact));
            if (heavy) A.q_heavy[atomicAdd(&A.rq->n_heavy, 1u)] = li;
            else A.q_light[atomicAdd(&A.rq->n_light, 1u)] = li;
        }
    }
    }
}

// ---- K3 (defined first: K2 runs it in its last CTA) --------------------------------------------
// body of K3 for a CTA of NT threads (NT <= 512).  The inputs may have been written by other CTAs of the
// SAME kernel (fused tail): they are read with ld.global.cg (L2), never from a stale L1 line.
template <int NT>
__device__ __forceinline__ void tail_body(const TailArgs& T) {
    __shared__ double s_md[16];
    __shared__ int s_ff[16], s_fi[16];
    const int tid = threadIdx.x;
    const int step = T.ctrl ? T.ctrl->step : 0;
    double md = 0.0;
    int ff = 0x7fffffff, fi = 0x7fffffff;  // lowest failing agent / lowest agent whose QP was infeasible
    // four agents per thread and pass: all loads of a pass are issued before anything is stored, so their
    // L2 round trips overlap (the last CTA is alone on the chip here: this is pure latency)
    constexpr int U = 4;
    const bool rec = T.traj_p && step < T.S;
    for (int base = tid; base < T.N; base += NT * U) {
        double pp[U][3], gg[U][3], r1[U][3], r2[U][3], r3[U][3];
        int stv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int n = base + u * NT;
            stv[u] = ST_SOLVED;
            if (n < T.N) {
#pragma unroll
                for (int x = 0; x < 3; ++x) {
                    if (T.p) {
                        pp[u][x] = __ldcg(T.p + (size_t)T.ld * n + x);
                        gg[u][x] = T.pf[3 * n + x];
                    }
                    if (rec) {
                        r1[u][x] = __ldcg(T.p1 + 3 * n + x);
                        r2[u][x] = __ldcg(T.v1 + 3 * n + x);
                        r3[u][x] = __ldcg(T.a1 + 3 * n + x);
                    }
                }
                if (T.status && n >= T.n0 && n < T.n1) stv[u] = __ldcg(T.status + n);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int n = base + u * NT;
            if (n >= T.N) continue;
            if (T.p) {
                // ReachedGoal.m:4-5
                const double dx = pp[u][0] - gg[u][0], dy = pp[u][1] - gg[u][1], dz = pp[u][2] - gg[u][2];
                md = fmax(md, sqrt(dx * dx + dy * dy + dz * dz));
            }
            if (T.status && n >= T.n0 && n < T.n1) {
                const int st = stv[u];
                if ((!(st & ST_SOLVED) || (st & ST_OUTBOUND)) && n < ff) ff = n;
                if ((st & (ST_INFEASIBLE | ST_QPFAIL)) && n < fi) fi = n;
                if (T.status_hist && step < T.S) T.status_hist[(size_t)step * T.N + n] = st;
            }
            if (rec) {
                const size_t o = 3 * ((size_t)(step + 1) + (size_t)(T.S + 1) * n);
#pragma unroll
                for (int x = 0; x < 3; ++x) {
                    T.traj_p[o + x] = r1[u][x];
                    T.traj_v[o + x] = r2[u][x];
                    T.traj_a[o + x] = r3[u][x];
                }
            }
        }
    }
    for (int o = 16; o; o >>= 1) {
        md = fmax(md, __shfl_xor_sync(0xffffffffu, md, o));
        ff = min(ff, __shfl_xor_sync(0xffffffffu, ff, o));
        fi = min(fi, __shfl_xor_sync(0xffffffffu, fi, o));
    }
    if ((tid & 31) == 0) {
        s_md[tid >> 5] = md;
        s_ff[tid >> 5] = ff;
        s_fi[tid >> 5] = fi;
    }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < NT / 32; ++w) {
            md = fmax(md, s_md[w]);
            ff = min(ff, s_ff[w]);
            fi = min(fi, s_fi[w]);
        }
        const int reached = (T.p && md < T.goal_tol) ? 1 : 0;
        if (T.goal_out) {
            T.goal_out[0] = md;
            T.goal_out[1] = (double)reached;
        }
        if (T.fail_out) *T.fail_out = (ff == 0x7fffffff) ? -1 : ff;
        if (T.rescue_next) *T.rescue_next = 0;
        if (T.ctrl) {
            Ctrl* c = T.ctrl;
            c->goal_dist = md;
            if (ff != 0x7fffffff && c->fail_step < 0) {
                c->fail_step = step;
                c->fail_agent = ff;
            }
            c->step = step + 1;
            if (reached) {
                c->reached = 1;
                c->done = 1;
            }
            // stop_on_fail 1: any failing agent ends the loop; 2: only an infeasible QP does -- what the reference's
            // driver does (test/failure_rate.m:112-124 breaks the trial on ~feasible; `coll` and `outbound` leave
            // feasible = 1 in solveSoftDMPCbound.m and the trial goes on)
            if (c->stop_on_fail == 1 && ff != 0x7fffffff) c->done = 1;
            if (c->stop_on_fail == 2 && fi != 0x7fffffff) {
                c->done = 1;
                c->infeasible = 1;
            }
            if (step + 1 >= c->max_steps) c->done = 1;
        }
    }
    if (T.copy_bytes) {
        __syncthreads();  // first_fail is written
        const int4* src = reinterpret_cast<const int4*>(T.copy_src);
        int4* dst = reinterpret_cast<int4*>(T.copy_dst);
        for (size_t i = tid; i < T.copy_bytes / 16; i += NT) dst[i] = __ldcg(src + i);
    }
}

// ---- K2 -------------------------------------------------------------------------------------
// shared-memory tables of K2: the blob of the register-resident solver (multiple of 32 bytes)
DMPC_HD size_t qp_table_bytes(int K) { return (size_t)tab_fast_size(K) * sizeof(double); }
// per-agent workspace: the register-resident solver's (qp_warp.cuh) or the generic solver's, whichever is larger
DMPC_HD size_t qp_agent_bytes(int K, int QMAX, int RCAP) {
    const size_t a = agent_smem_bytes(K, QMAX, RCAP), b = align_up(qw_smem_bytes(), 16);
    return a > b ? a : b;
}
DMPC_HD size_t qp_smem_bytes(int K, int W, int QMAX, int RCAP) {
    return align_up(qp_table_bytes(K), 16) + (size_t)W * qp_agent_bytes(K, QMAX, RCAP) + 16;
}

// One agent of a QP launch: li = index in [0, n_scen * (n1 - n0)).  MODE 0: the full path -- register-resident
// solver with QC = 64 where its preconditions hold, generic solver (qp_core.cuh) for solveHardDMPC / more than
// 64 rows, global-memory rescue slot for an active set beyond the on-chip capacity.  MODE 1 (light agents of the
// throughput layout): register-resident solver with QC = kLightQ only; returns false when the agent needs the
// full path instead (nothing useful has been written then).
constexpr int kLightQ = 32;
// Inlined into its kernel (a call would cost the hot loop a quarter of its speed): every kernel has exactly ONE
// call site per MODE, so within a kernel an agent's bits do not depend on how it was scheduled.  Two kernels
// (classic / throughput layout) carry separately optimised copies whose FMA contraction may differ: across
// layouts results agree to rounding (~1e-14), not bit for bit.
// HARD: the kernel serves solveHardDMPC (its own instantiation of the register-resident solver, rows on several
// horizon indices); the other kernels keep exactly the code of the on-demand variants.
template <int KT, int MODE, bool HARD = false>
DMPC_D bool qp_agent(const StepArgs& A, int li, const double* tab_s, unsigned char* scratch) {
    const int K = KT ? KT : A.P.K, n3 = 3 * K;
    const int lane = threadIdx.x & 31;
    const int nl_s = A.n1 - A.n0;
    const int scen = (A.n_scen > 1) ? li / nl_s : 0;
    const int n = A.n0 + (li - scen * nl_s);               // agent index within its scenario
    const int ng = scen * A.P.N + n;                        // index into the per-agent arrays
    const double* const l_prev = A.l_prev + (size_t)scen * A.lstride;
    double* const l_new = A.l_new + (size_t)scen * A.lstride;
    if (A.n_scen > 1 && A.ctrl && A.ctrl[scen].done) {
        // this scenario's loop has ended (goal reached / failed): its state is carried over unchanged
        for (int i = lane; i < n3; i += 32) l_new[(size_t)n * n3 + i] = l_prev[(size_t)n * n3 + i];
        if (lane < 3) {
            A.p1[3 * ng + lane] = A.pk[3 * ng + lane];
            A.v1[3 * ng + lane] = A.vk[3 * ng + lane];
            A.a1[3 * ng + lane] = A.ak[3 * ng + lane];
        }
        return true;
    }
#if defined(DMPC_PROF_AGENT)
    const long long agent_t0 = clock64();
#endif
    const ScanRec sr = A.scan[li];
    AgentIO io;
    io.po = A.pk + 3 * ng;
    io.pf = A.pf + 3 * ng;
    io.vo = A.vk + 3 * ng;
    io.ao = A.ak + 3 * ng;
    io.bounds = A.bounds ? A.bounds + 6 * scen : nullptr;
    io.kstar = sr.kstar;
    io.nv = sr.nv;
    io.scanflag = sr.flag;
    io.RMAX = A.RMAX;
    io.grow = A.grow + (size_t)li * 5 * A.RMAX;
    io.gkc = A.gkc + (size_t)li * A.RMAX;
    io.gscr_d = A.gscr_d + (size_t)li * 4 * A.RMAX;
    io.gscr_i = A.gscr_i + (size_t)li * 4 * A.RMAX;
    io.out_p = l_new + (size_t)n * n3;
    io.out_v = A.v_hor ? A.v_hor + (size_t)ng * n3 : nullptr;
    io.out_a = A.a_hor ? A.a_hor + (size_t)ng * n3 : nullptr;
    io.p1 = A.p1 + 3 * ng;
    io.v1 = A.v1 + 3 * ng;
    io.a1 = A.a1 + 3 * ng;
    io.l_prev_n = l_prev + (size_t)n * n3;
    io.dbg_n = n;

    AgentDiag dg;
    int st = 0, it0 = 0;
    // fast path: the register-resident warp solver (qp_warp.cuh) -- every agent of the soft variants whose
    // rows fit (more than 64 rows: a working set of 64 with exact row exchange, qp_warp.cuh; solveHardDMPC
    // with its rows on many horizon steps: the MK instantiation).  The generic solver (qp_core.cuh) takes
    // the rest: horizons beyond 21 steps and -- in a global-memory rescue slot of capacity QBIG -- agents whose active set outgrew the on-chip capacity or
    // whose working set has no free slot left.  One call site, so the generic solver exists once.
    const bool fast_ok = (n3 <= kQW) && (sr.nv <= kRowsFastMax) && !sr.flag && (HARD || A.P.variant != VAR_HARD);
    if (MODE == 1) {
        if (sr.flag) {
            // the scan already decided (predicted collision at the next step / row overflow): the reference
            // returns empty p, v, a -- the agent keeps its horizon and state
            st = sr.flag;
            dg.kstar = sr.kstar; dg.nv = sr.nv; dg.iters = 0; dg.nact = 0;
            for (int i = lane; i < n3; i += 32) io.out_p[i] = io.l_prev_n[i];
            if (lane < 3) {
                io.p1[lane] = io.po[lane];
                io.v1[lane] = io.vo[lane];
                io.a1[lane] = io.ao[lane];
            }
        } else {
            if (!fast_ok) return false;
            if (sr.nv > kQW) return false;  // (row working set: full path only)
            st = agent_solve_fast<KT, kLightQ, false, false>(A.P, tab_s, scratch, A.QMAX, io, &dg);
            if (st & ST_OVERFLOW) return false;  // the active set outgrew the light capacity: full path
        }
    } else {
        bool generic = !fast_ok, rescue = false;
        int cap = A.QMAX;
        if (fast_ok) {
            st = agent_solve_fast<KT, kQW, HARD>(A.P, tab_s, scratch, A.QMAX, io, &dg);
            if ((st & ST_OVERFLOW) && A.rescue) {
                generic = true;
                rescue = true;
                it0 = dg.iters;
                io.start_tries = (st >> 8) & 0xff;  // (the tries already found infeasible are not repeated)
            }
        }
        while (generic) {
            if (rescue) {
                int slot = 0;
                if (lane == 0) slot = atomicAdd(A.rescue_next, 1);
                slot = __shfl_sync(0xffffffffu, slot, 0);
                if (slot >= A.n_rescue) break;
                if (lane == 0 && A.ctrl) atomicAdd(&A.ctrl[scen].rescue_used, 1);
                scratch = A.rescue + (size_t)slot * A.rescue_bytes;
                cap = A.QBIG;
            }
            st = agent_solve<0>(A.P, A.tab, scratch, cap, A.RCAP, io, &dg);  // tables from global memory (L1/L2)
            dg.iters += it0;
            if (rescue || !(st & ST_OVERFLOW) || sr.flag || !A.rescue) break;
            rescue = true;
            it0 = dg.iters;
        }
    }
#if defined(DMPC_PROF_AGENT)
    dg.nact = (int)((clock64() - agent_t0) >> 4);  // profiling build: cycles / 16 of this agent's solve
#endif
    if (lane == 0) {
        A.status[ng] = st;
        if (A.diag) A.diag[ng] = dg;
    }
    return true;
}

template <int W, int KT, bool HARD = false>
__global__ void __launch_bounds__(W * 32, 1) qp_kernel(const __grid_constant__ StepArgs A) {
    if (A.ctrl && A.n_scen == 1 && A.ctrl->done) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int K = KT ? KT : A.P.K;
    const size_t tab_bytes = align_up(qp_table_bytes(K), 16);
    const size_t per_warp = qp_agent_bytes(K, A.QMAX, A.RCAP);
    double* tab_s = reinterpret_cast<double*>(smem_raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + tab_bytes + (size_t)W * per_warp);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        mbar_expect_tx(bar, (uint32_t)qp_table_bytes(K));
        tma_bulk_g2s(tab_s, A.tab + tab_fast_offset(K), (uint32_t)qp_table_bytes(K), bar);
    }
    // everything above touches constants only; the scan kernel's rows are read below
    asm volatile("griddepcontrol.wait;" ::: "memory");
    __syncthreads();
    // one agent per warp while the grid covers the swarm; beyond that the CTAs are persistent (one per SM:
    // the per-agent workspace fills shared memory) and every warp takes its next agent from a device
    // counter as soon as it is done -- the solve times differ by 50x, a static assignment would leave
    // three warps of a CTA idle behind its slowest agent
    const int nl = (A.n1 - A.n0) * A.n_scen;  // agents of this launch
    const bool queued = (int)gridDim.x * W < nl;
    int li = blockIdx.x * W + warp;
    unsigned char* const scratch = smem_raw + tab_bytes + (size_t)warp * per_warp;
    while (li < nl) {
        mbar_wait(bar, 0);  // tables have landed
        qp_agent<KT, 0, HARD>(A, li, tab_s, scratch);
        if (!queued) break;
        // (requesting the next index BEFORE the solve, to hide the atomic's round trip, was measured: a warp stuck on a
        // heavy agent then holds a claimed agent hostage -- N=2000 QP 304 -> 356 us, C5 773 -> 798 us)
        if (lane == 0) li = (int)gridDim.x * W + (int)atomicAdd(A.work_cnt, 1u);
        li = __shfl_sync(0xffffffffu, li, 0);
        __syncwarp();
    }  // agents of this warp
    {
        // the last CTA to arrive has every agent's result behind it: it resets the counters and, in the
        // resident loop and the host step, runs the tail of the step
        __shared__ int s_last;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_last = (atomicAdd(A.done_cnt, 1u) == gridDim.x - 1) ? 1 : 0;
        __syncthreads();
        if (s_last) {
            __threadfence();
            if (A.fuse_tail) tail_body<W * 32>(A.T);
            if (threadIdx.x == 0) {
                *A.done_cnt = 0;
                *A.work_cnt = 0;
            }
        }
    }
}

// ---- K2, throughput layout (more agents than one wave of the layout above) -----------------------------
// The solve times of a step differ by 50x and most agents end with a small active set.  Light agents run
// with an active-set capacity of kLightQ = 32: M shrinks from 34 KB to 8.7 KB, the per-agent workspace to
// 18 KB.  Heavy agents need the 64-capacity workspace (46 KB).  A CTA (one per SM, persistent) has
// kHeavyW = 2 warps that own a heavy workspace and kLightW - kHeavyW = 6 warps that own a light one: eight
// agents are resident per SM (two warps per sub-partition: the solver is a chain of dependent fixed-latency
// instructions, a second warp fills issue slots the first one leaves empty; shared-memory bandwidth -- ~60 KB
// of traffic per active-set iteration -- is what stops more).  Who is heavy is PREDICTED by the scan kernel from
// the size of the agent's active set in the previous MPC step (route queues) and corrected on the fly: a
// light agent that outgrows its capacity is pushed to the heavy queue.  There are no phases: a heavy-capable
// warp serves the heavy queue whenever it has an entry and light agents otherwise; when the light queue is
// empty it drains the heavy queue and leaves once its queue index stays empty after EVERY warp of the grid
// has left the light queue (nothing can be pushed any more).  The result of an agent does not depend on the
// route it took (same arithmetic in the same order; only the capacity differs).
DMPC_HD size_t qp2_light_bytes() { return align_up(qw_smem_bytes(kLightQ), 16); }
DMPC_HD size_t qp2_smem_bytes(int K, int QMAX, int RCAP) {
    const size_t a = (size_t)kHeavyW * qp_agent_bytes(K, QMAX, RCAP) + (size_t)(kLightW - kHeavyW) * qp2_light_bytes();
    const size_t b = (size_t)kHeavyMax * qp_agent_bytes(K, QMAX, RCAP);
    return align_up(qp_table_bytes(K), 16) + (a > b ? a : b) + 16;
}
DMPC_D unsigned ld_volatile_u32(const unsigned* p) { return *reinterpret_cast<const volatile unsigned*>(p); }
DMPC_D int ld_volatile_i32(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

template <int KT>
__global__ void __launch_bounds__(kLightW * 32, 1) qp2_kernel(const __grid_constant__ StepArgs A) {
    if (A.ctrl && A.n_scen == 1 && A.ctrl->done) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int K = KT ? KT : A.P.K;
    const size_t tab_bytes = align_up(qp_table_bytes(K), 16);
    const size_t per_heavy = qp_agent_bytes(K, A.QMAX, A.RCAP), per_light = qp2_light_bytes();
    double* tab_s = reinterpret_cast<double*>(smem_raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + qp2_smem_bytes(K, A.QMAX, A.RCAP) - 16);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    RouteQ* const rq = A.rq;

    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        mbar_expect_tx(bar, (uint32_t)qp_table_bytes(K));
        tma_bulk_g2s(tab_s, A.tab + tab_fast_offset(K), (uint32_t)qp_table_bytes(K), bar);
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");  // the scan kernel's rows and route queues are complete
    __syncthreads();
    mbar_wait(bar, 0);  // tables have landed
    const unsigned nlight = ld_volatile_u32(&rq->n_light), nheavy0 = ld_volatile_u32(&rq->n_heavy);
    // the mix of this step decides how the SM's shared memory is split (uniform over the grid: every CTA reads
    // the same counters): many heavy agents -> 4 heavy workspaces and no light-only warps (the classic split;
    // the heavy agents are the critical path then), else 2 heavy + 6 light workspaces
    const int n_heavy_ws = (10u * nheavy0 > nlight + nheavy0) ? kHeavyMax : kHeavyW;
    const int n_ws = (n_heavy_ws == kHeavyMax) ? kHeavyMax : kLightW;  // warps that take agents at all
    const bool heavy_capable = warp < n_heavy_ws;
    unsigned char* const scratch =
        heavy_capable ? smem_raw + tab_bytes + (size_t)warp * per_heavy
                      : smem_raw + tab_bytes + (size_t)n_heavy_ws * per_heavy + (size_t)(warp - n_heavy_ws) * per_light;
    const unsigned all_warps = gridDim.x * (unsigned)kLightW;
    int pending = -1;  // heavy-queue index this warp has taken but not yet served
    bool light_left = warp < n_ws;
    if (!light_left && lane == 0) atomicAdd(&rq->light_done, 1u);  // a warp without a workspace pushes nothing
    bool drain = false;
    while (warp < n_ws) {
        // ---- which agent next: a heavy one if one is waiting (heavy-capable warps), else a light one; when the
        //      light queue is empty, wait for heavy entries until none can come any more -----------------------
        int li = -1, mode = -1;  // mode 0: full path, 1: light path
        if (heavy_capable) {
            int fin = 0;
            if (lane == 0) {
                if (pending < 0 && (drain || ld_volatile_u32(&rq->n_heavy) > ld_volatile_u32(&rq->head_heavy)))
                    pending = (int)atomicAdd(&rq->head_heavy, 1u);
                while (pending >= 0) {
                    li = ld_volatile_i32(&A.q_heavy[pending]);
                    if (li >= 0 || !drain) break;
                    if (ld_volatile_u32(&rq->light_done) == all_warps) {
                        __threadfence();
                        li = ld_volatile_i32(&A.q_heavy[pending]);  // nothing can be pushed any more
                        fin = li < 0;
                        break;
                    }
                    __nanosleep(256);
                }
                if (li >= 0) {
                    A.q_heavy[pending] = -1;  // the queue cleans itself
                    pending = -1;
                }
            }
            li = __shfl_sync(0xffffffffu, li, 0);
            if (__shfl_sync(0xffffffffu, fin, 0)) break;
            if (li >= 0) mode = 0;
        }
        if (mode < 0 && light_left) {
            unsigned idx = 0;
            if (lane == 0) idx = atomicAdd(&rq->head_light, 1u);
            idx = __shfl_sync(0xffffffffu, idx, 0);
            if (idx < nlight) {
                li = A.q_light[idx];
                // (classic split: every agent takes the full path at once -- active sets grow from step to step
                // in the dense phase, a 32-capacity attempt would often be wasted)
                mode = (n_heavy_ws == kHeavyMax) ? 0 : 1;
            } else {
                light_left = false;  // this warp pushes nothing any more
                if (lane == 0) {
                    __threadfence();
                    atomicAdd(&rq->light_done, 1u);
                }
            }
        }
        if (mode < 0) {
            if (light_left) continue;      // (a heavy-capable warp whose pending entry is not there yet)
            if (!heavy_capable) break;
            drain = true;
            continue;
        }
        if (mode == 0) {
            qp_agent<KT, 0>(A, li, tab_s, scratch);
        } else if (!qp_agent<KT, 1>(A, li, tab_s, scratch)) {
            if (lane == 0) {
                const unsigned pos = atomicAdd(&rq->n_heavy, 1u);
                *reinterpret_cast<volatile int*>(&A.q_heavy[pos]) = li;
                __threadfence();
            }
        }
        __syncwarp();
    }
    {
        __shared__ int s_last;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_last = (atomicAdd(&rq->done_ctas, 1u) == gridDim.x - 1) ? 1 : 0;
        __syncthreads();
        if (s_last) {
            __threadfence();
            if (A.fuse_tail) tail_body<kLightW * 32>(A.T);
            if (threadIdx.x == 0) {
                rq->n_light = 0; rq->n_heavy = 0; rq->head_light = 0; rq->head_heavy = 0;
                rq->light_done = 0; rq->done_ctas = 0;
            }
        }
    }
}

}  // namespace dmpc

// small_kernels.cuh -- the non-template kernels of the library (one translation unit: dmpc_b200.cu):
// stand-alone tail (ReachedGoal.m), initDMPC.m for all agents, the per-agent helper drop-ins.
#pragma once
#include "dmpc_kernels.cuh"

namespace dmpc {

__global__ void __launch_bounds__(256) tail_kernel(const __grid_constant__ TailArgs T) {
    if (T.ctrl && T.ctrl->done) return;
    tail_body<256>(T);
}

// per-scenario tail of a batched step (n_scen > 1): one CTA per scenario runs ReachedGoal.m, the lowest
// failing agent, the trajectory record and the scenario's own loop control word
__global__ void __launch_bounds__(128) tail_batch_kernel(const __grid_constant__ TailArgs T0) {
    const int s = blockIdx.x;
    if (s == 0 && threadIdx.x == 0 && T0.rescue_next) *T0.rescue_next = 0;
    TailArgs T = T0;
    const size_t o3 = 3 * (size_t)s * T.N;
    T.ctrl = T0.ctrl + s;
    if (T.ctrl->done) return;  // this scenario's loop has ended
    T.p = T0.p + o3;
    T.pf = T0.pf + o3;
    T.status = T0.status + (size_t)s * T.N;
    T.p1 = T0.p1 + o3;
    T.v1 = T0.v1 + o3;
    T.a1 = T0.a1 + o3;
    if (T0.traj_p) {
        const size_t ot = 3 * (size_t)(T.S + 1) * T.N * s;
        T.traj_p = T0.traj_p + ot;
        T.traj_v = T0.traj_v + ot;
        T.traj_a = T0.traj_a + ot;
    }
    if (T0.status_hist) T.status_hist = T0.status_hist + (size_t)s * T.S * T.N;
    T.goal_out = T0.goal_out + 2 * s;
    T.fail_out = T0.fail_out + s;
    T.rescue_next = nullptr;
    T.copy_bytes = 0;
    tail_body<128>(T);
}

// ---- initDMPC.m:1-13 for all agents -------------------------------------------------------------
__global__ void init_kernel(int N, int K, double h, double init_div, const double* __restrict__ po,
                            const double* __restrict__ pf, double* l, double* pk, double* vk, double* ak) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    for (int x = 0; x < 3; ++x) {
        const double o = po[3 * n + x];
        const double d = pf[3 * n + x] - o;
        for (int k = 0; k < K; ++k) {
            // t = 0:h:(K-1)*h ; p(:,i) = po + 1*t(i)*diff/10
            const double t = (double)k * h;
            l[(size_t)n * 3 * K + 3 * k + x] = o + __ddiv_rn(__dmul_rn(t, d), init_div);
        }
        pk[3 * n + x] = o;
        vk[3 * n + x] = 0.0;
        ak[3 * n + x] = 0.0;
    }
}

// ---- randomTest.m:1-57 / randomExchange.m:1-56 on the device: one CTA per scenario ---------------------------
// Rejection sampling of N start points (and N goals) with pairwise distance > rmin inside the arena; a point
// that cannot be placed in max_iter tries restarts its whole set like the reference does (:9-27).  The random
// stream is counter based (splitmix64 of (seed, scenario, set, draw index)), so the CPU oracle generates the
// same scenarios bit for bit; MATLAB's own stream cannot be reproduced (the scripts never seed it).
//   mode 0 (randomTest): start and goal sets independent, ellipsoidal distance ||E1 (p - q)||, E1 = diag(1,1,1/c)
//   mode 1 (randomExchange): Euclidean distance; goals = a random permutation of the starts in which every
//                            agent moves (:32-49)
DMPC_HD unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
DMPC_HD unsigned long long gen_key(unsigned long long seed, int scen, int set) {
    return splitmix64(seed + 0x632BE59BD9B4E019ull * (unsigned long long)(2 * scen + set + 1));
}
DMPC_HD double gen_u01(unsigned long long key, unsigned long long idx) {
    return (double)(splitmix64(key + idx) >> 11) * (1.0 / 9007199254740992.0);  // [0, 1), 53 bits
}

constexpr int kGenMaxN = 4096;  // points of a scenario held in shared memory (3 x 8 B each)
__global__ void __launch_bounds__(256) gen_scenarios_kernel(int N, int mode, unsigned long long seed, double rmin,
                                                            double inv_c, double pmin0, double pmin1, double pmin2,
                                                            double pmax0, double pmax1, double pmax2, int max_iter,
                                                            double* po, double* pf) {
    extern __shared__ double g_pts[];  // x[N], y[N], z[N]
    const int scen = blockIdx.x, tid = threadIdx.x;
    double* px = g_pts;
    double* py = g_pts + N;
    double* pz = g_pts + 2 * N;
    const double lo[3] = {pmin0, pmin1, pmin2};
    const double rg[3] = {__dsub_rn(pmax0, pmin0), __dsub_rn(pmax1, pmin1), __dsub_rn(pmax2, pmin2)};
    const double ic = (mode == 0) ? inv_c : 1.0;
    const int nsets = (mode == 0) ? 2 : 1;
    for (int set = 0; set < nsets; ++set) {
        const unsigned long long key = gen_key(seed, scen, set);
        unsigned long long t = 0;  // candidates drawn so far (every thread keeps the same count)
        int n = 0, tries = 0;
        while (n < N) {
            const double cx = __dadd_rn(lo[0], __dmul_rn(rg[0], gen_u01(key, 3 * t)));
            const double cy = __dadd_rn(lo[1], __dmul_rn(rg[1], gen_u01(key, 3 * t + 1)));
            const double cz = __dadd_rn(lo[2], __dmul_rn(rg[2], gen_u01(key, 3 * t + 2)));
            ++t;
            int ok = 1;
            for (int j = tid; j < n; j += 256) {
                const double dx = __dsub_rn(px[j], cx), dy = __dsub_rn(py[j], cy), ez = __dmul_rn(ic, __dsub_rn(pz[j], cz));
                const double dist = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(ez, ez)));
                ok &= dist > rmin;
            }
            ok = __syncthreads_and(ok);  // (also orders the reads above before the store below)
            if (ok) {
                if (tid == 0) { px[n] = cx; py[n] = cy; pz[n] = cz; }
                ++n;
                tries = 0;
            } else if (++tries > max_iter) {  // randomTest.m:23-25: give up on this set and start it again
                n = 0;
                tries = 0;
            }
            __syncthreads();
        }
        double* out = (set == 0 ? po : pf) + 3 * (size_t)N * scen;
        for (int j = tid; j < N; j += 256) {
            out[3 * j] = px[j];
            out[3 * j + 1] = py[j];
            out[3 * j + 2] = pz[j];
        }
        __syncthreads();
    }
    if (mode == 1) {
        // randomExchange.m:32-49: perm(i) drawn from what is left without i; the last two steps make sure that the
        // last agent does not keep its own place.  Sequential (N steps of O(N)), one thread; `array` lives in the
        // x coordinates' shared memory after they have been written out.
        if (tid == 0) {
            int* array = reinterpret_cast<int*>(g_pts);          // remaining indices, ascending
            int* perm = array + N;
            const unsigned long long key = gen_key(seed, scen, 1);
            int left = N;
            for (int i = 0; i < N; ++i) array[i] = i;
            for (int i = 0; i < N; ++i) {
                // array_aux = array without i
                int pick;
                if (i == N - 1) {
                    pick = array[0];
                } else {
                    int last_aux = array[left - 1];
                    if (last_aux == i) last_aux = array[left - 2];
                    if (i == N - 2 && last_aux == N - 1) {
                        pick = N - 1;
                    } else {
                        int j = (int)(gen_u01(key, (unsigned long long)i) * (double)(N - i - 1));  // randi([1 N-i]) - 1
                        // j-th element of array_aux
                        int cnt = -1, e = 0;
                        for (; e < left; ++e) {
                            if (array[e] == i) continue;
                            if (++cnt == j) break;
                        }
                        pick = array[e];
                    }
                }
                perm[i] = pick;
                int w = 0;
                for (int e = 0; e < left; ++e)
                    if (array[e] != pick) array[w++] = array[e];
                left = w;
            }
            const double* src = po + 3 * (size_t)N * scen;
            double* dst = pf + 3 * (size_t)N * scen;
            for (int i = 0; i < N; ++i)
                for (int x = 0; x < 3; ++x) dst[3 * i + x] = src[3 * perm[i] + x];
        }
        __syncthreads();
    }
}

// ---- CheckCollSoftDMPC.m:1-17 for one agent and one horizon step (helper drop-in) ---------------
// out_d[0] = min distance, out_i[0] = any(violation)
__global__ void __launch_bounds__(256) check_coll_kernel(DevParams P, double px, double py, double pz, const double* l,
                                                         int n, int k1, unsigned char* violation,
                                                         unsigned char* viol_constr, double* out_d, int* out_i) {
    __shared__ double s_md[8];
    __shared__ int s_any[8];
    const int tid = threadIdx.x;
    const double thr = neigh_thr(P, k1);
    double md = INFINITY;
    int any = 0;
    for (int i = tid; i < P.N; i += 256) {
        unsigned char v = 0, vc = 0;
        if (i != n) {
            const double* pj = l + 3 * ((size_t)(k1 - 1) + (size_t)P.K * i);
            const double dist = ell_dist(px - pj[0], py - pj[1], pz - pj[2], P.c);
            v = dist < P.rmin;
            vc = dist < thr;
            md = fmin(md, dist);
            any |= v;
        }
        violation[i] = v;
        viol_constr[i] = vc;
    }
    for (int o = 16; o; o >>= 1) {
        md = fmin(md, __shfl_xor_sync(0xffffffffu, md, o));
        any |= __shfl_xor_sync(0xffffffffu, any, o);
    }
    if ((tid & 31) == 0) {
        s_md[tid >> 5] = md;
        s_any[tid >> 5] = any;
    }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 8; ++w) {
            md = fmin(md, s_md[w]);
            any |= s_any[w];
        }
        out_d[0] = md;
        out_i[0] = any;
    }
}

// ---- CollConstr*DMPC.m for one agent and one horizon step: dense rows like the reference -------
// one warp; rows in ascending neighbour order.  Ain cap x 3K column-major.
__global__ void __launch_bounds__(32) coll_constr_kernel(DevParams P, const double* __restrict__ tab, double px,
                                                         double py, double pz, double po0, double po1, double po2,
                                                         double vo0, double vo1, double vo2, int n, int k1,
                                                         const double* l, const unsigned char* mask, int cap,
                                                         double* Ain, double* bin, double* prev_dist, int* nrows) {
    const int K = P.K, N = P.N, lane = threadIdx.x;
    const bool hard = (P.variant == VAR_HARD);
    const int kc1 = (P.variant == VAR_SOFT_BOUND2) ? k1 - 1 : k1;  // CollConstrSoftDMPC2.m:8
    const double* lam = tab;
    const double* tt = tab + K * K;
    const double c2 = P.c * P.c;
    int nv = 0;
    for (int base = 0; base < N; base += 32) {
        const int i = base + lane;
        bool hit = false;
        double dx = 0, dy = 0, dz = 0, dist = 0;
        if (i < N && i != n && (hard || mask[i])) {
            const double* pj = l + 3 * ((size_t)(k1 - 1) + (size_t)K * i);
            dx = px - pj[0];
            dy = py - pj[1];
            dz = pz - pj[2];
            dist = ell_dist(dx, dy, dz, P.c);
            hit = hard ? (dist < P.hard_radius) : true;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (hit) {
            const int slot = nv + __popc(bal & ((1u << lane) - 1u));
            if (slot < cap) {
                const double d0 = dx, d1 = dy, d2 = dz / c2;
                const double dp = d0 * px + d1 * py + d2 * pz;
                const double tk = (kc1 >= 1) ? tt[kc1 - 1] : 0.0;
                const double dq = d0 * (po0 + tk * vo0) + d1 * (po1 + tk * vo1) + d2 * (po2 + tk * vo2);
                const double r = dist * ((P.rmin - dist) + dp / dist) - dq;
                for (int j = 0; j < K; ++j) {
                    const double lj = (kc1 >= 1) ? lam[(kc1 - 1) * K + j] : 0.0;
                    Ain[slot + (size_t)cap * (3 * j + 0)] = -d0 * lj;
                    Ain[slot + (size_t)cap * (3 * j + 1)] = -d1 * lj;
                    Ain[slot + (size_t)cap * (3 * j + 2)] = -d2 * lj;
                }
                bin[slot] = -r;
                prev_dist[slot] = dist;
            }
        }
        nv += __popc(bal);
    }
    if (lane == 0) *nrows = nv;
}

// ---- propStatedmpc.m:1-8 for a batch: a 3K x B -> p, v 3K x B -----------------------------------
__global__ void prop_state_kernel(int B, int K, const double* __restrict__ tab, double h, const double* po,
                                  const double* vo, const double* a, double* p, double* v) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int n3 = 3 * K;
    if (idx >= B * n3) return;
    const int b = idx / n3, i = idx - b * n3, k = i / 3, x = i - 3 * k;
    const double* lam = tab;
    const double* tt = tab + K * K;
    double sp = 0.0, sv = 0.0;
    for (int j = 0; j <= k; ++j) {
        const double aj = a[(size_t)b * n3 + 3 * j + x];
        sp = fma(lam[k * K + j], aj, sp);
        sv += h * aj;
    }
    p[idx] = sp + (po[3 * b + x] + tt[k] * vo[3 * b + x]);
    v[idx] = sv + vo[3 * b + x];
}

}  // namespace dmpc

// postprocess.cuh -- post-processing of a finished transition on the device.
//
// Reference: test/failure_rate.m:134-195 (the part of t_dmpc that follows the MPC loop): time scaling to
// the velocity / acceleration limits (:136-162), 100 Hz cubic-spline interpolation (`spline`, not-a-knot,
// :164-168), the O(N^2 T) pairwise collision check on the interpolated positions (:170-181), total
// distance (:183) and trajectory time (:185-194).  C++: dmpc.cpp:1912-2086.
// Layout: trajectories are MATLAB's 3 x S x N (agent-major, [N][S][3]); interpolated 3 x nt x N.
// Arithmetic that feeds bit-compared results (r_factor, the re-integration) uses separately rounded
// multiply / add / divide in MATLAB's order.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace dmpc {

// ---- r_factor = min over agents and steps of amax/|a|, vmax/|v|  (failure_rate.m:141-144) -----------------
// positive doubles (and +inf) order like their bit patterns: atomicMin on the bits
__global__ void __launch_bounds__(256) pp_rfactor_kernel(int n, const double* __restrict__ vk, const double* __restrict__ ak,
                                                         double vmax, double amax, unsigned long long* out_bits) {
    double m = INFINITY;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const double* a = ak + 3 * (size_t)e;
        const double* v = vk + 3 * (size_t)e;
        const double na = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(a[0], a[0]), __dmul_rn(a[1], a[1])), __dmul_rn(a[2], a[2])));
        const double nv = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(v[0], v[0]), __dmul_rn(v[1], v[1])), __dmul_rn(v[2], v[2])));
        m = fmin(m, fmin(__ddiv_rn(amax, na), __ddiv_rn(vmax, nv)));
    }
    for (int o = 16; o; o >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMin(out_bits, (unsigned long long)__double_as_longlong(m));
}

// ---- ak *= r; vk, pk re-integrated with h_scaled  (failure_rate.m:156-162), one thread per (agent, axis) ---
__global__ void pp_rescale_kernel(int N, int S, double r, double hs, double* pk, double* vk, double* ak) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 3 * N) return;
    const int n = e / 3, x = e - 3 * n;
    double* p = pk + (size_t)n * S * 3 + x;
    double* v = vk + (size_t)n * S * 3 + x;
    double* a = ak + (size_t)n * S * 3 + x;
    const double hh = __ddiv_rn(__dmul_rn(hs, hs), 2.0);
    double pc = p[0], vc = v[0];
    for (int k = 0; k < S - 1; ++k) {
        const double as = __dmul_rn(a[3 * k], r);
        a[3 * k] = as;
        const double vn = __dadd_rn(vc, __dmul_rn(hs, as));
        const double pn = __dadd_rn(__dadd_rn(pc, __dmul_rn(hs, vc)), __dmul_rn(hh, as));
        v[3 * (k + 1)] = vn;
        p[3 * (k + 1)] = pn;
        vc = vn;
        pc = pn;
    }
}

// ---- slopes of MATLAB's `spline` (not-a-knot) on the uniform knots tk = i*hs: one thread per series ------
// series e = (q, n, x): q = 0 pk, 1 vk, 2 ak.  The system is tridiagonal (spline.m); Thomas algorithm with
// the modified diagonal kept in `work` (S doubles per series).
__global__ void pp_slopes_kernel(int N, int S, int nq, double hs, const double* pk, const double* vk, const double* ak,
                                 double* slopes, double* work) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nq * 3 * N) return;
    const int q = e / (3 * N), r = e - q * 3 * N, n = r / 3, x = r - 3 * n;
    const double* y = (q == 0 ? pk : (q == 1 ? vk : ak)) + (size_t)n * S * 3 + x;
    double* s = slopes + (size_t)e * S;
    double* w = work + (size_t)e * S;
    // knots tk[i] = i * hs as the reference builds them (0:h_scaled:T)
    auto tk = [hs](int i) { return (double)i * hs; };
    auto dx = [&](int i) { return tk(i + 1) - tk(i); };
    auto dd = [&](int i) { return (y[3 * (i + 1)] - y[3 * i]) / dx(i); };
    const int nn = S;
    // row 0: dx1 s0 + x31 s1 = rhs0
    const double x31 = tk(2) - tk(0), xn = tk(nn - 1) - tk(nn - 3);
    double diag = dx(1), up = x31;
    double rhs = ((dx(0) + 2.0 * x31) * dx(1) * dd(0) + dx(0) * dx(0) * dd(1)) / x31;
    w[0] = up / diag;
    s[0] = rhs / diag;
    for (int i = 1; i < nn - 1; ++i) {
        const double lo = dx(i), di = 2.0 * (dx(i - 1) + dx(i)), u = dx(i - 1);
        const double bi = 3.0 * (dx(i) * dd(i - 1) + dx(i - 1) * dd(i));
        const double den = di - lo * w[i - 1];
        w[i] = u / den;
        s[i] = (bi - lo * s[i - 1]) / den;
    }
    {
        const double lo = xn, di = dx(nn - 2);
        const double bi = (dx(nn - 2) * dx(nn - 2) * dd(nn - 3) + (2.0 * xn + dx(nn - 2)) * dx(nn - 3) * dd(nn - 2)) / xn;
        const double den = di - lo * w[nn - 2];
        s[nn - 1] = (bi - lo * s[nn - 2]) / den;
    }
    for (int i = nn - 2; i >= 0; --i) s[i] -= w[i] * s[i + 1];
}

// ---- evaluation at t_m = m*Ts (ppval: right-continuous pieces, last point in the last piece) --------------
// one thread per (q, agent, sample); out[q]: [N][nt][3]
__global__ void pp_eval_kernel(int N, int S, int nt, int nq, double hs, double Ts, const double* pk, const double* vk,
                               const double* ak, const double* slopes, double* op, double* ov, double* oa) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)nq * N * nt) return;
    const int q = (int)(e / ((long long)N * nt));
    const long long r = e - (long long)q * N * nt;
    const int n = (int)(r / nt), m = (int)(r - (long long)n * nt);
    const double t = (double)m * Ts;
    int i = (int)(t / hs);
    if (i > S - 2) i = S - 2;
    while (i > 0 && (double)i * hs > t) --i;
    while (i < S - 2 && (double)(i + 1) * hs <= t) ++i;
    const double x0 = (double)i * hs, hx = (double)(i + 1) * hs - x0, tt = t - x0;
    const double* ysrc = (q == 0 ? pk : (q == 1 ? vk : ak)) + (size_t)n * S * 3;
    double* out = (q == 0 ? op : (q == 1 ? ov : oa)) + ((size_t)n * nt + m) * 3;
#pragma unroll
    for (int x = 0; x < 3; ++x) {
        const double* s = slopes + ((size_t)(q * 3 * N + 3 * n + x)) * S;
        const double y0 = ysrc[3 * i + x], y1 = ysrc[3 * (i + 1) + x], s0 = s[i], s1 = s[i + 1];
        const double d = (y1 - y0) / hx;
        const double c2 = (3.0 * d - 2.0 * s0 - s1) / hx;
        const double c3 = (s0 - 2.0 * d + s1) / (hx * hx);
        out[x] = y0 + tt * (s0 + tt * (c2 + tt * c3));
    }
}

// ---- pairwise post-interpolation check (failure_rate.m:170-181): min over pairs and samples of
//      ||E1 (p_i - p_j)||, E1 = diag(1,1,1/c).  16 x 16 agent tiles, time in chunks staged in shared memory;
//      the minimum of the SQUARED metric is reduced (sqrt is monotone and correctly rounded) ----------------
constexpr int kPairTile = 16, kPairChunk = 48;
// C_POW2: c is a power of two, so dz / c == dz * (1/c) exactly and the division can be a multiplication
template <bool C_POW2>
__global__ void __launch_bounds__(kPairTile* kPairTile) pp_pairs_kernel(int N, int nt, double c, double inv_c,
                                                                        const double* __restrict__ p,
                                                                        unsigned long long* out_bits) {
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj < bi) return;  // unordered pairs: upper triangle of tiles
    // [agent][sample][xyz], one sample of padding per agent: the 16 agents of a tile then sit in different
    // banks (without it every lane of a half-warp hits the same bank: measured 4.49 ms -> 1.28 ms at N=500)
    __shared__ double si[kPairTile][kPairChunk + 1][3];
    __shared__ double sj[kPairTile][kPairChunk + 1][3];
    const int ti = threadIdx.y, tj = threadIdx.x, tid = ti * kPairTile + tj;
    const int i = bi * kPairTile + ti, j = bj * kPairTile + tj;
    const bool active = i < N && j < N && i < j;
    double best = INFINITY, best2 = INFINITY;
    for (int m0 = 0; m0 < nt; m0 += kPairChunk) {
        const int len = (nt - m0 < kPairChunk) ? nt - m0 : kPairChunk;
        for (int e = tid; e < kPairTile * len * 3; e += kPairTile * kPairTile) {
            const int a = e / (len * 3), rr = e - a * len * 3;
            const int gi = bi * kPairTile + a, gj = bj * kPairTile + a;
            (&si[a][0][0])[rr] = gi < N ? p[((size_t)gi * nt + m0) * 3 + rr] : 0.0;
            (&sj[a][0][0])[rr] = gj < N ? p[((size_t)gj * nt + m0) * 3 + rr] : 0.0;
        }
        __syncthreads();
        if (active) {
#pragma unroll 4
            for (int m = 0; m < len; ++m) {
                const double dx = sj[tj][m][0] - si[ti][m][0], dy = sj[tj][m][1] - si[ti][m][1];
                const double d2 = sj[tj][m][2] - si[ti][m][2];
                const double dz = C_POW2 ? __dmul_rn(d2, inv_c) : __ddiv_rn(d2, c);
                const double s = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                if (m & 1) best2 = fmin(best2, s);
                else best = fmin(best, s);
            }
        }
        __syncthreads();
    }
    best = fmin(best, best2);
    for (int o = 16; o; o >>= 1) best = fmin(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((tid & 31) == 0) atomicMin(out_bits, (unsigned long long)__double_as_longlong(best));
}

// ---- per agent: travelled distance (failure_rate.m:183) and the last sample farther than goal_radius from
//      the goal (:185-193); one warp per agent, fixed summation order ------------------------------------------
__global__ void __launch_bounds__(128) pp_stats_kernel(int N, int nt, double goal_radius, const double* __restrict__ p,
                                                       const double* __restrict__ pf, double* dist_out, int* tidx_out) {
    const int n = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= N) return;
    const double* pp = p + (size_t)n * nt * 3;
    const double g0 = pf[3 * n], g1 = pf[3 * n + 1], g2 = pf[3 * n + 2];
    double sum = 0.0;
    int last = -1;
    for (int m = lane; m < nt; m += 32) {
        if (m + 1 < nt) {
            const double dx = pp[3 * (m + 1)] - pp[3 * m], dy = pp[3 * (m + 1) + 1] - pp[3 * m + 1],
                         dz = pp[3 * (m + 1) + 2] - pp[3 * m + 2];
            sum += sqrt(dx * dx + dy * dy + dz * dz);
        }
        const double ex = pp[3 * m] - g0, ey = pp[3 * m + 1] - g1, ez = pp[3 * m + 2] - g2;
        if (sqrt(__dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez))) >= goal_radius) last = m;
    }
    for (int o = 16; o; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
    }
    if (lane == 0) {
        dist_out[n] = sum;
        tidx_out[n] = last >= 0 ? last + 2 : 0;  // MATLAB: find(...,'last') is 1-based, time_index = hola + 1
    }
}

}  // namespace dmpc

// model_tables.cpp -- see model_tables.h.  Host code, runs once per handle.
#include "model_tables.h"

#include <cmath>
#include <cstring>

namespace dmpc {

// Scalar form of the reference recurrence (getPosMat.m:8-20, dmpc_soft_bound.m:92-108):
//   new_row = Aux*prev_row + add_b,  Aux = [I hI; 0 I]  =>  per axis  p <- p + h*v (+h^2/2 on the
//   diagonal block), v <- v (+h on the diagonal block), t <- t + h.
// Evaluated in that order with separately rounded multiply and add the result equals the
// reference's saved workspaces bit for bit (tests/test_model_mats.py), so keep the temporaries
// volatile: no FMA contraction whatever the compiler flags are.
static void scalar_model(double h, int K, std::vector<double>& lam, std::vector<double>& vel,
                         std::vector<double>& tt) {
    lam.assign((size_t)K * K, 0.0);
    vel.assign((size_t)K * K, 0.0);
    tt.assign(K, 0.0);
    std::vector<double> pc(K, 0.0), vc(K, 0.0);
    double t = 0.0;
    const double hh2 = h * h / 2;
    for (int k = 0; k < K; ++k) {
        for (int j = 0; j < K; ++j) {
            volatile double hv = h * vc[j];
            volatile double s = pc[j] + hv;
            pc[j] = s;
        }
        {
            volatile double s = pc[k] + hh2;
            pc[k] = s;
            volatile double w = vc[k] + h;
            vc[k] = w;
            volatile double tn = t + h;
            t = tn;
        }
        for (int j = 0; j < K; ++j) {
            lam[(size_t)k * K + j] = pc[j];
            vel[(size_t)k * K + j] = vc[j];
        }
        tt[k] = t;
    }
}

void model_mats(double h, int K, double* A_p, double* A_v, double* A_initp, double* Delta) {
    const size_t n = 3 * (size_t)K;
    std::vector<double> lam, vel, tt;
    scalar_model(h, K, lam, vel, tt);
    if (A_p) std::memset(A_p, 0, sizeof(double) * n * n);
    if (A_v) std::memset(A_v, 0, sizeof(double) * n * n);
    if (A_initp) std::memset(A_initp, 0, sizeof(double) * n * 6);
    if (Delta) std::memset(Delta, 0, sizeof(double) * n * n);
    for (int k = 0; k < K; ++k)
        for (int d = 0; d < 3; ++d) {
            const size_t row = 3 * (size_t)k + d;
            for (int j = 0; j < K; ++j) {
                const size_t col = 3 * (size_t)j + d;
                if (A_p) A_p[row + n * col] = lam[(size_t)k * K + j];
                if (A_v) A_v[row + n * col] = vel[(size_t)k * K + j];
            }
            if (A_initp) {
                A_initp[row + n * d] = 1.0;
                A_initp[row + n * (3 + d)] = tt[k];
            }
        }
    if (Delta) {
        for (size_t i = 0; i < n; ++i) Delta[i + n * i] = 1.0;
        for (size_t i = 3; i < n; ++i) Delta[i + n * (i - 3)] = -1.0;
    }
}

void build_tables(double h, int K, const double qs[3][2], std::vector<double>& out) {
    std::vector<double> lam, vel, tt;
    scalar_model(h, K, lam, vel, tt);
    out.assign(tables_size(K), 0.0);
    for (int i = 0; i < K * K; ++i) out[i] = lam[i];
    for (int k = 0; k < K; ++k) {
        out[K * K + k] = tt[k];
        long double s = 0;
        for (int j = 0; j < K; ++j) s += (long double)lam[k * K + j] * lam[k * K + j];
        out[K * K + K + k] = (double)sqrtl(s);
        out[K * K + 2 * K + k] = 1.0 / out[K * K + K + k];
    }
    typedef long double ld;
    std::vector<ld> Hm((size_t)K * K), L((size_t)K * K), Li((size_t)K * K), Gm((size_t)K * K),
        Bm((size_t)K * K);
    for (int w = 0; w < 3; ++w) {
        const ld q = qs[w][0], s = qs[w][1];
        // H_K = 2 (q lamK lamK' + s Delta'Delta + I), lamK = lam[K-1,:]   (spd = 1: Q on last block)
        for (int i = 0; i < K; ++i)
            for (int j = 0; j < K; ++j) {
                ld v = q * (ld)lam[(K - 1) * K + i] * (ld)lam[(K - 1) * K + j];
                if (i == j) v += s * ((i == K - 1) ? 1.0L : 2.0L) + 1.0L;
                if (i == j + 1 || j == i + 1) v -= s;
                Hm[i * K + j] = 2.0L * v;
            }
        // Cholesky H = L L'
        for (int j = 0; j < K; ++j) {
            ld d = Hm[j * K + j];
            for (int k = 0; k < j; ++k) d -= L[j * K + k] * L[j * K + k];
            d = sqrtl(d);
            L[j * K + j] = d;
            for (int i = j + 1; i < K; ++i) {
                ld v = Hm[i * K + j];
                for (int k = 0; k < j; ++k) v -= L[i * K + k] * L[j * K + k];
                L[i * K + j] = v / d;
            }
            for (int i = 0; i < j; ++i) L[i * K + j] = 0;
        }
        // Li = L^{-1} (lower), G = Li' Li
        for (int c = 0; c < K; ++c)
            for (int i = 0; i < K; ++i) {
                if (i < c) {
                    Li[i * K + c] = 0;
                    continue;
                }
                ld v = (i == c) ? 1.0L : 0.0L;
                for (int k = c; k < i; ++k) v -= L[i * K + k] * Li[k * K + c];
                Li[i * K + c] = v / L[i * K + i];
            }
        for (int i = 0; i < K; ++i)
            for (int j = 0; j < K; ++j) {
                ld v = 0;
                for (int k = (i > j ? i : j); k < K; ++k) v += Li[k * K + i] * Li[k * K + j];
                Gm[i * K + j] = v;
            }
        double* G = out.data() + tables_set_offset(K, w);
        double* B = G + K * K;
        double* C = B + K * K;
        for (int i = 0; i < K; ++i)
            for (int k = 0; k < K; ++k) {
                ld v = 0;  // B[i][k] = sum_j G[i][j] lam[k][j]
                for (int j = 0; j < K; ++j) v += Gm[i * K + j] * (ld)lam[k * K + j];
                Bm[i * K + k] = v;
            }
        for (int i = 0; i < K; ++i)
            for (int k = 0; k < K; ++k) {
                ld v = 0;  // C[i][k] = sum_j lam[i][j] B[j][k]
                for (int j = 0; j < K; ++j) v += (ld)lam[i * K + j] * Bm[j * K + k];
                G[i * K + k] = (double)Gm[i * K + k];
                B[i * K + k] = (double)Bm[i * K + k];
                C[i * K + k] = (double)v;
            }
        // symmetrise the rounded G and C exactly
        for (int i = 0; i < K; ++i)
            for (int k = 0; k < i; ++k) {
                G[i * K + k] = G[k * K + i];
                C[i * K + k] = C[k * K + i];
            }
    }
    // the shared-memory blob of the register-resident solver: header copy + interleaved T4 per weight set
    double* F = out.data() + tables_fast_offset(K);
    for (int i = 0; i < K * K + 4 * K; ++i) F[i] = out[i];
    for (int w = 0; w < 3; ++w) {
        const double* G = out.data() + tables_set_offset(K, w);
        const double* B = G + K * K;
        const double* C = B + K * K;
        double* T4 = F + tables_fast_header(K) + (size_t)w * 4 * K * K;
        for (int k = 0; k < K; ++k)
            for (int j = 0; j < K; ++j) {
                double* t = T4 + 4 * ((size_t)k * K + j);
                t[0] = G[k * K + j];
                t[1] = B[j * K + k];
                t[2] = B[k * K + j];
                t[3] = C[k * K + j];
            }
    }
}

}  // namespace dmpc

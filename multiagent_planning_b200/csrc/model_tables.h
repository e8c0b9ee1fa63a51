// model_tables.h -- host-side precompute of the DMPC model and of the constant tables the QP
// kernel looks its Schur-complement entries up in.
//
// Reference: getPosMat.m:1-23, getDeltaMat.m:1-9, dmpc_soft_bound.m:81-108 (A_p, A_v, A_initp),
// solveSoftDMPCbound.m:43-58,85-98 (weights, H).  The reference rebuilds the dense 3K x 3K Hessian
// H = 2(A'QA + Delta'S Delta + R) for every agent at every step; here H = kron(H_K, I_3) is never
// materialised: for each of the three (q,s) weight sets that can occur we keep
//      G = H_K^{-1},  B = G Lam',  C = Lam G Lam'      (all K x K)
// where Lam is the scalar (per-axis) acceleration->position map.
#pragma once
#include <vector>

namespace dmpc {

// flat table layout (doubles), K = horizon:
//   [0, K*K)            lam   row-major lam[k*K + j] = A_p(3k+d, 3j+d)
//   [K*K, K*K+K)        tt    tt[k] = A_initp(3k+d, 3+d) = (k+1) h (summed like the reference)
//   [.., +K)            lnorm lnorm[k] = || lam[k,:] ||_2
//   [.., +K)            ilnorm = 1 / lnorm
//   then for each weight set w = 0 (far), 1 (near), 2 (collision):  G, B, C  (K*K each, row-major)
//   [tables_fast_offset(K), +tables_fast_size(K))   the blob the register-resident solver (qp_warp.cuh)
//       stages in shared memory with ONE bulk copy: a copy of the header (lam, tt, lnorm, ilnorm, padded to a
//       multiple of 4 doubles) followed, per weight set, by the interleaved table
//           T4[k][j] = { G[k][j], B[j][k], B[k][j], C[k][j] }        (32-byte records, row-major in k, j)
//       so that two 128-bit loads fetch everything the direction needs for (k, j), and ONE fetches the pair
//       (acceleration part, position image) of H^-1 n for a normal in acceleration space (first half) or
//       in position space (second half).
inline int tables_set_offset(int K, int w) { return K * K + 4 * K + w * 3 * K * K; }
inline int tables_base_size(int K) { return K * K + 4 * K + 9 * K * K; }
inline int tables_fast_offset(int K) { return (tables_base_size(K) + 3) / 4 * 4; }
inline int tables_fast_header(int K) { return (K * K + 4 * K + 3) / 4 * 4; }
inline int tables_fast_size(int K) { return tables_fast_header(K) + 12 * K * K; }
inline int tables_size(int K) { return tables_fast_offset(K) + tables_fast_size(K); }

// qs[w] = {q, s} for w = far, near, collision
void build_tables(double h, int K, const double qs[3][2], std::vector<double>& out);

// dense reference-layout matrices (column-major), any pointer may be null
void model_mats(double h, int K, double* A_p, double* A_v, double* A_initp, double* Delta);

}  // namespace dmpc

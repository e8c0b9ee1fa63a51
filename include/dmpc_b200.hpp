// dmpc_b200.hpp -- header-only C++ facade over the C-ABI of libdmpc_b200.so with the public surface of the
// reference's `class DMPC` (dmpc/cpp/dmpc.h:70-182), so that a user of the C++ port (dmpc/cpp/main.cpp:16-73,
// cluster_test.cpp) can switch to the B200 path by changing an #include and the link line:
//
//     #include "dmpc_b200.hpp"            // instead of "dmpc.h"
//     using namespace dmpcb200;           // Params, Trajectory, DMPC
//     DMPC test("ooqp", p);  test.set_boundaries(pmin, pmax);  test.set_final_pts(pf);  test.set_initial_pts(po);
//     std::vector<Trajectory> sol = test.solveParallelDMPCv2();   test.trajectories2file(sol, "trajectories.txt");
//
// No Eigen: matrices are the small column-major `Mat` below (3 x cols, operator()(r, c), cols()).
// What differs from the reference, on purpose:
//   * the solver name is accepted and ignored (the QP arithmetic is the library's own exact dual active-set
//     solver on the GPU; eigen-quadprog / OOQP / CPLEX are not linked);
//   * set_cluster_num is accepted and ignored (agents are warps of one B200, not host threads);
//   * the post-checks of solveParallelDMPCv2 (dmpc.cpp:1716-1733) run the library's post-processing, which
//     restates the MATLAB pipeline (time scaling to vlim / alim, not-a-knot splines, pairwise check,
//     test/failure_rate.m:134-195) -- the C++ port interpolates with a Boost cubic B-spline and skips the time
//     scaling (`scale_solution` is commented out, dmpc.cpp:1720); `successful` means the same thing: no pair
//     closer than rmin - collision_tol on the interpolated trajectories;
//   * gen_rand_pts / gen_rand_perm take a seed (the reference seeds with srand(time(0)), dmpc.cpp:40: its inputs
//     are not reproducible); same rejection sampling and the same "everybody moves" permutation rule.
// Errors of the library surface as std::runtime_error with dmpcb200_last_error().
#ifndef DMPC_B200_HPP
#define DMPC_B200_HPP

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "dmpc_b200.h"

namespace dmpcb200 {

// rows x cols, column-major (element (r, c) at d[r + rows * c]) -- the layout of Eigen::MatrixXd and of MATLAB
struct Mat {
    int r = 0, c = 0;
    std::vector<double> d;
    Mat() {}
    Mat(int rows, int cols) : r(rows), c(cols), d((size_t)rows * cols, 0.0) {}
    double& operator()(int i, int j) { return d[(size_t)i + (size_t)r * j]; }
    double operator()(int i, int j) const { return d[(size_t)i + (size_t)r * j]; }
    int rows() const { return r; }
    int cols() const { return c; }
    double* data() { return d.data(); }
    const double* data() const { return d.data(); }
};
struct Vec3 {
    double v[3];
    Vec3(double x = 0, double y = 0, double z = 0) : v{x, y, z} {}
    double& operator()(int i) { return v[i]; }
    double operator()(int i) const { return v[i]; }
};

// dmpc.h:40-44
struct Trajectory {
    Mat pos, vel, acc;  // 3 x steps
};

// dmpc.h:50-63 (same fields, same order)
struct Params {
    float h;
    int T;
    int k_hor;
    int order;
    float c;
    float rmin;
    float alim;
    float vlim;
    int freq;
    float goal_tol;
    float collision_tol;
    int speed;
};
// dmpc.h:65-67
static const Params default_params = {0.2f, 10, 12, 2, 1.5f, 0.5f, 2.0f, 2.0f, 100, 0.05f, 0.05f, 1};

class DMPC {
public:
    explicit DMPC(std::string solver_name = "quadprog", Params params = default_params, int device = 0)
        : successful(false), _solver(std::move(solver_name)), _p(params), _device(device), _k_factor(0),
          _clusters(8), _h_scaled(params.h), _pmin(-2.5, -2.5, 0.2), _pmax(2.5, 2.5, 2.2) {
        if (params.order != 2) throw std::runtime_error("dmpcb200: only order = 2 (ellipsoid) is implemented");
        if (params.speed != 1) throw std::runtime_error("dmpcb200: only speed = 1 (terminal weight on the last step) is implemented");
    }

    std::vector<Trajectory> solution_short;  // solution before interpolation (dmpc.h:83)
    bool successful;                         // no collision after interpolation (dmpc.h:84)

    // ---- dmpc.cpp:188-227: rejection sampling of N points with pairwise (Euclidean) distance > rmin ----------
    Mat gen_rand_pts(int N, const Vec3& pmin, const Vec3& pmax, float rmin, uint64_t seed = 1) {
        std::mt19937_64 rng(seed);
        std::uniform_real_distribution<double> U(0.0, 1.0);
        Mat pts(3, N);
        for (int n = 0; n < N; ++n) {
            for (;;) {
                double cand[3];
                for (int x = 0; x < 3; ++x) cand[x] = pmin(x) + (pmax(x) - pmin(x)) * U(rng);
                bool pass = true;
                for (int k = 0; k < n && pass; ++k) {
                    const double dx = pts(0, k) - cand[0], dy = pts(1, k) - cand[1], dz = pts(2, k) - cand[2];
                    pass = std::sqrt(dx * dx + dy * dy + dz * dz) > rmin;
                }
                if (pass) {
                    for (int x = 0; x < 3; ++x) pts(x, n) = cand[x];
                    break;
                }
            }
        }
        return pts;
    }
    // ---- dmpc.cpp:229-265: random permutation of the start points in which every agent moves -------------------
    Mat gen_rand_perm(const Mat& po, uint64_t seed = 2) {
        const int N = po.cols();
        std::mt19937_64 rng(seed);
        std::vector<int> array(N), perm(N);
        for (int i = 0; i < N; ++i) array[i] = i;
        for (int i = 0; i < N; ++i) {
            std::vector<int> aux = array;
            aux.erase(std::remove(aux.begin(), aux.end(), i), aux.end());
            if (i == N - 1) {
                perm[i] = array.at(0);
            } else if (i == N - 2 && aux.back() == N - 1) {
                perm[i] = aux.back();
                array.erase(std::remove(array.begin(), array.end(), perm[i]), array.end());
            } else {
                const int j = (int)(rng() % (uint64_t)(N - i - 1));
                perm[i] = aux.at(j);
                array.erase(std::remove(array.begin(), array.end(), aux.at(j)), array.end());
            }
        }
        Mat pf(3, N);
        for (int i = 0; i < N; ++i)
            for (int x = 0; x < 3; ++x) pf(x, i) = po(x, perm[i]);
        return pf;
    }

    // ---- setters (dmpc.h:126-144); out-of-bounds points are moved onto the boundary like dmpc.cpp:266-360 ------
    void set_boundaries(const Vec3& pmin, const Vec3& pmax) { _pmin = pmin; _pmax = pmax; }
    void set_initial_pts(const Mat& po) { _po = clip(po); }
    void set_final_pts(const Mat& pf) { _pf = clip(pf); }
    void set_k_factor(int k_factor) {
        if (k_factor != 0 && k_factor != -1) throw std::runtime_error("dmpcb200: k_factor must be 0 or -1 (dmpc.cpp:516)");
        _k_factor = k_factor;
    }
    void set_cluster_num(int num) { _clusters = num; }

    // ---- the solves.  solveDMPC / solveParallelDMPC (fixed-length runs of the older solveQP) map onto the same
    //      batched GPU step; solveParallelDMPCv2 (dmpc.cpp:1570-1740) is the one the reference's drivers call ----
    std::vector<Trajectory> solveParallelDMPCv2() { return solve(true); }
    std::vector<Trajectory> solveParallelDMPC() { return solve(false); }
    std::vector<Trajectory> solveDMPC() { return solve(false); }

    // ---- dmpc.cpp:2088-2126 -----------------------------------------------------------------------------------
    void trajectories2file(const std::vector<Trajectory>& src, char const* pathAndName) {
        const int N = _po.cols(), N_cmd = (int)src.size();
        if (!N_cmd) throw std::runtime_error("dmpcb200: trajectories2file: empty solution");
        const int T = src[0].pos.cols();
        std::vector<double> pos((size_t)3 * T * N_cmd), vel(pos.size()), acc(pos.size());
        for (int i = 0; i < N_cmd; ++i) {
            std::copy(src[i].pos.d.begin(), src[i].pos.d.end(), pos.begin() + (size_t)3 * T * i);
            std::copy(src[i].vel.d.begin(), src[i].vel.d.end(), vel.begin() + (size_t)3 * T * i);
            std::copy(src[i].acc.d.begin(), src[i].acc.d.end(), acc.begin() + (size_t)3 * T * i);
        }
        if (dmpcb200_write_trajectories(pathAndName, N, N_cmd, T, _h_scaled, _pmin.v, _pmax.v, _po.data(), _pf.data(),
                                        pos.data(), vel.data(), acc.data()))
            throw std::runtime_error(std::string("dmpcb200: cannot write ") + pathAndName);
    }

    // figures of the last solve
    int steps() const { return _steps; }
    bool reached_goal() const { return _reached; }
    double min_distance() const { return _min_dist; }
    double trajectory_time() const { return _traj_time; }

private:
    std::string _solver;
    Params _p;
    int _device, _k_factor, _clusters;
    double _h_scaled;
    Vec3 _pmin, _pmax;
    Mat _po, _pf;
    int _steps = 0;
    bool _reached = false;
    double _min_dist = 0.0, _traj_time = 0.0;

    Mat clip(const Mat& m) const {
        Mat o = m;
        for (int j = 0; j < o.cols(); ++j)
            for (int x = 0; x < 3; ++x) o(x, j) = std::min(std::max(o(x, j), _pmin(x)), _pmax(x));
        return o;
    }
    static void ck(int rc, const char* what) {
        if (rc) throw std::runtime_error(std::string("dmpcb200: ") + what + ": " + dmpcb200_last_error());
    }

    std::vector<Trajectory> solve(bool stop_at_goal) {
        const int N = _po.cols(), N_cmd = _pf.cols();
        if (N < 1 || N_cmd < 1 || N_cmd > N) throw std::runtime_error("dmpcb200: set_initial_pts / set_final_pts first (N_cmd <= N)");
        dmpcb200_params q;
        dmpcb200_default_params_cpp(&q, _k_factor);
        q.h = _p.h; q.K = _p.k_hor; q.c = _p.c; q.rmin = _p.rmin; q.alim = _p.alim;
        q.goal_tol = stop_at_goal ? _p.goal_tol : -1.0;  // (a negative tolerance is never met: fixed-length run)
        q.coll_tol = _p.collision_tol;
        dmpcb200_t* h = nullptr;
        ck(dmpcb200_create(&q, N, 0, N, 1, _device, 0, &h), "create");
        struct Guard { dmpcb200_t* h; ~Guard() { dmpcb200_destroy(h); } } guard{h};
        ck(dmpcb200_set_bounds(h, _pmin.v, _pmax.v), "set_bounds");
        Mat pf_all(3, N);  // un-commanded agents: static obstacles (dmpc.cpp:1633-1649)
        for (int j = 0; j < N; ++j)
            for (int x = 0; x < 3; ++x) pf_all(x, j) = j < N_cmd ? _pf(x, j) : _po(x, j);
        ck(dmpcb200_set_goals(h, pf_all.data()), "set_goals");
        if (N_cmd < N) ck(dmpcb200_set_static_obstacles(h, N_cmd), "set_static_obstacles");
        ck(dmpcb200_init_horizons(h, _po.data(), nullptr, nullptr, nullptr, nullptr), "init_horizons");
        const int K = (int)(_p.T / _p.h);  // dmpc.cpp:52: _K = T / h
        const int S = K - 1;
        std::vector<double> tp((size_t)3 * (S + 1) * N), tv(tp.size()), ta(tp.size());
        int32_t steps = 0, reached = 0, fs = -1, fa = -1;
        ck(dmpcb200_run(h, S, /*stop_on_fail*/ 1, 0, tp.data(), tv.data(), ta.data(), nullptr, &steps, &reached, &fs, &fa),
           "run");
        _steps = steps;
        _reached = reached != 0;
        const int T = steps + 1;
        solution_short.clear();
        for (int i = 0; i < N_cmd; ++i) {
            Trajectory t;
            t.pos = Mat(3, T); t.vel = Mat(3, T); t.acc = Mat(3, T);
            const size_t o = (size_t)3 * (S + 1) * i;
            std::copy(tp.begin() + o, tp.begin() + o + 3 * T, t.pos.d.begin());
            std::copy(tv.begin() + o, tv.begin() + o + 3 * T, t.vel.d.begin());
            std::copy(ta.begin() + o, ta.begin() + o + 3 * T, t.acc.d.begin());
            solution_short.push_back(std::move(t));
        }
        successful = false;
        _h_scaled = _p.h;
        std::vector<Trajectory> solution;
        if (_reached && fs < 0 && T >= 4 && N_cmd == N) {
            // post-checks (dmpc.cpp:1716-1733) through the library's post-processing
            std::vector<double> pk((size_t)3 * T * N), vk(pk.size()), ak(pk.size());
            for (int i = 0; i < N; ++i) {
                std::copy(solution_short[i].pos.d.begin(), solution_short[i].pos.d.end(), pk.begin() + (size_t)3 * T * i);
                std::copy(solution_short[i].vel.d.begin(), solution_short[i].vel.d.end(), vk.begin() + (size_t)3 * T * i);
                std::copy(solution_short[i].acc.d.begin(), solution_short[i].acc.d.end(), ak.begin() + (size_t)3 * T * i);
            }
            dmpcb200_post res;
            std::vector<double> scratch_p(pk), scratch_v(vk), scratch_a(ak);
            ck(dmpcb200_postprocess(h, T, scratch_p.data(), scratch_v.data(), scratch_a.data(), _p.vlim, _p.alim,
                                    1.0 / _p.freq, 0.05, nullptr, nullptr, nullptr, 0, nullptr, &res), "postprocess");
            const int nt = res.nt;
            std::vector<double> ip((size_t)3 * nt * N), iv(ip.size()), ia(ip.size());
            ck(dmpcb200_postprocess(h, T, pk.data(), vk.data(), ak.data(), _p.vlim, _p.alim, 1.0 / _p.freq, 0.05,
                                    ip.data(), iv.data(), ia.data(), nt, nullptr, &res), "postprocess");
            successful = res.violation == 0;
            _h_scaled = res.h_scaled;
            _min_dist = res.min_dist;
            _traj_time = res.traj_time;
            for (int i = 0; i < N; ++i) {
                Trajectory t;
                t.pos = Mat(3, nt); t.vel = Mat(3, nt); t.acc = Mat(3, nt);
                const size_t o = (size_t)3 * nt * i;
                std::copy(ip.begin() + o, ip.begin() + o + 3 * nt, t.pos.d.begin());
                std::copy(iv.begin() + o, iv.begin() + o + 3 * nt, t.vel.d.begin());
                std::copy(ia.begin() + o, ia.begin() + o + 3 * nt, t.acc.d.begin());
                solution.push_back(std::move(t));
            }
        }
        return solution;
    }
};

}  // namespace dmpcb200
#endif

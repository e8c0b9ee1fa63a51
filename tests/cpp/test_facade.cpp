// Compiled test of include/dmpc_b200.hpp: the reference's dmpc/cpp/main.cpp:16-73 shape (N random agents in a
// box, solveParallelDMPCv2, trajectories2file) on the facade.
//   test_facade nogpu <out.txt>        host-only parts (generators, setters, file format) + the loud failure
//                                      of the solve without a CUDA device
//   test_facade gpu <out.txt> <N> <N_cmd> <k_factor>
//                                      the whole thing; prints one line of figures for the Python side
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "dmpc_b200.hpp"

using namespace dmpcb200;

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    const bool gpu = !std::strcmp(argv[1], "gpu");
    const char* out = argv[2];
    const int N = argc > 3 ? std::atoi(argv[3]) : 20;
    const int N_cmd = argc > 4 ? std::atoi(argv[4]) : N;
    const int k_factor = argc > 5 ? std::atoi(argv[5]) : 0;
    Params p = {0.2f, 20, 15, 2, 2.0f, 0.35f, 1.0f, 2.0f, 100, 0.01f, 0.05f, 1};  // main.cpp:19-31 values
    DMPC test("ooqp", p);
    const Vec3 pmin(-2.5, -2.5, 0.2), pmax(2.5, 2.5, 2.2);
    test.set_boundaries(pmin, pmax);
    Mat po = test.gen_rand_pts(N, pmin, pmax, p.rmin + 0.2f, 11);
    Mat pf_all = test.gen_rand_perm(po, 12);
    // every agent moves, every goal is a start point
    for (int i = 0; i < N; ++i) {
        bool same = true;
        for (int x = 0; x < 3; ++x) same = same && po(x, i) == pf_all(x, i);
        if (same) { std::printf("FAIL: agent %d does not move\n", i); return 1; }
    }
    for (int i = 0; i < N; ++i)
        for (int j = i + 1; j < N; ++j) {
            double d2 = 0;
            for (int x = 0; x < 3; ++x) d2 += (po(x, i) - po(x, j)) * (po(x, i) - po(x, j));
            if (!(std::sqrt(d2) > p.rmin + 0.2f)) { std::printf("FAIL: start points too close\n"); return 1; }
        }
    Mat pf(3, N_cmd);
    for (int i = 0; i < N_cmd; ++i)
        for (int x = 0; x < 3; ++x) pf(x, i) = pf_all(x, i);
    test.set_final_pts(pf);
    test.set_initial_pts(po);
    test.set_k_factor(k_factor);
    test.set_cluster_num(8);
    try {
        test.set_k_factor(3);
        std::printf("FAIL: k_factor 3 accepted\n");
        return 1;
    } catch (const std::runtime_error&) {
    }
    if (!gpu) {
        // the file format needs no device
        std::vector<Trajectory> fake;
        for (int i = 0; i < N_cmd; ++i) {
            Trajectory t;
            t.pos = Mat(3, 5); t.vel = Mat(3, 5); t.acc = Mat(3, 5);
            for (int k = 0; k < 5; ++k)
                for (int x = 0; x < 3; ++x) { t.pos(x, k) = po(x, i) + 0.1 * k; t.vel(x, k) = 0.5 * k; t.acc(x, k) = -1.0 / (k + 1); }
            fake.push_back(t);
        }
        test.trajectories2file(fake, out);
        try {
            test.solveParallelDMPCv2();
            std::printf("FAIL: solve without a CUDA device did not throw\n");
            return 1;
        } catch (const std::runtime_error& e) {
            std::printf("ok nogpu: %s\n", e.what());
            return 0;
        }
    }
    std::vector<Trajectory> sol = test.solveParallelDMPCv2();
    test.trajectories2file(test.solution_short, out);
    std::printf("ok gpu: N=%d N_cmd=%d steps=%d reached=%d successful=%d interp_cols=%d min_dist=%.6f traj_time=%.3f\n", N,
                N_cmd, test.steps(), (int)test.reached_goal(), (int)test.successful, sol.empty() ? 0 : sol[0].pos.cols(),
                test.min_distance(), test.trajectory_time());
    return 0;
}

// traj_io.cpp -- the reference's on-disk trajectory format (host code, no CUDA).
//
// DMPC::trajectories2file (dmpc/cpp/dmpc.cpp:2088-2126) dumps a solved transition as text that
// dmpc/cpp_results/read_result.m:1-42 reads back with dlmread:
//     line 1:  N  N_cmd  h_scaled  pmin(3)  pmax(3)
//     3 lines: po (3 x N), 3 lines: pf (3 x N_cmd)
//     then for every commanded agent its positions (3 lines of T numbers), then all velocities, then all
//     accelerations.
// Every matrix goes through Eigen's default operator<< : 6 significant digits ("%.6g"), coefficients
// right-aligned to the widest one OF THAT MATRIX, one blank between columns, one line per row.  The writer
// below reproduces that byte for byte (checked against the reference's own dump of a 200-agent transition).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/dmpc_b200.h"

namespace {

// Eigen default IOFormat of a rows x cols matrix (element (r,c) at data[r*rs + c*cs])
void eigen_format(std::string& out, int rows, int cols, const double* data, size_t rs, size_t cs) {
    std::vector<std::string> cell((size_t)rows * cols);
    size_t width = 0;
    char buf[64];
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) {
            std::snprintf(buf, sizeof buf, "%.6g", data[r * rs + c * cs]);
            cell[(size_t)r * cols + c] = buf;
            width = std::max(width, std::strlen(buf));
        }
    for (int r = 0; r < rows; ++r) {
        if (r) out += '\n';
        for (int c = 0; c < cols; ++c) {
            if (c) out += ' ';
            const std::string& s = cell[(size_t)r * cols + c];
            out.append(width - s.size(), ' ');
            out += s;
        }
    }
}

}  // namespace

extern "C" {

int dmpcb200_format_matrix(int rows, int cols, const double* col_major, char* buf, int cap) {
    if (rows < 1 || cols < 1 || !col_major) return DMPCB200_ERR_ARG;
    std::string s;
    eigen_format(s, rows, cols, col_major, 1, (size_t)rows);
    if (buf && cap > 0) {
        const size_t n = std::min((size_t)cap - 1, s.size());
        std::memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return (int)s.size();
}

int dmpcb200_write_trajectories(const char* path, int N, int N_cmd, int T, double h_scaled, const double* pmin,
                                const double* pmax, const double* po, const double* pf, const double* pos,
                                const double* vel, const double* acc) {
    if (!path || !pmin || !pmax || !po || !pf || !pos || !vel || !acc || N < 1 || N_cmd < 1 || N_cmd > N || T < 1)
        return DMPCB200_ERR_ARG;
    std::string s;
    char buf[64];
    std::snprintf(buf, sizeof buf, "%d %d %.6g ", N, N_cmd, h_scaled);
    s += buf;
    eigen_format(s, 1, 3, pmin, 0, 1);  // _pmin.transpose()
    s += ' ';
    eigen_format(s, 1, 3, pmax, 0, 1);
    s += '\n';
    eigen_format(s, 3, N, po, 1, 3);
    s += '\n';
    eigen_format(s, 3, N_cmd, pf, 1, 3);
    s += '\n';
    const double* blocks[3] = {pos, vel, acc};
    for (const double* b : blocks)
        for (int i = 0; i < N_cmd; ++i) {
            eigen_format(s, 3, T, b + (size_t)3 * T * i, 1, 3);
            s += '\n';
        }
    FILE* f = std::fopen(path, "w");
    if (!f) return DMPCB200_ERR_STATE;
    const bool ok = std::fwrite(s.data(), 1, s.size(), f) == s.size();
    return (std::fclose(f) == 0 && ok) ? 0 : DMPCB200_ERR_STATE;
}

int dmpcb200_read_trajectories(const char* path, int32_t* N, int32_t* N_cmd, int32_t* T, double* h_scaled,
                               double* pmin, double* pmax, double* po, double* pf, double* pos, double* vel,
                               double* acc) {
    if (!path || !N || !N_cmd || !T) return DMPCB200_ERR_ARG;
    FILE* f = std::fopen(path, "r");
    if (!f) return DMPCB200_ERR_STATE;
    std::vector<std::vector<double>> rows;
    std::string line;
    int ch;
    auto flush = [&]() {
        std::vector<double> v;
        const char* p = line.c_str();
        char* end = nullptr;
        for (;;) {
            const double x = std::strtod(p, &end);
            if (end == p) break;
            v.push_back(x);
            p = end;
        }
        if (!v.empty()) rows.push_back(std::move(v));
        line.clear();
    };
    while ((ch = std::fgetc(f)) != EOF) {
        if (ch == '\n') flush();
        else line += (char)ch;
    }
    flush();
    std::fclose(f);
    if (rows.size() < 7 || rows[0].size() < 9) return DMPCB200_ERR_ARG;
    const int n = (int)rows[0][0], nc = (int)rows[0][1];
    if (n < 1 || nc < 1 || nc > n || rows.size() != (size_t)7 + 9 * (size_t)nc) return DMPCB200_ERR_ARG;
    const int t = (int)rows[7].size();
    *N = n;
    *N_cmd = nc;
    *T = t;
    if (h_scaled) *h_scaled = rows[0][2];
    for (int x = 0; x < 3; ++x) {
        if (pmin) pmin[x] = rows[0][3 + x];
        if (pmax) pmax[x] = rows[0][6 + x];
    }
    for (int x = 0; x < 3; ++x) {
        if ((int)rows[1 + x].size() != n || (int)rows[4 + x].size() != nc) return DMPCB200_ERR_ARG;
        if (po) for (int i = 0; i < n; ++i) po[3 * i + x] = rows[1 + x][i];
        if (pf) for (int i = 0; i < nc; ++i) pf[3 * i + x] = rows[4 + x][i];
    }
    double* blocks[3] = {pos, vel, acc};
    for (int b = 0; b < 3; ++b)
        for (int i = 0; i < nc; ++i)
            for (int x = 0; x < 3; ++x) {
                const std::vector<double>& r = rows[7 + (size_t)3 * nc * b + 3 * i + x];
                if ((int)r.size() != t) return DMPCB200_ERR_ARG;
                if (blocks[b])
                    for (int k = 0; k < t; ++k) blocks[b][3 * ((size_t)k + (size_t)t * i) + x] = r[k];
            }
    return 0;
}

}  // extern "C"

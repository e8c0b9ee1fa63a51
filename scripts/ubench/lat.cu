// micro-benchmarks: dependent-chain latency and per-warp throughput of the instructions the QP kernel
// lives on (DFMA, DADD, LDS, SHFL, REDUX, MUFU.RCP64H).  nvcc -arch=sm_100a -O3 lat.cu -o lat
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double* out, long long* cyc, int iters, int ilp_dummy) {
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 1.0 + i * 1e-9;
    __syncthreads();
    double a = out[threadIdx.x], b = 1.0000001, c = 1e-9;
    double a2 = a + 1, a3 = a + 2, a4 = a + 3;
    int idx = threadIdx.x & 31;
    unsigned u = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) { a = fma(a, b, c); }                                   // DFMA latency
        if (MODE == 1) { a = fma(a, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c); a4 = fma(a4, b, c); }  // 4 indep
        if (MODE == 2) { a = a + c; }                                          // DADD latency
        if (MODE == 3) { idx = (int)sm[idx & 1023 ] ; idx += threadIdx.x & 31; }   // LDS + cvt dependent
        if (MODE == 4) { a = __shfl_xor_sync(0xffffffffu, a, 1); }             // SHFL 64-bit
        if (MODE == 5) { u = __reduce_max_sync(0xffffffffu, u) + 1; }          // REDUX
        if (MODE == 6) { a = __drcp_rn(a) + 1.5; }                             // RCP64 + add
        if (MODE == 7) { a = 1.0 / a + 1.5; }                                  // full division
        if (MODE == 8) { a = sm[(i * 33 + threadIdx.x) & 1023] + a; }          // LDS independent address + DADD
        if (MODE == 9) { a = sqrt(a) + 1.5; }
    }
    long long t1 = clock64();
    out[threadIdx.x] = a + a2 + a3 + a4 + idx + u;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, int threads) {
    double* out; long long* cyc;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8 * 8);
    cudaMemset(out, 0, 1024 * 8);
    const int iters = 4096;
    k<MODE><<<1, threads>>>(out, cyc, iters, 0);
    k<MODE><<<1, threads>>>(out, cyc, iters, 0);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s threads %4d : %7.1f cycles / iteration\n", name, threads, (double)h / iters);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int th : {32, 128, 512}) {
        run<0>("DFMA dependent", th);
        run<1>("DFMA x4 independent", th);
        run<2>("DADD dependent", th);
        run<3>("LDS->cvt dependent", th);
        run<4>("SHFL.64 dependent", th);
        run<5>("REDUX.MAX dependent", th);
        run<6>("drcp_rn + add", th);
        run<7>("ddiv + add", th);
        run<8>("LDS + DADD", th);
        run<9>("dsqrt + add", th);
    }
    return 0;
}

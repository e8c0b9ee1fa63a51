// launch.cuh -- host-side launchers of the two template kernels of a step.  Every instantiation of
// qp_kernel / scan_kernel lives in its own translation unit (k_*.cu) so that the library builds in
// parallel; dmpc_b200.cu only sees the declarations below.  (-DDMPC_SINGLE_TU: the profiling build
// includes the k_*.cu files into dmpc_b200.cu, because its cycle counters are one __device__ array.)
#pragma once
#include <algorithm>

#include "dmpc_kernels.cuh"

namespace dmpc {

constexpr int kMaxDevices = 64;
inline int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}
inline int sm_count(int dev) {
    static int n_sm[kMaxDevices] = {0};
    if (!n_sm[dev]) {
        cudaDeviceGetAttribute(&n_sm[dev], cudaDevAttrMultiProcessorCount, dev);
        if (n_sm[dev] < 1) n_sm[dev] = 148;
    }
    return n_sm[dev];
}

// ---- declarations (defined in k_qp15.cu, k_qp20.cu, k_qpgen.cu, k_scan.cu) ------------------------
cudaError_t launch_qp_4_15(const StepArgs& A, int nl, size_t smem, cudaStream_t s);
cudaError_t launch_qp_4_20(const StepArgs& A, int nl, size_t smem, cudaStream_t s);
cudaError_t launch_qp_4_0(const StepArgs& A, int nl, size_t smem, cudaStream_t s);
cudaError_t launch_qp_3_0(const StepArgs& A, int nl, size_t smem, cudaStream_t s);
// solveHardDMPC (k_qphard.cu): K = 15 compiled in, any other horizon generic
cudaError_t launch_qp_hard_4_15(const StepArgs& A, int nl, size_t smem, cudaStream_t s);
cudaError_t launch_qp_hard_4_0(const StepArgs& A, int nl, size_t smem, cudaStream_t s);
cudaError_t launch_qp_hard_3_0(const StepArgs& A, int nl, size_t smem, cudaStream_t s);
// throughput layout (two-role persistent kernel, 8 light / 4 heavy agents per SM), horizons 15 and 20
cudaError_t launch_qp2_15(const StepArgs& A, int nl, cudaStream_t s);
cudaError_t launch_qp2_20(const StepArgs& A, int nl, cudaStream_t s);
// scan layouts: W agents per CTA x S warps per agent, horizon KT (0: run time)
enum ScanLayout { SCAN_1_2_0, SCAN_4_2_15, SCAN_4_2_20, SCAN_4_4_0, SCAN_8_1_15, SCAN_8_1_20, SCAN_8_2_0,
                  SCAN_RT_4_8_15, SCAN_RT_4_8_20, SCAN_RT_8_8_15, SCAN_RT_8_8_20, SCAN_RT_14_8_15,  // register-tile layout (scan_rt_kernel)
                  SCAN_LAYOUTS };
cudaError_t launch_scan_layout(ScanLayout id, const StepArgs& A, int nl, int K, cudaStream_t s);

// ---- templates (instantiated by the k_*.cu files only) ---------------------------------------------
#if defined(DMPC_LAUNCH_IMPL)
template <int W, int S, int KT, bool OWNREG = (KT > 0)>
cudaError_t launch_scan_w(const StepArgs& A, int nl, int K, cudaStream_t s) {
    const int Npad = round_up(A.P.N, kTile);
    int stages = scan_stages(K, A.P.N, W, A.RMAX);
    if (stages < 1) return cudaErrorInvalidConfiguration;
    if (stages > S) stages = stages / S * S;  // rounds of S tiles map onto distinct stages
    const size_t smem = scan_smem_bytes(K, W, stages, Npad, A.RMAX);
    // the opt-in above 48 KB is a per-device (per-context) attribute: one cache entry per device
    static size_t attr_smem[kMaxDevices] = {0};
    const int dev = current_device();
    if (attr_smem[dev] < smem) {
        cudaError_t e = cudaFuncSetAttribute(scan_kernel<W, S, KT, OWNREG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_smem[dev] = smem;
    }
    scan_kernel<W, S, KT, OWNREG><<<(nl + W - 1) / W * A.n_scen, W * S * 32, smem, s>>>(A, stages);
    return cudaGetLastError();
}
template <int W, int NW, int KT>
cudaError_t launch_scan_rt_w(const StepArgs& A, int nl, int K, cudaStream_t s) {
    const int Npad = round_up(A.P.N, kTile);
    const int stages = scan_stages(K, A.P.N, W, A.RMAX);
    if (stages < 1) return cudaErrorInvalidConfiguration;
    const size_t smem = scan_smem_bytes(K, W, stages, Npad, A.RMAX);
    static size_t attr_smem[kMaxDevices] = {0};
    const int dev = current_device();
    if (attr_smem[dev] < smem) {
        cudaError_t e = cudaFuncSetAttribute(scan_rt_kernel<W, NW, KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_smem[dev] = smem;
    }
    scan_rt_kernel<W, NW, KT><<<(nl + W - 1) / W * A.n_scen, NW * 32, smem, s>>>(A, stages);
    return cudaGetLastError();
}
template <int W, int KT, bool HARD = false>
cudaError_t launch_qp_w(const StepArgs& A, int nl, size_t smem, cudaStream_t s) {
    static size_t attr_smem[kMaxDevices] = {0};  // per device, like the SM count below
    const int dev = current_device();
    if (attr_smem[dev] < smem) {  // (the kernel also has a few hundred bytes of static shared memory)
        cudaError_t e =
            cudaFuncSetAttribute(qp_kernel<W, KT, HARD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_smem[dev] = smem;
    }
    // one CTA per SM at most (shared memory): larger swarms run a persistent grid with an agent queue
    const int ctas = std::min((nl * A.n_scen + W - 1) / W, sm_count(dev));
    // programmatic stream serialization: the grid may launch while the scan kernel drains (the kernel
    // itself waits for the scan's completion before it reads the rows)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)ctas);
    cfg.blockDim = dim3(W * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, qp_kernel<W, KT, HARD>, A);
}
template <int KT>
cudaError_t launch_qp2_w(const StepArgs& A, int nl, cudaStream_t s) {
    static size_t attr_smem[kMaxDevices] = {0};
    const int dev = current_device();
    const size_t smem = qp2_smem_bytes(KT, A.QMAX, A.RCAP);
    if (attr_smem[dev] < smem) {
        cudaError_t e = cudaFuncSetAttribute(qp2_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_smem[dev] = smem;
    }
    // one persistent CTA per SM (never more: a CTA in its drain phase waits for the light phases of all others)
    const int ctas = std::min((nl * A.n_scen + kLightW - 1) / kLightW, sm_count(dev));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)ctas);
    cfg.blockDim = dim3(kLightW * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, qp2_kernel<KT>, A);
}
#endif

}  // namespace dmpc

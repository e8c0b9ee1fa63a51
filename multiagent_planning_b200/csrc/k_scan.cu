// k_scan.cu -- one translation unit of libdmpc_b200.so: kernel instantiation(s) + launcher (launch.cuh)
#define DMPC_LAUNCH_IMPL
#include "launch.cuh"

namespace dmpc {
cudaError_t launch_scan_layout(ScanLayout id, const StepArgs& A, int nl, int K, cudaStream_t s) {
    switch (id) {
        case SCAN_1_2_0: return launch_scan_w<1, 2, 0>(A, nl, K, s);
        case SCAN_4_2_15: return launch_scan_w<4, 2, 15>(A, nl, K, s);
        case SCAN_4_2_20: return launch_scan_w<4, 2, 20>(A, nl, K, s);
        case SCAN_4_4_0: return launch_scan_w<4, 4, 0>(A, nl, K, s);
        case SCAN_8_1_15: return launch_scan_w<8, 1, 15>(A, nl, K, s);
        case SCAN_8_1_20: return launch_scan_w<8, 1, 20>(A, nl, K, s);
        case SCAN_RT_4_8_15: return launch_scan_rt_w<4, 8, 15>(A, nl, K, s);
        case SCAN_RT_4_8_20: return launch_scan_rt_w<4, 8, 20>(A, nl, K, s);
        case SCAN_RT_8_8_15: return launch_scan_rt_w<8, 8, 15>(A, nl, K, s);
        case SCAN_RT_8_8_20: return launch_scan_rt_w<8, 8, 20>(A, nl, K, s);
        case SCAN_RT_14_8_15: return launch_scan_rt_w<14, 8, 15>(A, nl, K, s);
        default: return launch_scan_w<8, 2, 0>(A, nl, K, s);
    }
}
}  // namespace dmpc

// k_qp20.cu -- one translation unit of libdmpc_b200.so: kernel instantiation(s) + launcher (launch.cuh)
#define DMPC_LAUNCH_IMPL
#include "launch.cuh"

namespace dmpc {
cudaError_t launch_qp_4_20(const StepArgs& A, int nl, size_t smem, cudaStream_t s) { return launch_qp_w<4, 20>(A, nl, smem, s); }
}  // namespace dmpc

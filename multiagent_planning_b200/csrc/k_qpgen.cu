// k_qpgen.cu -- one translation unit of libdmpc_b200.so: kernel instantiation(s) + launcher (launch.cuh)
#define DMPC_LAUNCH_IMPL
#include "launch.cuh"

namespace dmpc {
// W = 4 for every horizon up to 21, 3 beyond (the tables grow with K^2)
cudaError_t launch_qp_4_0(const StepArgs& A, int nl, size_t smem, cudaStream_t s) { return launch_qp_w<4, 0>(A, nl, smem, s); }
cudaError_t launch_qp_3_0(const StepArgs& A, int nl, size_t smem, cudaStream_t s) { return launch_qp_w<3, 0>(A, nl, smem, s); }
}  // namespace dmpc

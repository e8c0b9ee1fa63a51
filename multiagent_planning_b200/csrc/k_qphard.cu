// k_qphard.cu -- one translation unit of libdmpc_b200.so: the QP kernels of solveHardDMPC (rows on several horizon
// indices: the MK instantiation of the register-resident solver), classic layout
#define DMPC_LAUNCH_IMPL
#include "launch.cuh"

namespace dmpc {
cudaError_t launch_qp_hard_4_15(const StepArgs& A, int nl, size_t smem, cudaStream_t s) { return launch_qp_w<4, 15, true>(A, nl, smem, s); }
cudaError_t launch_qp_hard_4_0(const StepArgs& A, int nl, size_t smem, cudaStream_t s) { return launch_qp_w<4, 0, true>(A, nl, smem, s); }
cudaError_t launch_qp_hard_3_0(const StepArgs& A, int nl, size_t smem, cudaStream_t s) { return launch_qp_w<3, 0, true>(A, nl, smem, s); }
}  // namespace dmpc

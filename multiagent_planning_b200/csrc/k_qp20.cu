// k_qp20.cu -- one translation unit of libdmpc_b200.so: the QP kernels of horizon K = 20, classic layout
// (one agent per SM sub-partition) and throughput layout (two-role persistent kernel).  Both live HERE so that
// they share one compiled body of the per-agent solver (qp_agent is not inlined): an agent's bits do not
// depend on the layout that solved it.
#define DMPC_LAUNCH_IMPL
#include "launch.cuh"

namespace dmpc {
cudaError_t launch_qp_4_20(const StepArgs& A, int nl, size_t smem, cudaStream_t s) { return launch_qp_w<4, 20>(A, nl, smem, s); }
cudaError_t launch_qp2_20(const StepArgs& A, int nl, cudaStream_t s) { return launch_qp2_w<20>(A, nl, s); }
}  // namespace dmpc

// Instantiates the shared-memory lane-per-tree DPW kernel (dpw_fast.cuh) for SimpleGridWorld, one variant per
// rollout-policy kind.
#include "dpw_fast.cuh"

namespace mctsb200 {

template <class M, int POLICY>
static int32_t launch_dpw_fast_one(const KernelArgs<M>& args, int grid, int block, size_t smem, cudaStream_t stream) {
    auto kernel = dpw_fast_kernel<M, POLICY>;
#ifndef MCTSB200_HOST_EMU
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return MCTSB200_ERR_CUDA;
    prefer_smallest_carveout(kernel, smem);
#endif
    MCTSB200_LAUNCH(kernel, (unsigned)grid, (unsigned)block, smem, stream, args);
    return MCTSB200_OK;
}

template <>
int32_t launch_dpw_fast<GridWorldDev>(const KernelArgs<GridWorldDev>& args, int policy, int grid, int block, size_t smem, cudaStream_t stream) {
    switch (policy) {
        case FAST_POLICY_RANDOM: return launch_dpw_fast_one<GridWorldDev, FAST_POLICY_RANDOM>(args, grid, block, smem, stream);
        case FAST_POLICY_CONSTANT: return launch_dpw_fast_one<GridWorldDev, FAST_POLICY_CONSTANT>(args, grid, block, smem, stream);
        default: return launch_dpw_fast_one<GridWorldDev, FAST_POLICY_TABLE>(args, grid, block, smem, stream);
    }
}

}  // namespace mctsb200

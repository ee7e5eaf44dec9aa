// Instantiates the shared-memory lane-per-tree UCT kernel (uct_fast.cuh) for SimpleGridWorld: one variant per
// rollout-policy kind and descent mode.
#include "uct_fast.cuh"

namespace mctsb200 {

template <class M, int POLICY, int MODE>
static int32_t launch_fast_one(const KernelArgs<M>& args, int grid, int block, size_t smem, cudaStream_t stream) {
    auto kernel = uct_fast_kernel<M, POLICY, MODE>;
#ifndef MCTSB200_HOST_EMU
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return MCTSB200_ERR_CUDA;
    prefer_smallest_carveout(kernel, smem);
#endif
    MCTSB200_LAUNCH(kernel, (unsigned)grid, (unsigned)block, smem, stream, args);
    return MCTSB200_OK;
}

template <class M, int MODE>
static int32_t launch_fast_policy(const KernelArgs<M>& args, int policy, int grid, int block, size_t smem, cudaStream_t stream) {
    switch (policy) {
        case FAST_POLICY_RANDOM: return launch_fast_one<M, FAST_POLICY_RANDOM, MODE>(args, grid, block, smem, stream);
        case FAST_POLICY_CONSTANT: return launch_fast_one<M, FAST_POLICY_CONSTANT, MODE>(args, grid, block, smem, stream);
        default: return launch_fast_one<M, FAST_POLICY_TABLE, MODE>(args, grid, block, smem, stream);
    }
}

template <>
int32_t launch_uct_fast<GridWorldDev>(const KernelArgs<GridWorldDev>& args, int policy, int mode, int grid, int block, size_t smem,
                                      cudaStream_t stream) {
    switch (mode) {
        case FAST_MODE_DET: return launch_fast_policy<GridWorldDev, FAST_MODE_DET>(args, policy, grid, block, smem, stream);
        case FAST_MODE_AHEAD: return launch_fast_policy<GridWorldDev, FAST_MODE_AHEAD>(args, policy, grid, block, smem, stream);
        default: return launch_fast_policy<GridWorldDev, FAST_MODE_PLAIN>(args, policy, grid, block, smem, stream);
    }
}

}  // namespace mctsb200

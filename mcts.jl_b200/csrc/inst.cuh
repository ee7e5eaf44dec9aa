// inst.cuh -- explicit instantiation of the search kernels for one (MDP, solver kind) pair, all tile widths.
#pragma once
#include "search.cuh"

namespace mctsb200 {

template <class M, int KIND, int G>
static void launch_one(const KernelArgs<M>& args, cudaStream_t stream) {
    const long long threads = args.n_trees * G;
    const unsigned grid = (unsigned)((threads + kBlockThreads - 1) / kBlockThreads);
    auto kernel = search_kernel<M, KIND, G>;
    MCTSB200_LAUNCH(kernel, grid, kBlockThreads, stream, args);
}

template <class M, int KIND, int G>
static int occupancy_one() {
    int blocks = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, search_kernel<M, KIND, G>, kBlockThreads, 0);
    return blocks;
}

#define MCTSB200_INSTANTIATE(M, KIND)                                                               \
    template <>                                                                                     \
    void launch_search<M, KIND>(const KernelArgs<M>& args, int G, cudaStream_t stream) {            \
        switch (G) {                                                                                \
            case 1: launch_one<M, KIND, 1>(args, stream); break;                                    \
            case 2: launch_one<M, KIND, 2>(args, stream); break;                                    \
            case 4: launch_one<M, KIND, 4>(args, stream); break;                                    \
            case 8: launch_one<M, KIND, 8>(args, stream); break;                                    \
            case 16: launch_one<M, KIND, 16>(args, stream); break;                                  \
            default: launch_one<M, KIND, 32>(args, stream); break;                                  \
        }                                                                                           \
    }                                                                                               \
    template <>                                                                                     \
    int search_occupancy<M, KIND>(int G) {                                                          \
        switch (G) {                                                                                \
            case 1: return occupancy_one<M, KIND, 1>();                                             \
            case 2: return occupancy_one<M, KIND, 2>();                                             \
            case 4: return occupancy_one<M, KIND, 4>();                                             \
            case 8: return occupancy_one<M, KIND, 8>();                                             \
            case 16: return occupancy_one<M, KIND, 16>();                                           \
            default: return occupancy_one<M, KIND, 32>();                                           \
        }                                                                                           \
    }

}  // namespace mctsb200

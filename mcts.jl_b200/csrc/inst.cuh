// inst.cuh -- explicit instantiation of the search kernels for one (MDP, solver kind) pair, all tile widths.
#pragma once
#include "search.cuh"

namespace mctsb200 {

template <class M, int KIND, int G>
static void launch_one(const KernelArgs<M>& args, cudaStream_t stream) {
    const long long threads = args.n_slots * G;
    const unsigned grid = (unsigned)((threads + kBlockThreads - 1) / kBlockThreads);
    auto kernel = search_kernel<M, KIND, G>;
    const size_t smem = args.path_smem ? path_smem_bytes(args.geo.depth, G, args.path_tiles) : 0;
    MCTSB200_LAUNCH(kernel, grid, kBlockThreads, smem, stream, args);
}

template <class M, int KIND, int G>
static int occupancy_one(size_t smem) {
    int blocks = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, search_kernel<M, KIND, G>, kBlockThreads, smem);
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
    int search_occupancy<M, KIND>(int G, size_t smem) {                                                          \
        switch (G) {                                                                                \
            case 1: return occupancy_one<M, KIND, 1>(smem);                                             \
            case 2: return occupancy_one<M, KIND, 2>(smem);                                             \
            case 4: return occupancy_one<M, KIND, 4>(smem);                                             \
            case 8: return occupancy_one<M, KIND, 8>(smem);                                             \
            case 16: return occupancy_one<M, KIND, 16>(smem);                                           \
            default: return occupancy_one<M, KIND, 32>(smem);                                           \
        }                                                                                           \
    }

}  // namespace mctsb200

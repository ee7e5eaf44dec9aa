// Search kernels for GridWorldDev / MCTSB200_KIND_DPW (see search.cuh).
#include "inst.cuh"
namespace mctsb200 {
MCTSB200_INSTANTIATE(GridWorldDev, MCTSB200_KIND_DPW)
}

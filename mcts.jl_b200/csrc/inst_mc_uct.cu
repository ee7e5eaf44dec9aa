// Search kernels for MountainCarDev / MCTSB200_KIND_UCT (see search.cuh).
#include "inst.cuh"
namespace mctsb200 {
MCTSB200_INSTANTIATE(MountainCarDev, MCTSB200_KIND_UCT)
}

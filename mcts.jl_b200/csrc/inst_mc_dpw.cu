// Search kernels for MountainCarDev / MCTSB200_KIND_DPW (see search.cuh).
#include "inst.cuh"
namespace mctsb200 {
MCTSB200_INSTANTIATE(MountainCarDev, MCTSB200_KIND_DPW)
}

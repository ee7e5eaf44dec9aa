// GaussianBox (continuous states and actions): only DPW plans for it (src/vanilla.jl:297 enumerates actions(mdp, s)).
#include "inst.cuh"
namespace mctsb200 {
MCTSB200_INSTANTIATE(GaussianBoxDev, MCTSB200_KIND_DPW)
}

// plugin_rt.h -- run-time registration of MDP plugins: NVRTC compilation of the tile kernel around user source and the
// handles of the compiled kernels (implementation: plugin_rt.cu; trait adapter: custom.cuh).
#pragma once
#include <string>

namespace mctsb200_rt {

struct PluginInfo {
    std::string name, source, struct_name;
    int state_doubles = 0, n_actions = 0, max_branch = 0;
    int action_doubles = 0;  // > 0: a continuous action space (CustomBoxDev<state_doubles, action_doubles>; only DPW plans for it)
};

// kernels of one (plugin, solver kind) program
enum KernelId { K_SEARCH_G1 = 0, K_SEARCH_G2, K_SEARCH_G4, K_SEARCH_G8, K_SEARCH_G16, K_SEARCH_G32, K_EPISODE_INIT, K_EPISODE_STEP, K_COUNT };

// Registers a plugin and compiles its vanilla-UCT program (continuous actions: its DPW program) for sm_100a (this needs NVRTC but no GPU, so interface errors
// surface here with the compiler's log in `err`). Returns the plugin index (>= 0) or -1. A name registered before is
// replaced if the definition differs (its index stays).
int register_plugin(const PluginInfo& info, std::string& err);
int lookup_plugin(const char* name);            // index or -1
const PluginInfo* plugin_info(int index);       // null if unknown
// The kernel handle (a cudaKernel_t, usable with cudaLaunchKernel / cudaFuncGetAttributes / the occupancy API) of one
// kernel of the (plugin, kind) program on the current device; compiles and loads on first use. Null + err on failure.
const void* plugin_kernel(int index, int kind, int kernel_id, std::string& err);
// The generated program text of (plugin, kind): what NVRTC is given besides the embedded library headers.
std::string plugin_program_text(const PluginInfo& info, int kind);
// Programs of this process served from the on-disk cache ($MCTSB200_CACHE_DIR) / compiled with NVRTC.
void plugin_cache_stats(long long* hits, long long* misses);

}  // namespace mctsb200_rt

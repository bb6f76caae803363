// fg_abi_f32.cu -- fp32 entry points + the precision-independent ones (see fg_abi_impl.cuh).
#include "fg_abi_impl.cuh"

namespace fgabi { thread_local char g_err[512] = ""; }

extern "C" {

int fg_policy_bfs(const void* pos, const void* ideal_shape, const void* ideal_vel, void* act, int E, int N,
                  int num_agents_per_layer, void* stream) {
    return policy_bfs_impl<float>(pos, ideal_shape, ideal_vel, act, E, N, num_agents_per_layer, stream);
}

int fg_abi_version(void) { return FG_ABI_VERSION; }

const char* fg_last_error(void) { return g_err; }

int fg_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    cudaDeviceProp prop;
    if (err == cudaSuccess) err = cudaGetDeviceProperties(&prop, dev);
    if (err != cudaSuccess) return fail(FG_ERR_CUDA, "fg_device_info: %s", cudaGetErrorString(err));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return FG_OK;
}

int fg_launch_geometry(int N, int* envs_per_cta, int* threads_per_cta) {
    if (N < 1 || N > FG_MAX_AGENTS) return fail(FG_ERR_ARG, "N must be in [1, FG_MAX_AGENTS=256]%s");
    if (envs_per_cta) *envs_per_cta = fg::kBlock / N;
    if (threads_per_cta) *threads_per_cta = fg::kBlock;
    return FG_OK;
}

int fg_world_step(const fg_params* p, const fg_buffers* b, int E, int N, uint64_t seed, uint32_t tick,
                  uint32_t env_offset, void* stream) {
    return world_step_impl<float>(p, b, E, N, seed, tick, env_offset, stream);
}

int fg_obs_reward(const fg_params* p, const fg_buffers* b, int scenario, int E, int N, int L, void* stream) {
    return obs_reward_impl<float>(p, b, scenario, E, N, L, stream);
}

int fg_step_fused(const fg_params* p, const fg_buffers* b, int scenario, int E, int N, int L, int n_steps,
                  int random_actions, int auto_reset, uint64_t seed, uint32_t tick, uint32_t env_offset,
                  void* stream) {
    return step_fused_impl<float>(p, b, scenario, E, N, L, n_steps, random_actions, auto_reset, seed, tick,
                                  env_offset, stream);
}

int fg_reset(const fg_params* p, const fg_buffers* b, int scenario, int E, int N, int L, const uint8_t* mask,
             uint64_t seed, uint32_t tick, uint32_t env_offset, void* stream) {
    return reset_impl<float>(p, b, scenario, E, N, L, mask, seed, tick, env_offset, stream);
}

int fg_random_actions(void* act, int E, int N, uint64_t seed, uint32_t tick, uint32_t env_offset,
                      const uint32_t* tick_dev, void* stream) {
    return random_actions_impl<float>(act, E, N, seed, tick, env_offset, tick_dev, stream);
}

}  // extern "C"

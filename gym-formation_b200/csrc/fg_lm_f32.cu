// fg_lm_f32.cu -- fp32 instantiations of the partial-observation warp kernel (fg_lm.cuh, fg_warp_lm.cuh).
#include "fg_lm.cuh"

namespace fgabi {
int launch_lm_warp(const fg::KArgs<float>& a, int scenario, cudaStream_t st) { return launch_lm_dispatch<float>(a, scenario, st); }
}  // namespace fgabi

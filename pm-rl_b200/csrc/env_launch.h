// env_launch.h — internal: launchers shared between the env translation units.
#pragma once
#include <cuda_runtime.h>
#include "pmrl_device.cuh"

// Warp-specialised TMA pipeline variant of the fused step+obs kernel (env_step_tma.cu).
// Returns -100 if the shape is not supported by this variant (caller falls back).
int pmrl_launch_step_obs_tma(pmrl::StepParams& p, int npl, int stages, int group, cudaStream_t s);

// Register-staged fused step+obs kernel (env_step_fast.cu).  Returns -100 if the shape is not covered.
int pmrl_launch_step_obs_fast(pmrl::StepParams& p, int npl, int group, int ctas_per_sm, cudaStream_t s);
void pmrl_set_fast_variant(int v);

// Variant of the fused kernel that brings each env's weight ring in with one TMA bulk load (env_step_rt.cu).
int pmrl_launch_step_obs_rt(pmrl::StepParams& p, int npl, int group, int ctas_per_sm, cudaStream_t s);

// Variant with tensor-map TMA loads of the feature windows (env_step_tm.cu).
int pmrl_launch_step_obs_tm(pmrl::StepParams& p, int npl, int group, int ctas_per_sm, cudaStream_t s);

// host_util.h — host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

// largest obs tile staged in shared memory by k_obs_build (fits the default 48 KB dynamic limit)
static const size_t kObsTileCapBytes = 40 * 1024;

int pmrl_fail(int code, const char* msg);          // records msg for pmrl_last_error(), returns code
int pmrl_check_launch(const char* what);           // cudaGetLastError() → 0 or positive cudaError_t
int pmrl_sm_count(void);                           // SMs of the current device (cached per device)

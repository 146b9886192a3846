// hz_common.cuh — host-side plumbing shared by the .cu files: launch accounting and CUDA
// error capture for the C ABI (no exceptions cross the boundary).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/harmonies_b200.h"

// defined in hz_abi.cu
extern "C" int hz_record_launch(int n_kernels, cudaError_t err);

// call right after a kernel launch: counts it and converts a launch error into a status
static inline int hz_launched(int n_kernels) { return hz_record_launch(n_kernels, cudaGetLastError()); }

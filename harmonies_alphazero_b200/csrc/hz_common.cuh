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

// ---- bounds-checked debug build (-DHZ_DEBUG_BOUNDS, profiles/run_bounds.sh) ------------------------
// compute-sanitizer is closed on this pool, so the arenas guard themselves in a debug build: every index
// into a tree arena (nodes, edges, hash table, paths), the shared-memory water queue and the network
// buffers is checked; a violation records its site id in hz_bounds_fault and traps the launch (the host
// then sees cudaErrorLaunchFailure and every test of the suite fails loudly).  Release builds compile the
// checks away.
#ifdef HZ_DEBUG_BOUNDS
static __device__ unsigned int hz_bounds_fault_site;   // one per translation unit (the trap is what the host sees)
#define HZ_BOUND(idx, limit, site)                                                       \
    do {                                                                                 \
        if ((unsigned long long)(idx) >= (unsigned long long)(limit)) {                  \
            atomicExch(&hz_bounds_fault_site, (unsigned int)(site));                     \
            __threadfence_system();                                                      \
            __trap();                                                                    \
        }                                                                                \
    } while (0)
#else
#define HZ_BOUND(idx, limit, site) ((void)0)
#endif

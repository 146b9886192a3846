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

// ---- programmatic dependent launch (sm_90+) -------------------------------------------------------------
// The kernels of a simulation step (select -> tower -> FC heads -> expand/backup) are launched with the
// programmatic-stream-serialization attribute: the next kernel's blocks are scheduled as the previous kernel's
// blocks retire and run their prologue (tables, barrier set-up, weight prefetch) before hz_grid_dep_wait(), which
// returns when the previous grid has completed and its writes are visible.  Every kernel launched through
// hz_launch() calls hz_grid_dep_wait() before it touches anything a predecessor wrote.  HZ_NO_PDL=1 in the
// environment launches them plainly (A/B switch).
#include <stdlib.h>

#include <utility>
#ifdef __CUDACC__
__device__ __forceinline__ void hz_grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
static inline bool hz_pdl_enabled() {
    static const bool on = getenv("HZ_NO_PDL") == nullptr;
    return on;
}
template <typename... Exp, typename... Act>
static inline cudaError_t hz_launch(void (*kernel)(Exp...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Act&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = hz_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Act>(args)...);
}

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

// hz_abi.cu — ABI bookkeeping: version, status strings, launch counter, last CUDA error.
#include <atomic>
#include <cstring>

#include "hz_common.cuh"

static std::atomic<uint64_t> g_launches{0};
static char g_last_error[256] = "";

extern "C" {

int hz_record_launch(int n_kernels, cudaError_t err) {
    if (err != cudaSuccess) {
        std::strncpy(g_last_error, cudaGetErrorString(err), sizeof(g_last_error) - 1);
        return HZ_ERR_CUDA;
    }
    g_launches.fetch_add((uint64_t)n_kernels, std::memory_order_relaxed);
    return HZ_OK;
}

int hz_abi_version(void) { return HZ_ABI_VERSION; }

const char* hz_status_string(int status) {
    switch (status) {
        case HZ_OK: return "ok";
        case HZ_ERR_ARG: return "invalid argument";
        case HZ_ERR_CUDA: return "CUDA error";
        case HZ_ERR_NO_DEVICE: return "no CUDA device";
        case HZ_ERR_WORKSPACE: return "workspace too small or misaligned";
        default: return "unknown status";
    }
}

const char* hz_last_cuda_error(void) { return g_last_error; }

uint64_t hz_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

}  // extern "C"

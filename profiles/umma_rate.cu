// Tensor-pipe rate probe (B200, via gpurun): how many SM cycles does one tcgen05.mma (M=128, K=16, bf16)
// take for the N values the tower uses, with A re-read from shared memory every time versus held in the
// A collector, and B K-major SWIZZLE_128B versus MN-major without swizzle?  One CTA per SM on all SMs
// (so that the power/clock state is the loaded one), 512 MMAs back to back per measurement.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../harmonies_alphazero_b200/csrc/hz_sm100.cuh"
using namespace hz::sm100;

__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}

template <int MODE>   // 0: plain, 1: collector fill/use/lastuse in groups of 3, 2: groups of 3 with the same A but no hints
__global__ void __launch_bounds__(128, 1) rate(int N, int bmn, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_init_fence(); }
    for (int i = threadIdx.x; i < (160 * 1024) / 16; i += 128) ((uint4*)sm)[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0, 0);
    fence_proxy_async_smem();
    __syncthreads();
    if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tb = slot;
    if (threadIdx.x == 0) {
        uint32_t idesc = idesc_bf16_f32(128, N) | (bmn ? (1u << 16) : 0u);
        uint32_t sA = smem_u32(sm), sB = smem_u32(sm + 64 * 1024);
        long long t0 = clock64();
        for (int it = 0; it < 512; it++) {
            int grp = it / 3, r = it % 3;
            uint32_t aoff = (MODE == 0 ? (it & 15) : (grp & 15)) * 4096 + (it & 3) * 0;   // a different 128x16 weight slice per MMA / per group
            uint64_t da = smem_desc_sw128(sA + (aoff & 0xFFFF));
            uint32_t boff = (uint32_t)((it % 5) * 7 * 16) * (bmn ? 16 : 128);
            uint64_t db = bmn ? desc_mn(sB + boff, 8960, 128) : smem_desc_sw128(sB + boff);
            uint32_t d = tb + (it & 3) * 128;
            if (MODE == 1) {
                if (r == 0) umma_bf16_coll<COLL_FILL>(d, da, db, idesc, 1);
                else if (r == 1) umma_bf16_coll<COLL_USE>(d, da, db, idesc, 1);
                else umma_bf16_coll<COLL_LASTUSE>(d, da, db, idesc, 1);
            } else umma_bf16(d, da, db, idesc, 1);
        }
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0, nullptr, 0);
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
    long long* d;
    cudaMalloc(&d, 8);
    int smem = 162 * 1024;
    cudaFuncSetAttribute(rate<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(rate<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(rate<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int Ns[] = {64, 96, 112, 128, 192, 224, 256};
    for (int bmn = 0; bmn < 2; bmn++)
        for (int N : Ns)
            for (int mode = 0; mode < 3; mode++) {
                long long best = 1LL << 60;
                for (int rep = 0; rep < 3; rep++) {
                    if (mode == 0) rate<0><<<148, 128, smem>>>(N, bmn, d);
                    else if (mode == 1) rate<1><<<148, 128, smem>>>(N, bmn, d);
                    else rate<2><<<148, 128, smem>>>(N, bmn, d);
                    long long h;
                    if (cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("CUDA error\n"); return 1; }
                    if (h < best) best = h;
                }
                printf("B %s N=%3d %-28s: %6.1f cycles/MMA (tensor floor N/2 = %d)\n", bmn ? "MN-major/none " : "K-major/SW128 ", N,
                       mode == 0 ? "A re-read per MMA" : mode == 1 ? "A collector (groups of 3)" : "same A x3, no hints", best / 512.0, N / 2);
            }
    return 0;
}

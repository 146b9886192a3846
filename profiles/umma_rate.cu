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

template <int MODE>   // 0: a different A per MMA, 1: collector fill/use/lastuse in groups of 3, 2: groups of 3 with the same A, no hints
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
    if (warp == 0) {
        const uint32_t idesc = idesc_bf16_f32(128, N) | (bmn ? (1u << 16) : 0u);
        const uint32_t sA = smem_u32(sm), sB = smem_u32(sm + 64 * 1024);
        const uint32_t a_lo = (1u << 16) | (sA >> 4);
        const uint32_t a_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t b_lo = bmn ? (((8960u >> 4) << 16) | (sB >> 4)) : ((1u << 16) | (sB >> 4));
        const uint32_t b_hi = bmn ? ((128u >> 4) | (1u << 14)) : a_hi;
        const uint32_t ncol = N > 128 ? 256u : 128u;              // accumulator slots of this width
        long long t0 = clock64();
        uint32_t pred;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
        if (pred) {
#pragma unroll 1
            for (int it = 0; it < 32; it++) {
#pragma unroll
                for (int j = 0; j < 12; j++) {       // 4 "k-steps" x 3 "rows", like one weight stage of the tower
                    const int kk = j / 3, r = j % 3;
                    const uint64_t da = ((uint64_t)a_hi << 32) | (a_lo + (uint32_t)((MODE == 0 ? j : kk) * 256));   // 4 KB apart
                    const uint64_t db = ((uint64_t)b_hi << 32) | (b_lo + (uint32_t)(r * (bmn ? 112 : 896)));
                    const uint32_t d = tb + (uint32_t)(r * ncol) % 512u;
                    if (MODE == 1) {
                        if (r == 0) umma_bf16_coll<COLL_FILL>(d, da, db, idesc, 1);
                        else if (r == 1) umma_bf16_coll<COLL_USE>(d, da, db, idesc, 1);
                        else umma_bf16_coll<COLL_LASTUSE>(d, da, db, idesc, 1);
                    } else umma_bf16(d, da, db, idesc, 1);
                }
            }
            umma_commit(smem_u32(&bar));
        }
        __syncwarp();
        mbar_wait(smem_u32(&bar), 0, nullptr, 0);
        long long t1 = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
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
    int Ns[] = {32, 64, 80, 96, 112, 128, 160, 192, 208, 224, 240, 256};
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
                       mode == 0 ? "A re-read per MMA" : mode == 1 ? "A collector (groups of 3)" : "same A x3, no hints", best / 384.0, N / 2);
            }
    return 0;
}

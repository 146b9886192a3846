// Primitive probe for the sm_100a building blocks of csrc/hz_tower.cu (run on a B200 through
// gpurun; not part of the library).  One variant per process so that a trapped variant cannot
// poison the others:   umma_probe <N> <b_row_off> <d_col_off> <use_bulk> <n_kblocks>
// Checks D[128 x N] = A[128 x K] * B[N x K]^T (bf16 in, fp32 accumulate in TMEM) where both operands
// sit in shared memory as K-major SWIZZLE_128B tiles, B is a window starting `b_row_off` rows into
// a taller tile (the shifted-window trick of the 3x3 taps) and D starts `d_col_off` columns into
// the TMEM allocation.  use_bulk = 1 loads the operands with cp.async.bulk + mbarrier.
// bmode (7th argument): 0 = B K-major SWIZZLE_128B; 1 = B MN-major, no swizzle (core matrix = 8 k
// x 8 positions, positions contiguous; LBO = K-direction stride, SBO = MN-direction stride);
// 2 = the same image with LBO/SBO swapped in the descriptor (must FAIL if 1 is right).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include "../harmonies_alphazero_b200/csrc/hz_sm100.cuh"

using namespace hz::sm100;

constexpr int MAXROWS_B = 288;
constexpr int A_BYTES = 128 * 128;          // one k-block: 128 rows x 64 bf16
constexpr int B_BYTES = MAXROWS_B * 128;

constexpr int LBO_B = (MAXROWS_B / 8) * 128;   // MN-major image: stride between 8-channel groups

__device__ __forceinline__ uint64_t smem_desc_mn_none(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

__global__ void __launch_bounds__(128) probe(const uint8_t* Aimg, const uint8_t* Bimg, float* D, int N, int brow_off, int dcol_off,
                                             int use_bulk, int nkb, unsigned int* fault, int bmode) {
    extern __shared__ uint8_t raw[];
    uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = sm;                         // [nkb][16 KB]
    uint8_t* sB = sm + 2 * A_BYTES;           // [nkb][B_BYTES]
    uint64_t* bars = (uint64_t*)(sm + 2 * A_BYTES + 2 * B_BYTES);
    uint32_t* tmem_slot = (uint32_t*)(bars + 4);
    int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(smem_u32(&bars[0]), 1);
        mbar_init(smem_u32(&bars[1]), 1);
        mbar_init_fence();
    }
    __syncthreads();
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tbase = *tmem_slot;
    if (use_bulk) {
        if (tid == 0) {
            mbar_expect_tx(smem_u32(&bars[0]), (uint32_t)nkb * (A_BYTES + B_BYTES));
            for (int kb = 0; kb < nkb; kb++) {
                bulk_g2s(smem_u32(sA + kb * A_BYTES), Aimg + (size_t)kb * A_BYTES, A_BYTES, smem_u32(&bars[0]));
                bulk_g2s(smem_u32(sB + kb * B_BYTES), Bimg + (size_t)kb * B_BYTES, B_BYTES, smem_u32(&bars[0]));
            }
        }
        mbar_wait(smem_u32(&bars[0]), 0, fault, 1);
    } else {
        for (int i = tid; i < nkb * A_BYTES / 16; i += 128) ((uint4*)sA)[i] = ((const uint4*)Aimg)[i];
        for (int kb = 0; kb < nkb; kb++)
            for (int i = tid; i < B_BYTES / 16; i += 128) ((uint4*)(sB + kb * B_BYTES))[i] = ((const uint4*)(Bimg + (size_t)kb * B_BYTES))[i];
        fence_proxy_async_smem();
        __syncthreads();
    }
    if (tid == 0) {
        tc_fence_after();
        uint32_t idesc = idesc_bf16_f32(128, N) | (bmode ? (1u << 16) : 0u);
        for (int kb = 0; kb < nkb; kb++)
            for (int k = 0; k < 4; k++) {
                uint64_t da = smem_desc_sw128(smem_u32(sA + kb * A_BYTES) + k * 32);
                uint64_t db;
                if (bmode == 0) db = smem_desc_sw128(smem_u32(sB + kb * B_BYTES) + brow_off * 128 + k * 32);
                else {
                    uint32_t a = smem_u32(sB + kb * B_BYTES) + (brow_off >> 3) * 128 + k * 2 * LBO_B;
                    db = bmode == 1 ? smem_desc_mn_none(a, LBO_B, 128) : smem_desc_mn_none(a, 128, LBO_B);
                }
                umma_bf16(tbase + dcol_off, da, db, idesc, (kb | k) ? 1u : 0u);
            }
        umma_commit(smem_u32(&bars[1]));
    }
    mbar_wait(smem_u32(&bars[1]), 0, fault, 2);
    tc_fence_after();
    for (int c = 0; c < 512; c += 16) {
        uint32_t v[16];
        tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + c, v);
        tmem_ld_wait();
        for (int j = 0; j < 16; j++) D[(size_t)(warp * 32 + lane) * 512 + c + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

static uint16_t f2bf(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static float bf2f(uint16_t h) {
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
static size_t sw128(int row, int k) { return (size_t)row * 128 + ((((k >> 3) ^ (row & 7)) << 4) | ((k & 7) * 2)); }

int main(int argc, char** argv) {
    int N = argc > 1 ? atoi(argv[1]) : 112, brow = argc > 2 ? atoi(argv[2]) : 0, dcol = argc > 3 ? atoi(argv[3]) : 0;
    int bulk = argc > 4 ? atoi(argv[4]) : 0, nkb = argc > 5 ? atoi(argv[5]) : 1, bmode = argc > 6 ? atoi(argv[6]) : 0;
    if (brow + N > MAXROWS_B || dcol + N > 512 || nkb < 1 || nkb > 2) { printf("bad args\n"); return 2; }
    std::vector<uint16_t> A((size_t)nkb * 128 * 64), B((size_t)nkb * MAXROWS_B * 64);
    std::vector<uint8_t> Aimg((size_t)nkb * A_BYTES), Bimg((size_t)nkb * B_BYTES);
    srand(1234 + N);
    for (auto& x : A) x = f2bf((float)(rand() % 17 - 8) / 8.0f);
    for (auto& x : B) x = f2bf((float)(rand() % 13 - 6) / 4.0f);
    for (int kb = 0; kb < nkb; kb++) {
        for (int r = 0; r < 128; r++)
            for (int k = 0; k < 64; k++) memcpy(&Aimg[(size_t)kb * A_BYTES + sw128(r, k)], &A[((size_t)kb * 128 + r) * 64 + k], 2);
        for (int r = 0; r < MAXROWS_B; r++)
            for (int k = 0; k < 64; k++) {
                size_t off = bmode == 0 ? sw128(r, k) : (size_t)(k >> 3) * LBO_B + (size_t)(r >> 3) * 128 + (k & 7) * 16 + (r & 7) * 2;
                memcpy(&Bimg[(size_t)kb * B_BYTES + off], &B[((size_t)kb * MAXROWS_B + r) * 64 + k], 2);
            }
    }
    uint8_t *dA, *dB;
    float* dD;
    unsigned int* fault;
    cudaMalloc(&dA, Aimg.size());
    cudaMalloc(&dB, Bimg.size());
    cudaMalloc(&dD, 128 * 512 * 4);
    cudaHostAlloc(&fault, 4, cudaHostAllocMapped);
    *fault = 0;
    cudaMemcpy(dA, Aimg.data(), Aimg.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, Bimg.data(), Bimg.size(), cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, 128 * 512 * 4);
    int smem = 2 * A_BYTES + 2 * B_BYTES + 1024 + 256;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<<<1, 128, smem>>>(dA, dB, dD, N, brow, dcol, bulk, nkb, fault, bmode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("N=%d brow=%d dcol=%d bulk=%d nkb=%d bmode=%d: CUDA ERROR %s (fault code %u)\n", N, brow, dcol, bulk, nkb, bmode, cudaGetErrorString(e), *fault);
        return 1;
    }
    std::vector<float> D(128 * 512);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    int bad = 0;
    for (int m = 0; m < 128; m++)
        for (int n = 0; n < N; n++) {
            double ref = 0;
            for (int kb = 0; kb < nkb; kb++)
                for (int k = 0; k < 64; k++)
                    ref += (double)bf2f(A[((size_t)kb * 128 + m) * 64 + k]) * (double)bf2f(B[((size_t)kb * MAXROWS_B + brow + n) * 64 + k]);
            double err = fabs(ref - (double)D[(size_t)m * 512 + dcol + n]);
            if (err > maxerr) maxerr = err;
            if (err > 1e-3) {
                if (bad < 4) printf("  mismatch m=%d n=%d ref=%f got=%f\n", m, n, ref, D[(size_t)m * 512 + dcol + n]);
                bad++;
            }
        }
    printf("N=%d brow=%d dcol=%d bulk=%d nkb=%d bmode=%d: maxerr=%g bad=%d -> %s\n", N, brow, dcol, bulk, nkb, bmode, maxerr, bad, bad ? "FAIL" : "OK");
    return bad ? 1 : 0;
}

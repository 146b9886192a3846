// hz_tower.cu — hand-written sm_100a 3x3 convolution for the residual tower of the reference
// network (model.py:325-357 stem conv+bn+relu, ResidualBlock.forward model.py:380-392), BatchNorm
// folded, bf16 operands, fp32 accumulation in tensor memory.  SURVEY.md §8 row f4.
//
// Formulation (per layer):  out[o][p] = sum over taps t=(dy,dx), channels i of W_t[o][i] * in[i][p + s(t)]
//   * tcgen05.mma, M = 128 output channels (A = one tap's weight tile, K-major SWIZZLE_128B),
//     N = board positions (B = a window of the activation tile, K-major SWIZZLE_128B), accumulators
//     in TMEM: lane = output channel, column = position.
//   * activations live in HBM as 16-board tiles, cell-major: row (cell*16 + board) of 128 bytes per
//     64-channel half ("T16" layout, already in the shared-memory swizzle so that one bulk copy
//     (TMA, cp.async.bulk) brings a tile in).  With cells outermost a tap is a shift by whole
//     cells: the B operand of tap (dy,dx) for output board-row r is the SAME resident tile read
//     through a descriptor whose start address moved by ((r+dy)*7 + max(dx,0)) cells.
//   * taps that fall off the 5x7 board are never multiplied: rows with r+dy outside 0..4 skip the
//     tap, and a dx = -1 / +1 tap covers only the six cells x = 1..6 / 0..5 (N = 96 instead of 112,
//     accumulator window moved by 16 columns).  Executed MACs = 247/315 of a zero-padded conv.
//   * a whole tile stays in shared memory for all nine taps (one HBM/L2 read per layer); the
//     weights stream through a 5-stage ring (16 KB = one tap x one channel half per stage).
//   * TMEM holds 512 positions (4 board rows of 16 boards) but a tile has 560, so a tile is two
//     passes over the weight stream: board rows {0,1,2} then {3,4}.  Row accumulators live in four
//     128-column units: rows 0..3 own units 0..3 and row 4 shares unit 0 with row 0.  The tap order
//     makes that free of stalls: pass 0 runs dy = +1, 0, -1, so row 0 (which has no dy = -1 tap) is
//     complete and drained before pass 1 starts; pass 1 runs dy = -1, 0, +1, so row 4 (no dy = +1)
//     is drained before the next tile's row 0 needs the unit.  The epilogue of one row therefore
//     always overlaps the MMAs of the others.
//   * warp roles: 0 = weight producer, 1 = MMA issuer, 2 = activation producer, 3 = TMEM
//     allocator, 4..7 = epilogue (bias + residual + ReLU + bf16, thread = output channel).
#include "hz_common.cuh"
#include "hz_sm100.cuh"

namespace hz {
namespace tower {
using namespace hz::sm100;

constexpr int G = 16;                          // boards per tile
constexpr int CELLS = 35, BROWS = 5, BCOLS = 7;
constexpr int TILE_ROWS = CELLS * G;           // 560 rows of 128 bytes per channel half
constexpr int KH_BYTES = TILE_ROWS * 128;      // 71,680
constexpr int ROW_BYTES = BCOLS * G * 128;     // 14,336: one board row of one channel half
constexpr int W_BYTES = 128 * 128;             // 16,384: [128 out][64 in] bf16
constexpr int NSTAGE = 5;
constexpr int NUNIT = 4, UNIT_COLS = 128;
constexpr int NTHREADS = 256;
constexpr int OFF_X = 0;
constexpr int OFF_W = 2 * KH_BYTES;
constexpr int OFF_BAR = OFF_W + NSTAGE * W_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;   // + alignment slack

// barrier indices
constexpr int B_WFULL = 0, B_WEMPTY = NSTAGE, B_AFULL = 2 * NSTAGE, B_AEMPTY = 2 * NSTAGE + 2, B_TFULL = 2 * NSTAGE + 4,
              B_TEMPTY = 2 * NSTAGE + 4 + NUNIT, B_DONE = 2 * NSTAGE + 4 + 2 * NUNIT, N_BARS = B_DONE + 1;

// tap = ky*3 + kx (dy = ky-1, dx = kx-1).  Within a dy group the dx = 0 tap comes first: the first
// MMA into a row accumulator overwrites it and must cover all 112 columns.
// pass 0 (rows 0,1,2): dy = +1, 0, -1  -> row 0 (no dy = -1) completes after two thirds of the pass
// pass 1 (rows 3,4)  : dy = -1, 0, +1  -> row 4 (no dy = +1) completes after two thirds of the pass
__constant__ int8_t TAP_ORDER[2][9] = {{7, 6, 8, 4, 3, 5, 1, 0, 2}, {1, 0, 2, 4, 3, 5, 7, 6, 8}};
__constant__ int8_t LAST_TAP[5] = {5, 2, 2, 8, 5};   // last tap (in its pass's order) that touches row r
__constant__ int8_t EPI_ORDER[5] = {0, 1, 2, 4, 3};  // order in which the row accumulators complete
// accumulator unit of board row r and how many times the unit has been used before (tile iteration it)
__device__ __forceinline__ int unit_of(int r) { return r == 4 ? 0 : r; }
__device__ __forceinline__ int use_of(int r, int it) { return r == 0 ? 2 * it : r == 4 ? 2 * it + 1 : it; }

struct Params {
    const uint8_t* x;      // input tiles  [n_tiles][nkh][560][128 B]
    const uint8_t* w;      // weight tiles [9][nkh][128][128 B]
    const float* bias;     // [128]
    const uint8_t* res;    // residual tiles [n_tiles][2][560][128 B] or null
    uint8_t* y;            // output: T16 tiles (2 halves) or NHWC [n_boards][35][128]
    int n_tiles, nkh, relu, out_nhwc;
    unsigned int* fault;
};

__device__ __forceinline__ uint32_t t16_offset(int rr, int c) {   // byte offset of (row rr, channel c) inside a tile (2 halves)
    return (uint32_t)(c >> 6) * KH_BYTES + (uint32_t)rr * 128u + ((uint32_t)(((c & 63) >> 3) ^ (rr & 7)) << 4) + (uint32_t)(c & 7) * 2u;
}

__global__ void __launch_bounds__(NTHREADS, 1) k_conv3x3(Params P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t sX = smem_u32(sm + OFF_X), sW = smem_u32(sm + OFF_W), sBar = smem_u32(sm + OFF_BAR);
    uint32_t* tmem_slot = (uint32_t*)(sm + OFF_BAR + N_BARS * 8);
    auto bar = [&](int i) { return sBar + 8u * (uint32_t)i; };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkh = P.nkh;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; i++) { mbar_init(bar(B_WFULL + i), 1); mbar_init(bar(B_WEMPTY + i), 1); }
        for (int i = 0; i < 2; i++) { mbar_init(bar(B_AFULL + i), 1); mbar_init(bar(B_AEMPTY + i), 1); }
        for (int i = 0; i < NUNIT; i++) { mbar_init(bar(B_TFULL + i), 1); mbar_init(bar(B_TEMPTY + i), 4); }
        mbar_init(bar(B_DONE), 1);
        mbar_init_fence();
    }
    if (warp == 3) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tmem_slot;

    if (warp == 0 && lane == 0) {
        // ---- weight producer: the tap stream of every pass of every tile, through the ring ----
        uint32_t stage = 0, ph = 0;
        for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x)
            for (int pass = 0; pass < 2; pass++)
                for (int kh = 0; kh < nkh; kh++)
                    for (int ti = 0; ti < 9; ti++) {
                        int tap = TAP_ORDER[pass][ti];
                        mbar_wait(bar(B_WEMPTY + stage), ph ^ 1, P.fault, 0x100 + stage);
                        mbar_expect_tx(bar(B_WFULL + stage), W_BYTES);
                        bulk_g2s(sW + stage * W_BYTES, P.w + (size_t)(tap * nkh + kh) * W_BYTES, W_BYTES, bar(B_WFULL + stage));
                        if (++stage == NSTAGE) { stage = 0; ph ^= 1; }
                    }
    } else if (warp == 2 && lane == 0) {
        // ---- activation producer: one channel half of a tile per buffer ----
        int it = 0;
        for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, it++)
            for (int kh = 0; kh < nkh; kh++) {
                mbar_wait(bar(B_AEMPTY + kh), (it & 1) ^ 1, P.fault, 0x200 + kh);
                mbar_expect_tx(bar(B_AFULL + kh), KH_BYTES);
                const uint8_t* src = P.x + ((size_t)tile * nkh + kh) * KH_BYTES;
                for (int r = 0; r < BROWS; r++)
                    bulk_g2s(sX + kh * KH_BYTES + r * ROW_BYTES, src + (size_t)r * ROW_BYTES, ROW_BYTES, bar(B_AFULL + kh));
            }
    } else if (warp == 1 && lane == 0) {
        // ---- MMA issuer ----
        const uint32_t idesc112 = idesc_bf16_f32(128, 112), idesc96 = idesc_bf16_f32(128, 96);
        uint32_t stage = 0, ph = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, it++) {
            uint32_t started = 0;                                  // rows whose accumulator holds this tile's sums
            for (int pass = 0; pass < 2; pass++) {
                const int r0 = pass ? 3 : 0, r1 = pass ? 5 : 3;
                for (int kh = 0; kh < nkh; kh++) {
                    if (pass == 0) mbar_wait(bar(B_AFULL + kh), it & 1, P.fault, 0x300 + kh);
                    for (int ti = 0; ti < 9; ti++) {
                        const int tap = TAP_ORDER[pass][ti];
                        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                        mbar_wait(bar(B_WFULL + stage), ph, P.fault, 0x400 + stage);
                        tc_fence_after();
                        const uint32_t wa = sW + stage * W_BYTES;
                        const uint32_t idesc = dx ? idesc96 : idesc112;
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const uint64_t da = smem_desc_sw128(wa + k * 32);
                            for (int r = r0; r < r1; r++) {
                                const int sr = r + dy;
                                if (sr < 0 || sr >= BROWS) continue;
                                const int unit = unit_of(r);
                                if (!((started >> r) & 1u)) {       // first use of the unit for this row: the previous tenant must be drained
                                    mbar_wait(bar(B_TEMPTY + unit), (use_of(r, it) & 1) ^ 1, P.fault, 0x500 + unit);
                                    tc_fence_after();
                                }
                                const uint32_t cell0 = (uint32_t)(sr * BCOLS + (dx > 0 ? 1 : 0));
                                const uint64_t db = smem_desc_sw128(sX + kh * KH_BYTES + cell0 * (G * 128) + k * 32);
                                const uint32_t d = tbase + unit * UNIT_COLS + (dx < 0 ? G : 0);
                                umma_bf16(d, da, db, idesc, (started >> r) & 1u);
                                started |= 1u << r;
                            }
                        }
                        umma_commit(bar(B_WEMPTY + stage));        // frees the weight stage when these MMAs have read it
                        if (kh == nkh - 1)
                            for (int r = r0; r < r1; r++)
                                if (tap == LAST_TAP[r]) umma_commit(bar(B_TFULL + unit_of(r)));
                        if (++stage == NSTAGE) { stage = 0; ph ^= 1; }
                    }
                    if (pass == 1) umma_commit(bar(B_AEMPTY + kh));   // the tile's channel half is no longer read
                }
            }
        }
        umma_commit(bar(B_DONE));
        mbar_wait(bar(B_DONE), 0, P.fault, 0x600);
    } else if (warp >= 4) {
        // ---- epilogue: thread = output channel; 16 boards of one cell per TMEM load ----
        const int q = warp & 3, c = q * 32 + lane;
        const float bias = P.bias[c];
        int it = 0;
        for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, it++) {
            const size_t tile_off = (size_t)tile * 2 * KH_BYTES;
            for (int ri = 0; ri < BROWS; ri++) {
                const int r = EPI_ORDER[ri], unit = unit_of(r);
                mbar_wait(bar(B_TFULL + unit), use_of(r, it) & 1, P.fault, 0x700 + unit);
                tc_fence_after();
                for (int x = 0; x < BCOLS; x++) {
                    uint32_t v[16];
                    tmem_ld16(tbase + ((uint32_t)(q * 32) << 16) + unit * UNIT_COLS + x * G, v);
                    tmem_ld_wait();
                    const int cell = r * BCOLS + x;
#pragma unroll
                    for (int b = 0; b < G; b++) {
                        const int rr = cell * G + b;
                        const uint32_t off = t16_offset(rr, c);
                        float o = __uint_as_float(v[b]) + bias;
                        if (P.res) o += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(P.res + tile_off + off));
                        if (P.relu) o = fmaxf(o, 0.0f);
                        if (P.out_nhwc)
                            reinterpret_cast<__nv_bfloat16*>(P.y)[((size_t)(tile * G + b) * CELLS + cell) * 128 + c] = __float2bfloat16_rn(o);
                        else
                            *reinterpret_cast<__nv_bfloat16*>(P.y + tile_off + off) = __float2bfloat16_rn(o);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(B_TEMPTY + unit));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 3) tmem_dealloc(tbase, 512);
}

// ---- layout conversion (interop with NHWC tensors: tests, the leaf encoder's output) ------------
// src [n][35][C] bf16 (C % 8 == 0, C <= 64*nkh) -> tiles [n_pad/16][nkh][560][128 B]; pad boards/channels = 0
__global__ void k_to_tiles(const uint4* __restrict__ src, uint4* __restrict__ dst, int64_t n, int C, int nkh, int64_t n_chunks) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_chunks; i += (int64_t)gridDim.x * blockDim.x) {
        int j = (int)(i & 7);                       // chunk position inside the 128-byte row (swizzled)
        int64_t row = i >> 3;                       // global row: ((tile*nkh + kh)*560 + rr)
        int rr = (int)(row % TILE_ROWS);
        int64_t tk = row / TILE_ROWS;
        int kh = (int)(tk % nkh);
        int64_t tile = tk / nkh;
        int cell = rr / G, b = rr % G;
        int chunk = j ^ (rr & 7);                   // logical 8-channel group stored at position j
        int ch = kh * 64 + chunk * 8;
        int64_t board = tile * G + b;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (board < n && ch < C) v = src[((board * CELLS + cell) * C + ch) >> 3];
        dst[i] = v;
    }
}
// tiles (2 halves) -> dst [n][35][128] bf16
__global__ void k_from_tiles(const uint4* __restrict__ src, uint4* __restrict__ dst, int64_t n, int64_t n_chunks) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_chunks; i += (int64_t)gridDim.x * blockDim.x) {
        int g = (int)(i & 15);                      // 8-channel group of the NHWC row
        int64_t pos = i >> 4;                       // board*35 + cell
        int cell = (int)(pos % CELLS);
        int64_t board = pos / CELLS;
        int64_t tile = board / G;
        int rr = cell * G + (int)(board % G);
        int kh = g >> 3, chunk = g & 7;
        size_t off = ((size_t)(tile * 2 + kh) * TILE_ROWS + rr) * 128 + (size_t)((chunk ^ (rr & 7)) << 4);
        dst[i] = src[off >> 4];
    }
}

static int g_max_ctas = 0;   // 0 = one CTA per SM; tests lower it to drive several tiles through one CTA

static int ensure_attr() {
    static bool done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_conv3x3, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return hz_record_launch(0, e);
        if (dev >= 0 && dev < 64) done[dev] = true;
    }
    return HZ_OK;
}

}  // namespace tower
}  // namespace hz

extern "C" {

size_t hz_tower_tile_bytes(int64_t n_boards, int channel_halves) {
    if (n_boards < 0 || channel_halves < 1 || channel_halves > 2) return 0;
    int64_t tiles = (n_boards + hz::tower::G - 1) / hz::tower::G;
    return (size_t)tiles * channel_halves * hz::tower::KH_BYTES;
}

int hz_tower_set_max_ctas(int max_ctas) {
    hz::tower::g_max_ctas = max_ctas > 0 ? max_ctas : 0;
    return HZ_OK;
}

int hz_tower_to_tiles(const void* src_nhwc, void* dst_tiles, int64_t n_boards, int channels, int channel_halves, void* stream) {
    if (!src_nhwc || !dst_tiles || n_boards <= 0 || channels <= 0 || (channels & 7) || channel_halves < 1 || channel_halves > 2 ||
        channels > 64 * channel_halves || ((uintptr_t)src_nhwc & 15) || ((uintptr_t)dst_tiles & 15))
        return HZ_ERR_ARG;
    int64_t tiles = (n_boards + hz::tower::G - 1) / hz::tower::G;
    int64_t n_chunks = tiles * channel_halves * hz::tower::TILE_ROWS * 8;
    int grid = (int)((n_chunks + 255) / 256 < 148 * 16 ? (n_chunks + 255) / 256 : 148 * 16);
    hz::tower::k_to_tiles<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)src_nhwc, (uint4*)dst_tiles, n_boards, channels,
                                                                   channel_halves, n_chunks);
    return hz_launched(1);
}

int hz_tower_from_tiles(const void* src_tiles, void* dst_nhwc, int64_t n_boards, void* stream) {
    if (!src_tiles || !dst_nhwc || n_boards <= 0 || ((uintptr_t)src_tiles & 15) || ((uintptr_t)dst_nhwc & 15)) return HZ_ERR_ARG;
    int64_t n_chunks = n_boards * hz::tower::CELLS * 16;
    int grid = (int)((n_chunks + 255) / 256 < 148 * 16 ? (n_chunks + 255) / 256 : 148 * 16);
    hz::tower::k_from_tiles<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)src_tiles, (uint4*)dst_nhwc, n_boards, n_chunks);
    return hz_launched(1);
}

int hz_tower_conv3x3(const void* x_tiles, int in_channel_halves, const void* w_tiles, const float* bias, const void* residual_tiles,
                     void* y, int64_t n_boards, int relu, int out_nhwc, unsigned int* fault, void* stream) {
    using namespace hz::tower;
    if (!x_tiles || !w_tiles || !bias || !y || n_boards <= 0 || (n_boards % G) || in_channel_halves < 1 || in_channel_halves > 2)
        return HZ_ERR_ARG;
    if (((uintptr_t)x_tiles | (uintptr_t)w_tiles | (uintptr_t)y | (uintptr_t)residual_tiles) & 15) return HZ_ERR_ARG;
    int st = ensure_attr();
    if (st != HZ_OK) return st;
    Params P;
    P.x = (const uint8_t*)x_tiles;
    P.w = (const uint8_t*)w_tiles;
    P.bias = bias;
    P.res = (const uint8_t*)residual_tiles;
    P.y = (uint8_t*)y;
    P.n_tiles = (int)(n_boards / G);
    P.nkh = in_channel_halves;
    P.relu = relu;
    P.out_nhwc = out_nhwc;
    P.fault = fault;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_max_ctas > 0 && g_max_ctas < sms) sms = g_max_ctas;
    int grid = P.n_tiles < sms ? P.n_tiles : sms;
    k_conv3x3<<<grid, NTHREADS, SMEM_BYTES, (cudaStream_t)stream>>>(P);
    return hz_launched(1);
}

}  // extern "C"

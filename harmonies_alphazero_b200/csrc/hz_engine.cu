// hz_engine.cu — batched engine kernels (sm_100a) behind the C ABI of
// include/harmonies_b200.h: new games, legal masks, move application, scoring, state
// encoding, canonical keys, outcomes, random-playout policy and the fused playout.
//
// Mapping: one thread per game, state held in registers (7 x LDG.128 in, 8 x STG.128 out),
// hex-parallel work done as 23-bit bit-board arithmetic; the neighbour expansion LUT lives in
// 3 KB of shared memory per block.  hz_encode uses a block-cooperative layout instead
// (write-bound: 5.5 KB out per 128 B in).
#include "hz_common.cuh"
#include "hz_core.cuh"

namespace hz {

constexpr int TPB = 128;

__global__ void __launch_bounds__(TPB) k_init(void* states, int64_t n, const uint64_t* keys,
                                              uint64_t seed, uint64_t first_id) {
    int64_t g = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (g >= n) return;
    State s;
    init_state(s, keys ? keys[g] : rand64(seed, first_id + (uint64_t)g));
    store_state(s, states, g);
}

__global__ void __launch_bounds__(TPB) k_legal(const void* states, int64_t n, uint32_t* mask) {
    int64_t g = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (g >= n) return;
    State s;
    load_state(s, states, g);
    uint32_t out[5];
    legal_words(legal_of(s), out);
#pragma unroll
    for (int k = 0; k < 5; k++) mask[g * 5 + k] = out[k];
}

__global__ void __launch_bounds__(TPB) k_apply(void* states, int64_t n, const int16_t* actions,
                                               const uint16_t* draws, uint8_t* status) {
    __shared__ NbrLut lut;
    build_nbr_lut(&lut);
    __syncthreads();
    int64_t g = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (g >= n) return;
    State s;
    load_state(s, states, g);
    uint32_t ex = draws ? (uint32_t)draws[g] : (uint32_t)HZ_NO_DRAW;
    int st = apply_move(s, (int)actions[g], ex, key_of(s), s.w[HZ_W_EVENT], true, &lut);
    if (st == HZ_MOVE_OK) store_state(s, states, g);
    if (status) status[g] = (uint8_t)st;
}

__global__ void __launch_bounds__(TPB) k_score(const void* states, int64_t n, int16_t* scores,
                                               int16_t* terms) {
    __shared__ NbrLut lut;
    build_nbr_lut(&lut);
    __syncthreads();
    int64_t g = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (g >= n) return;
    // only the 18 board words are needed: 92 -> 72 B of traffic per position
    const uint4* p = reinterpret_cast<const uint4*>(states) + g * 8;
    uint32_t w[20];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        uint4 v = p[k];
        w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
    }
#pragma unroll
    for (int pl = 0; pl < 2; pl++) {
        Board b;
#pragma unroll
        for (int k = 0; k < 9; k++) b.p[k] = w[pl * 9 + k];
        int t[5];
        score_board(&lut, b, t);
        if (terms) {
#pragma unroll
            for (int k = 0; k < 5; k++) terms[(g * 2 + pl) * 5 + k] = (int16_t)t[k];
        }
        if (scores) scores[g * 2 + pl] = (int16_t)(t[0] + t[1] + t[2] + t[3] + t[4]);
    }
}

__global__ void __launch_bounds__(TPB) k_hash(const void* states, int64_t n, int mode, uint64_t* out) {
    int64_t g = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (g >= n) return;
    State s;
    load_state(s, states, g);
    out[g] = canon_hash(s, mode);
}

__global__ void __launch_bounds__(TPB) k_outcome(const void* states, int64_t n, uint8_t* over, int8_t* outcome) {
    int64_t g = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (g >= n) return;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(states) + g * 32;
    State s;
    s.w[HZ_W_BAG1META] = w[HZ_W_BAG1META];
    if (over) over[g] = is_over(s) ? 1 : 0;
    if (outcome) outcome[g] = (int8_t)(is_over(s) ? outcome_of(s) : 0);
}

__global__ void __launch_bounds__(TPB) k_random_actions(const void* states, int64_t n, int16_t* actions) {
    int64_t g = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (g >= n) return;
    State s;
    load_state(s, states, g);
    actions[g] = (int16_t)random_action(s, legal_of(s));
}

// Fused playout (K9): the whole game stays in registers; HBM sees one load and one store of
// the state per game.  Threads of a warp run different games and diverge only on the phase
// (choose / place / end of turn); scoring runs once per game.
constexpr int PTPB = 64;   // 65,536 games -> 1024 blocks: 6.9 per SM, <2 % tail imbalance over 148 SMs
__global__ void __launch_bounds__(PTPB) k_playout(void* states, int64_t n, int max_steps, uint32_t* steps,
                                                  unsigned long long* total_steps) {
    __shared__ NbrLut lut;
    __shared__ uint64_t rtab[RTAB_N];
    build_nbr_lut(&lut);
    build_rand_table(rtab);
    __syncthreads();
    int64_t g = (int64_t)blockIdx.x * PTPB + threadIdx.x;
    uint32_t k = 0;
    if (g < n) {
        State s;
        load_state(s, states, g);
        if (player_of(s)) swap_boards(s);   // mover-relative board order inside the loop (see REL)
        while ((int)k < max_steps && phase_of(s) != HZ_PHASE_OVER) {
            int a = random_action(s, legal_of<true>(s), rtab);
            if (a < 0) break;  // stuck position (no legal move, not over): harmonies_engine.py:205-208
            if (apply_move<true, true>(s, a, HZ_NO_DRAW, key_of(s), s.w[HZ_W_EVENT], true, &lut, rtab) != HZ_MOVE_OK) break;
            k++;
        }
        if (player_of(s)) swap_boards(s);   // back to absolute order
        // final scoring deferred to here: the lanes of the warp are converged again
        if (phase_of(s) == HZ_PHASE_OVER && winner_code(s) == 0) finalize_scores(s, &lut);
        store_state(s, states, g);
        if (steps) steps[g] = k;
    }
    if (total_steps) {
        uint32_t sum = k;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
        if ((threadIdx.x & 31) == 0 && sum) atomicAdd(total_steps, (unsigned long long)sum);
    }
}

// ---- encode (process_game_state.py:15-137) ------------------------------------------------------
constexpr int ENC_S = 8;      // states per block iteration
constexpr int ENC_TPB = 256;
template <typename T, bool NHWC>
__global__ void __launch_bounds__(ENC_TPB) k_encode(const void* states, int64_t n, T* board, T* glob, int vec_ok) {
    __shared__ uint32_t smask[ENC_S][40];
    __shared__ float sphase[ENC_S];
    __shared__ float sglob[ENC_S][42];
    __shared__ uint8_t scell[36];   // shared copy: per-lane indices into __constant__ would serialise
    if (threadIdx.x < 35) scell[threadIdx.x] = CELL_HEX[threadIdx.x];
    const uint32_t* W = reinterpret_cast<const uint32_t*>(states);
    int64_t n_chunks = (n + ENC_S - 1) / ENC_S;
    for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        int64_t base = chunk * ENC_S;
        int cnt = (int)min((int64_t)ENC_S, n - base);
        for (int item = threadIdx.x; item < cnt * 80; item += ENC_TPB) {
            int s = item / 80, c = item - 80 * s;
            const uint32_t* w = W + (base + s) * 32;
            if (c < 38) {
                smask[s][c] = channel_mask(w, c);
                if (c == 37) sphase[s] = (float)((double)((w[HZ_W_BAG1META] >> 25) & 7u) / 3.0);
            } else {
                sglob[s][c - 38] = global_feature(w, c - 38);
            }
        }
        __syncthreads();
        // Write-bound part: every thread emits one 16-byte vector (4 fp32 / 8 bf16) per store.
        // A chunk of 8 states is 8*1330 elements = a whole number of vectors and starts
        // 16-byte aligned; a ragged last chunk falls back to scalar stores.
        T* bout = board + base * 1330;
        constexpr int VEC = 16 / (int)sizeof(T);
        int total = cnt * 1330;
        int nvec = (cnt == ENC_S && vec_ok) ? total / VEC : 0;
        for (int v = threadIdx.x; v < nvec; v += ENC_TPB) {
            int e = v * VEC;
            int s = e / 1330, r = e - 1330 * s;
            int c, cell;
            if (NHWC) { cell = r / 38; c = r - 38 * cell; } else { c = r / 35; cell = r - 35 * c; }
            alignas(16) T vals[VEC];
#pragma unroll
            for (int j = 0; j < VEC; j++) {
                uint32_t bit = (smask[s][c] >> scell[cell]) & 1u;    // bit 31 is never set: masked cells
                vals[j] = cvt<T>(bit ? (c == 37 ? sphase[s] : 1.0f) : 0.0f);
                if (NHWC) { if (++c == 38) { c = 0; if (++cell == 35) { cell = 0; s++; } } }
                else      { if (++cell == 35) { cell = 0; if (++c == 38) { c = 0; s++; } } }
            }
            reinterpret_cast<uint4*>(bout)[v] = *reinterpret_cast<const uint4*>(vals);
        }
        for (int e = nvec * VEC + threadIdx.x; e < total; e += ENC_TPB) {
            int s = e / 1330, r = e - 1330 * s;
            int c, cell;
            if (NHWC) { cell = r / 38; c = r - 38 * cell; } else { c = r / 35; cell = r - 35 * c; }
            uint32_t bit = (smask[s][c] >> scell[cell]) & 1u;
            bout[e] = cvt<T>(bit ? (c == 37 ? sphase[s] : 1.0f) : 0.0f);
        }
        T* gout = glob + base * 42;
        for (int e = threadIdx.x; e < cnt * 42; e += ENC_TPB) gout[e] = cvt<T>(sglob[e / 42][e % 42]);
        __syncthreads();
    }
}

}  // namespace hz

using namespace hz;

static inline int blocks_for(int64_t n, int tpb) { return (int)((n + tpb - 1) / tpb); }

extern "C" {

int hz_init_states(void* states, int64_t n, const uint64_t* keys, uint64_t seed, uint64_t first_id, void* stream) {
    if (n == 0) return HZ_OK;   // empty batch: nothing to do, pointers may be null
    if (!states || n < 0) return HZ_ERR_ARG;
    if (n == 0) return HZ_OK;
    k_init<<<blocks_for(n, TPB), TPB, 0, (cudaStream_t)stream>>>(states, n, keys, seed, first_id);
    return hz_launched(1);
}

int hz_legal_mask(const void* states, int64_t n, uint32_t* mask, void* stream) {
    if (n == 0) return HZ_OK;   // empty batch: nothing to do, pointers may be null
    if (!states || !mask || n < 0) return HZ_ERR_ARG;
    if (n == 0) return HZ_OK;
    k_legal<<<blocks_for(n, TPB), TPB, 0, (cudaStream_t)stream>>>(states, n, mask);
    return hz_launched(1);
}

int hz_apply(void* states, int64_t n, const int16_t* actions, const uint16_t* draws, uint8_t* status, void* stream) {
    if (n == 0) return HZ_OK;   // empty batch: nothing to do, pointers may be null
    if (!states || !actions || n < 0) return HZ_ERR_ARG;
    if (n == 0) return HZ_OK;
    k_apply<<<blocks_for(n, TPB), TPB, 0, (cudaStream_t)stream>>>(states, n, actions, draws, status);
    return hz_launched(1);
}

int hz_score(const void* states, int64_t n, int16_t* scores, int16_t* terms, void* stream) {
    if (n == 0) return HZ_OK;   // empty batch: nothing to do, pointers may be null
    if (!states || (!scores && !terms) || n < 0) return HZ_ERR_ARG;
    if (n == 0) return HZ_OK;
    k_score<<<blocks_for(n, TPB), TPB, 0, (cudaStream_t)stream>>>(states, n, scores, terms);
    return hz_launched(1);
}

int hz_encode(const void* states, int64_t n, void* board, void* glob, int dtype, int layout, void* stream) {
    if (n == 0) return HZ_OK;   // empty batch: nothing to do, pointers may be null
    if (!states || !board || !glob || n < 0) return HZ_ERR_ARG;
    if (dtype != HZ_DTYPE_F32 && dtype != HZ_DTYPE_BF16) return HZ_ERR_ARG;
    if (layout != HZ_LAYOUT_NCHW && layout != HZ_LAYOUT_NHWC) return HZ_ERR_ARG;
    if (n == 0) return HZ_OK;
    int64_t chunks = (n + ENC_S - 1) / ENC_S;
    int grid = (int)(chunks < 148 * 8 ? chunks : 148 * 8);
    cudaStream_t st = (cudaStream_t)stream;
    int vec_ok = ((uintptr_t)board & 15) == 0;   // 16-byte vector stores need an aligned base
    if (dtype == HZ_DTYPE_F32) {
        if (layout == HZ_LAYOUT_NCHW) k_encode<float, false><<<grid, ENC_TPB, 0, st>>>(states, n, (float*)board, (float*)glob, vec_ok);
        else k_encode<float, true><<<grid, ENC_TPB, 0, st>>>(states, n, (float*)board, (float*)glob, vec_ok);
    } else {
        if (layout == HZ_LAYOUT_NCHW) k_encode<__nv_bfloat16, false><<<grid, ENC_TPB, 0, st>>>(states, n, (__nv_bfloat16*)board, (__nv_bfloat16*)glob, vec_ok);
        else k_encode<__nv_bfloat16, true><<<grid, ENC_TPB, 0, st>>>(states, n, (__nv_bfloat16*)board, (__nv_bfloat16*)glob, vec_ok);
    }
    return hz_launched(1);
}

int hz_canon_hash(const void* states, int64_t n, int key_mode, uint64_t* hashes, void* stream) {
    if (n == 0) return HZ_OK;   // empty batch: nothing to do, pointers may be null
    if (!states || !hashes || n < 0 || (key_mode != HZ_KEY_EXACT && key_mode != HZ_KEY_REFERENCE)) return HZ_ERR_ARG;
    if (n == 0) return HZ_OK;
    k_hash<<<blocks_for(n, TPB), TPB, 0, (cudaStream_t)stream>>>(states, n, key_mode, hashes);
    return hz_launched(1);
}

int hz_outcome(const void* states, int64_t n, uint8_t* over, int8_t* outcome, void* stream) {
    if (n == 0) return HZ_OK;   // empty batch: nothing to do, pointers may be null
    if (!states || (!over && !outcome) || n < 0) return HZ_ERR_ARG;
    if (n == 0) return HZ_OK;
    k_outcome<<<blocks_for(n, TPB), TPB, 0, (cudaStream_t)stream>>>(states, n, over, outcome);
    return hz_launched(1);
}

int hz_random_actions(const void* states, int64_t n, int16_t* actions, void* stream) {
    if (n == 0) return HZ_OK;   // empty batch: nothing to do, pointers may be null
    if (!states || !actions || n < 0) return HZ_ERR_ARG;
    if (n == 0) return HZ_OK;
    k_random_actions<<<blocks_for(n, TPB), TPB, 0, (cudaStream_t)stream>>>(states, n, actions);
    return hz_launched(1);
}

int hz_playout(void* states, int64_t n, int max_steps, uint32_t* steps, unsigned long long* total_steps, void* stream) {
    if (n == 0) return HZ_OK;   // empty batch: nothing to do, pointers may be null
    if (!states || n < 0 || max_steps < 0) return HZ_ERR_ARG;
    if (n == 0) return HZ_OK;
    k_playout<<<blocks_for(n, PTPB), PTPB, 0, (cudaStream_t)stream>>>(states, n, max_steps, steps, total_steps);
    return hz_launched(1);
}

}  // extern "C"

// hz_engine.cu — batched engine kernels (sm_100a) behind the C ABI of
// include/harmonies_b200.h: new games, legal masks, move application, scoring, state
// encoding, canonical keys, outcomes, random-playout policy and the fused playout.
//
// Mapping: one thread per game, state held in registers (7 x LDG.128 in, 8 x STG.128 out),
// hex-parallel work done as 23-bit bit-board arithmetic; the neighbour expansion LUT lives in
// 3 KB of shared memory per block.  hz_encode is write-bound (5.5 KB out per 128 B in): a warp
// renders a group of records as a bit stream and stores 16-byte vectors through a lookup table
// (k_encode_w); hz_greedy_actions is warp-per-game with one lane per legal move.
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
    __shared__ __align__(16) uint4 tile[TPB / 32][256];
    int64_t g = (int64_t)blockIdx.x * TPB + threadIdx.x;
    State s;
    load_state_warp<6>(s, states, g - (threadIdx.x & 31), n, tile[threadIdx.x >> 5]);   // legal moves read words 0..22 only
    uint32_t out[5];
    legal_words(legal_of(s), out);
    // the warp's 32 x 5 mask words leave as five coalesced 128-byte stores, through the same tile (stride 5: conflict-free)
    uint32_t* tw = reinterpret_cast<uint32_t*>(tile[threadIdx.x >> 5]);
    const int lane = threadIdx.x & 31;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 5; k++) tw[lane * 5 + k] = out[k];
    __syncwarp();
    const int64_t g0 = g - lane, lim = (n - g0) * 5;
#pragma unroll
    for (int k = 0; k < 5; k++)
        if (k * 32 + lane < lim) mask[g0 * 5 + k * 32 + lane] = tw[k * 32 + lane];
}

__global__ void __launch_bounds__(TPB) k_apply(void* states, int64_t n, const int16_t* actions,
                                               const uint16_t* draws, uint8_t* status) {
    __shared__ NbrLut lut;
    int64_t g = (int64_t)blockIdx.x * TPB + threadIdx.x;
    State s;
    if (g < n) load_state(s, states, g);      // issue the record loads first: the LUT copy overlaps them
    build_nbr_lut(&lut);
    __syncthreads();
    if (g >= n) return;
    uint32_t ex = draws ? (uint32_t)draws[g] : (uint32_t)HZ_NO_DRAW;
    int st = apply_move(s, (int)actions[g], ex, key_of(s), s.w[HZ_W_EVENT], true, &lut);
    if (st == HZ_MOVE_OK) store_state(s, states, g);
    if (status) status[g] = (uint8_t)st;
}

// ---- the reference's private turn helpers, for callers that drive them directly (GUI/main.py:364-365,
// harness code): _end_turn_actions (:301-329), _replenish_piles (:132-137), _draw_tiles (:120-130)
__global__ void __launch_bounds__(TPB) k_end_turn(void* states, int64_t n, const uint16_t* draws, uint8_t* status) {
    __shared__ NbrLut lut;
    int64_t g = (int64_t)blockIdx.x * TPB + threadIdx.x;
    State s;
    if (g < n) load_state(s, states, g);
    build_nbr_lut(&lut);
    __syncthreads();
    if (g >= n) return;
    uint32_t ex = draws ? (uint32_t)draws[g] : (uint32_t)HZ_NO_DRAW;
    bool use_explicit = ex != HZ_NO_DRAW && n_piles_of(s) < 5;
    int st = HZ_MOVE_OK;
    if (use_explicit && !pile_available(bag_of(s), ex)) st = HZ_MOVE_BAD_DRAW;
    else {
        int pl = player_of(s);
        Tops t = tops_of(board_of(s, pl));
        st = end_turn<false, false>(s, pl, t.occ0, hand_of(s), use_explicit, ex, key_of(s), s.w[HZ_W_EVENT], true, &lut, nullptr);
        store_state(s, states, g);
    }
    if (status) status[g] = (uint8_t)st;
}

__global__ void __launch_bounds__(TPB) k_replenish(void* states, int64_t n) {
    int64_t g = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (g >= n) return;
    State s;
    load_state(s, states, g);
    int np;
    replenish_piles(s, hand_of(s), false, 0u, key_of(s), s.w[HZ_W_EVENT], nullptr, np);
    s.w[HZ_W_EVENT]++;
    store_state(s, states, g);
}

// draws min(count, bag total) tiles one by one, uniformly without replacement from the bag multiset
// (random.sample over the flattened bag, :122-129), with the arithmetic of draw_pile; tiles[g][0..14] =
// tile types in draw order, tiles[g][15] = how many were drawn.  One draw event of the state's stream.
__global__ void __launch_bounds__(TPB) k_draw_tiles(void* states, int64_t n, int count, uint8_t* tiles) {
    int64_t g = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (g >= n) return;
    State s;
    load_state(s, states, g);
    uint64_t bag = bag_of(s), key = key_of(s);
    int total = bag_total(bag), drawn = 0;
    for (int grp = 0; grp < 5 && drawn < count && total > 0; grp++) {
        uint64_t z = rand64(key, (uint64_t)s.w[HZ_W_EVENT] * 8 + (uint64_t)grp);
        for (int j = 0; j < 3 && drawn < count && total > 0; j++) {
            uint32_t x = (uint32_t)(z >> (21 * j)) & 0x1FFFFFu;
            uint32_t r = (x * (uint32_t)total) >> 21;
            int t = 0;
            uint32_t cum = 0;
            for (; t < 5; t++) {
                cum += (uint32_t)(bag >> (8 * t)) & 0xFFu;
                if (cum > r) break;
            }
            bag -= 1ull << (8 * t);
            total--;
            tiles[g * 16 + drawn++] = (uint8_t)t;
        }
    }
    tiles[g * 16 + 15] = (uint8_t)drawn;
    set_bag(s, bag);
    s.w[HZ_W_EVENT]++;
    store_state(s, states, g);
}

__global__ void __launch_bounds__(TPB) k_score(const void* states, int64_t n, int16_t* scores,
                                               int16_t* terms) {
    __shared__ NbrLut lut;
    __shared__ WaterQueue<2 * TPB, 2 * TPB> wq;
    int64_t g = (int64_t)blockIdx.x * TPB + threadIdx.x;
    // only the 18 board words are needed: 92 -> 80 B of traffic per position; the loads are
    // issued before the LUT copy so that the two overlap
    uint32_t w[20];
    if (g < n) {
        const uint4* p = reinterpret_cast<const uint4*>(states) + g * 8;
#pragma unroll
        for (int k = 0; k < 5; k++) {
            uint4 v = p[k];
            w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
        }
    }
    build_nbr_lut(&lut);
    water_queue_init(&wq);
    __syncthreads();
    int t[2][5];
    if (g < n) {
#pragma unroll
        for (int pl = 0; pl < 2; pl++) {
            Board b;
#pragma unroll
            for (int k = 0; k < 9; k++) b.p[k] = w[pl * 9 + k];
            score_board(&lut, b, t[pl], &wq, (int)threadIdx.x * 2 + pl);
        }
    }
    __syncthreads();
    resolve_water(&lut, &wq);          // the rare large water components, one warp each
    __syncthreads();
    if (g >= n) return;
#pragma unroll
    for (int pl = 0; pl < 2; pl++) {
        t[pl][4] += wq.extra[threadIdx.x * 2 + pl];
        if (terms) {
#pragma unroll
            for (int k = 0; k < 5; k++) terms[(g * 2 + pl) * 5 + k] = (int16_t)t[pl][k];
        }
        if (scores) scores[g * 2 + pl] = (int16_t)(t[pl][0] + t[pl][1] + t[pl][2] + t[pl][3] + t[pl][4]);
    }
}

// (the warp-staged load of k_legal was measured here too: 32.8 us against 29.7 us per 1 M records — the hash has enough
// arithmetic per record to hide the direct loads, and the staging adds a barrier in front of it)
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

// choose_move_greedy (evaluation.py:137-196): one warp per game, one lane per legal move; a lane
// places its tile on a copy of the mover's board, scores it (K2+K3 fused) and the warp keeps the
// first strict maximum in ascending action order.  Taking a pile never changes the mover's
// score, so in the choose phase the first pile wins, as in the reference's loop.
constexpr int GWPB = 4;
__global__ void __launch_bounds__(GWPB * 32) k_greedy(const void* states, int64_t n, int16_t* actions) {
    __shared__ NbrLut lut;
    __shared__ uint32_t sw[GWPB][32];
    build_nbr_lut(&lut);
    __syncthreads();
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int64_t g = (int64_t)blockIdx.x * GWPB + warp;
    if (g >= n) return;
    sw[warp][lane] = reinterpret_cast<const uint32_t*>(states)[g * 32 + lane];
    __syncwarp();
    State s;
#pragma unroll
    for (int i = 0; i < SW; i++) s.w[i] = sw[warp][i];
    Legal L = legal_of(s);
    int nl = legal_count(L), best_k = -1, best_score = -1000000;
    if (L.n_piles > 0) {
        best_k = 0;
    } else {
        Board b0 = board_of(s, player_of(s));
        Tops t = tops_of(b0);
        for (int base = 0; base < nl; base += 32) {
            int k = base + lane;
            if (k < nl) {
                int a = kth_action(L, k);
                int tile = (a - 5) / 23, hex = (a - 5) - 23 * tile;
                uint32_t bit = 1u << hex, code = (uint32_t)tile + 1u;
                uint32_t at0 = bit & ~t.occ0, at1 = bit & t.occ0 & ~t.occ1, at2 = bit & t.occ1;
                Board b = b0;
#pragma unroll
                for (int q = 0; q < 3; q++) {
                    uint32_t on = (code >> q) & 1u ? 0xFFFFFFFFu : 0u;
                    b.p[q] |= at0 & on; b.p[3 + q] |= at1 & on; b.p[6 + q] |= at2 & on;
                }
                int tm[5];
                score_board(&lut, b, tm);
                int sc = tm[0] + tm[1] + tm[2] + tm[3] + tm[4];
                if (sc > best_score) { best_score = sc; best_k = k; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            int os = __shfl_xor_sync(0xFFFFFFFFu, best_score, o), ok = __shfl_xor_sync(0xFFFFFFFFu, best_k, o);
            bool take = ok >= 0 && (best_k < 0 || os > best_score || (os == best_score && ok < best_k));
            if (take) { best_score = os; best_k = ok; }
        }
    }
    if (lane == 0) actions[g] = (int16_t)(best_k < 0 ? -1 : kth_action(L, best_k));
}

// Fused playout (K9): the whole game stays in registers; HBM sees one load and one store of
// the state per game.  Threads of a warp run different games and diverge only on the phase
// (choose / place / end of turn); scoring runs once per game.
constexpr int PTPB = 64;   // 65,536 games -> 1024 blocks: 6.9 per SM, <2 % tail imbalance over 148 SMs
// FROM_KEYS: the game is created in registers from its 64-bit key (k_init fused in) and only
// (meta word, score word, actions played) leave the chip: 8 B in, 12 B out per game.
template <bool FROM_KEYS>
__global__ void __launch_bounds__(PTPB) k_playout(void* states, int64_t n, int max_steps, uint32_t* steps,
                                                  unsigned long long* total_steps, const uint64_t* keys,
                                                  uint64_t seed, uint64_t first_id, uint32_t* results) {
    __shared__ NbrLut lut;
    __shared__ uint64_t rtab_s[RTAB_N];
    __shared__ WaterQueue<2 * PTPB, 2 * PTPB> wq;
    build_nbr_lut(&lut);
    build_rand_table(rtab_s);
    const RandTab rtab = rand_table_handle(rtab_s);
    water_queue_init(&wq);
    __syncthreads();
    int64_t g = (int64_t)blockIdx.x * PTPB + threadIdx.x;
    uint32_t k = 0;
    State s;
    bool fin = false;
    int s0 = 0, s1 = 0;
    if (g < n) {
        if (FROM_KEYS) init_state(s, keys ? keys[g] : rand64(seed, first_id + (uint64_t)g), rtab);
        else load_state(s, states, g);
#ifdef HZ_PLAYOUT_CHECKED   // A/B: the three general calls on a mover-relative state (validation + board copy in apply_move)
        if (player_of(s)) swap_boards(s);
        while ((int)k < max_steps && phase_of(s) != HZ_PHASE_OVER) {
            int a = random_action(s, legal_of<true>(s), rtab);
            if (a < 0) break;  // stuck position (no legal move, not over): harmonies_engine.py:205-208
            if (apply_move<true, true>(s, a, HZ_NO_DRAW, key_of(s), s.w[HZ_W_EVENT], true, &lut, rtab) != HZ_MOVE_OK) break;
            k++;
        }
        if (player_of(s)) swap_boards(s);
#else
        // one branch on the player per step: fresh games move in lockstep (same player, same phase in every lane until games
        // end), and inside a branch the mover's planes are fixed registers
        while ((int)k < max_steps && phase_of(s) != HZ_PHASE_OVER) {
            const bool ok = player_of(s) ? playout_step<1, true>(s, &lut, rtab) : playout_step<0, true>(s, &lut, rtab);
            if (!ok) break;  // stuck position (no legal move, not over): harmonies_engine.py:205-208
            k++;
        }
#endif
        // final scoring deferred to here: the lanes of the warp are converged again; rivers of >= 5
        // hexes go to the block's queue and get a whole warp each (resolve_water)
        fin = phase_of(s) == HZ_PHASE_OVER && winner_code(s) == 0;
        if (fin) partial_scores(s, &lut, &wq, s0, s1);
    }
    __syncthreads();
    resolve_water(&lut, &wq);
    __syncthreads();
    if (g < n) {
        if (fin) store_final_scores(s, s0 + wq.extra[threadIdx.x * 2], s1 + wq.extra[threadIdx.x * 2 + 1]);
        if (FROM_KEYS) {
            results[g * 3] = s.w[HZ_W_BAG1META];
            results[g * 3 + 1] = s.w[HZ_W_SCORES];
            results[g * 3 + 2] = s.w[HZ_W_MOVES];
        } else {
            store_state(s, states, g);
            if (steps) steps[g] = k;
        }
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
    __shared__ uint8_t shex[24];
    __shared__ __align__(16) uint32_t sw[ENC_S][32];
    // The tensor is ~94 % zeros and every non-zero is 1.0 except the phase plane, so the chunk
    // is first rendered as a BIT stream in output order (1 bit per element, 10,640 bits) and a
    // store is then one table lookup: VEC bits -> 16 bytes of 0.0/1.0 (then the phase patch).
    constexpr int VEC = 16 / (int)sizeof(T);
    __shared__ uint32_t stream[ENC_S * 1330 / 32 + 4], pstream[ENC_S * 1330 / 32 + 4];
    __shared__ uint64_t perm[3][256];   // hex-order byte -> cell-order bits (NCHW rows)
    __shared__ uint4 vlut[1 << VEC];
    __shared__ uint32_t sphase_bits[ENC_S];
    if (threadIdx.x < 35) scell[threadIdx.x] = CELL_HEX[threadIdx.x];
    if (threadIdx.x < 23) shex[threadIdx.x] = HEX_CELL[threadIdx.x];
    for (int i = threadIdx.x; i < (1 << VEC); i += ENC_TPB) {
        uint32_t o[4];
        if (VEC == 4) {
#pragma unroll
            for (int j = 0; j < 4; j++) o[j] = ((i >> j) & 1) ? 0x3F800000u : 0u;
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) o[j] = (((i >> (2 * j)) & 1) ? 0x3F80u : 0u) | (((i >> (2 * j + 1)) & 1) ? 0x3F800000u : 0u);
        }
        vlut[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    if (!NHWC) {
        for (int e = threadIdx.x; e < 768; e += ENC_TPB) {
            int c = e >> 8, v = e & 255;
            uint64_t m = 0;
            for (int b = 0; b < 8; b++) {
                int i = c * 8 + b;
                if (((v >> b) & 1) && i < 23) m |= 1ull << HEX_CELL[i];
            }
            perm[c][v] = m;
        }
    }
    const uint32_t* W = reinterpret_cast<const uint32_t*>(states);
    int64_t n_chunks = (n + ENC_S - 1) / ENC_S;
    for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        int64_t base = chunk * ENC_S;
        int cnt = (int)min((int64_t)ENC_S, n - base);
        for (int i = threadIdx.x; i < ENC_S * 1330 / 32 + 4; i += ENC_TPB) { stream[i] = 0; pstream[i] = 0; }
        // stage the chunk's records (1 KB) in shared memory with one coalesced 16-byte load per
        // thread: the 640 mask/feature tasks below each read 1-3 words of a record
        if (threadIdx.x < cnt * 8)
            reinterpret_cast<uint4*>(&sw[0][0])[threadIdx.x] = reinterpret_cast<const uint4*>(W)[base * 8 + threadIdx.x];
        __syncthreads();
        for (int item = threadIdx.x; item < cnt * 80; item += ENC_TPB) {
            int s = item / 80, c = item - 80 * s;
            const uint32_t* w = sw[s];
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
        int total = cnt * 1330;
        int nvec = (cnt == ENC_S && vec_ok) ? total / VEC : 0;
        if (nvec) {
            // ---- render the bit stream: one task per (state, channel) [NCHW] or (state, cell) [NHWC]
            if (threadIdx.x < ENC_S) {
                float pv = sphase[threadIdx.x];
                sphase_bits[threadIdx.x] = VEC == 4 ? __float_as_uint(pv) : (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(pv));
            }
            constexpr int ROWS = NHWC ? 35 : 38, ROW_BITS = NHWC ? 38 : 35;
            for (int item = threadIdx.x; item < ENC_S * ROWS; item += ENC_TPB) {
                int s = item / ROWS, row = item - ROWS * s;
                uint64_t bits = 0;
                if (NHWC) {
                    int hx = scell[row];
                    if (hx < 23) {
                        // a cell has at most 3+3 tiles: read the tile code of each (player, level)
                        // and set its channel bit (p*18 + type*3 + level) instead of probing 38 masks
                        const uint32_t* w = sw[s];
#pragma unroll
                        for (int pl = 0; pl < 6; pl++) {
                            uint32_t code = ((w[pl * 3] >> hx) & 1u) | (((w[pl * 3 + 1] >> hx) & 1u) << 1) | (((w[pl * 3 + 2] >> hx) & 1u) << 2);
                            int p = pl / 3, l = pl - 3 * p;
                            if (code) bits |= 1ull << (p * 18 + ((int)code - 1) * 3 + l);
                        }
                        bits |= (uint64_t)((smask[s][36] >> hx) & 1u) << 36;
                        bits |= (uint64_t)((smask[s][37] >> hx) & 1u) << 37;
                    }
                } else {
                    uint32_t m = smask[s][row];                      // hex order -> cell order: 3 table lookups
                    bits = perm[0][m & 255] | perm[1][(m >> 8) & 255] | perm[2][(m >> 16) & 127];
                }
                if (bits) {
                    int off = s * 1330 + row * ROW_BITS, wd = off >> 5, sh = off & 31;
                    uint64_t lo = bits << sh;
                    uint32_t hi = sh ? (uint32_t)(bits >> (64 - sh)) : 0u;
                    if ((uint32_t)lo) atomicOr(&stream[wd], (uint32_t)lo);
                    if ((uint32_t)(lo >> 32)) atomicOr(&stream[wd + 1], (uint32_t)(lo >> 32));
                    if (hi) atomicOr(&stream[wd + 2], hi);
                    // second stream: the set elements of the phase plane (channel 37, value k/3)
                    uint64_t pbits = NHWC ? (bits & (1ull << 37)) : (row == 37 ? bits : 0ull);
                    if (pbits) {
                        uint64_t plo = pbits << sh;
                        uint32_t phi = sh ? (uint32_t)(pbits >> (64 - sh)) : 0u;
                        if ((uint32_t)plo) atomicOr(&pstream[wd], (uint32_t)plo);
                        if ((uint32_t)(plo >> 32)) atomicOr(&pstream[wd + 1], (uint32_t)(plo >> 32));
                        if (phi) atomicOr(&pstream[wd + 2], phi);
                    }
                }
            }
            __syncthreads();
        }
        for (int v = threadIdx.x; v < nvec; v += ENC_TPB) {
            int e0 = v * VEC;
            uint32_t nib = (stream[e0 >> 5] >> (e0 & 31)) & ((1u << VEC) - 1u);   // VEC divides 32: never straddles
            uint32_t pm = (pstream[e0 >> 5] >> (e0 & 31)) & ((1u << VEC) - 1u);
            uint4 o = vlut[nib];
            while (pm) {                                             // rare: phase-plane elements hold k/3, not 1.0
                int j = __ffs(pm) - 1;
                pm &= pm - 1;
                uint32_t pb = sphase_bits[(e0 + j) / 1330];
                int wi = VEC == 4 ? j : (j >> 1);
                uint32_t cur = wi == 0 ? o.x : wi == 1 ? o.y : wi == 2 ? o.z : o.w;
                uint32_t nw = VEC == 4 ? pb : ((j & 1) ? ((cur & 0x0000FFFFu) | (pb << 16)) : ((cur & 0xFFFF0000u) | pb));
                o.x = wi == 0 ? nw : o.x; o.y = wi == 1 ? nw : o.y; o.z = wi == 2 ? nw : o.z; o.w = wi == 3 ? nw : o.w;
            }
            reinterpret_cast<uint4*>(bout)[v] = o;
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

// Fast path of hz_encode: one WARP renders a group of G records (G*1330 elements = 665 16-byte
// vectors: G = 2 for fp32, 4 for bf16) with no block-level barrier at all — each warp has its
// own staging area and only __syncwarp()s, so the SM always has other warps to issue from.
// Same bit-stream + table-lookup scheme as k_encode (which remains the tail / unaligned path).
constexpr int ENCW_WARPS = 8;
template <typename T, bool NHWC>
__global__ void __launch_bounds__(ENCW_WARPS * 32) k_encode_w(const void* states, int64_t n_groups, T* board, T* glob) {
    constexpr int VEC = 16 / (int)sizeof(T);
    constexpr int G = VEC / 2;                       // 2 (fp32) or 4 (bf16) records per group
    constexpr int NV = G * 1330 / VEC;               // 665 vectors per group
    constexpr int SWORDS = (G * 1330 + 31) / 32 + 3;
    constexpr int ROWS = NHWC ? 35 : 38, ROW_BITS = NHWC ? 38 : 35;
    __shared__ __align__(16) uint32_t sw[ENCW_WARPS][G][32];
    __shared__ uint32_t smask[ENCW_WARPS][G][40];
    __shared__ uint32_t stream[ENCW_WARPS][SWORDS], pstream[ENCW_WARPS][SWORDS];
    __shared__ uint32_t sphase_bits[ENCW_WARPS][G];
    __shared__ uint64_t perm[3][256];
    __shared__ uint4 vlut[1 << VEC];
    __shared__ uint8_t scell[36];
    if (threadIdx.x < 35) scell[threadIdx.x] = CELL_HEX[threadIdx.x];
    for (int i = threadIdx.x; i < (1 << VEC); i += ENCW_WARPS * 32) {
        uint32_t o[4];
        if (VEC == 4) {
#pragma unroll
            for (int j = 0; j < 4; j++) o[j] = ((i >> j) & 1) ? 0x3F800000u : 0u;
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) o[j] = (((i >> (2 * j)) & 1) ? 0x3F80u : 0u) | (((i >> (2 * j + 1)) & 1) ? 0x3F800000u : 0u);
        }
        vlut[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    if (!NHWC) {
        for (int e = threadIdx.x; e < 768; e += ENCW_WARPS * 32) {
            int c = e >> 8, v = e & 255;
            uint64_t m = 0;
            for (int b = 0; b < 8; b++) {
                int i = c * 8 + b;
                if (((v >> b) & 1) && i < 23) m |= 1ull << HEX_CELL[i];
            }
            perm[c][v] = m;
        }
    }
    __syncthreads();                                  // the only block barrier: tables are ready
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* my_stream = stream[warp];
    uint32_t* my_pstream = pstream[warp];
    const int64_t n_warps = (int64_t)gridDim.x * ENCW_WARPS;
    for (int64_t grp = (int64_t)blockIdx.x * ENCW_WARPS + warp; grp < n_groups; grp += n_warps) {
        const int64_t base = grp * G;
        for (int i = lane; i < G * 8; i += 32)
            reinterpret_cast<uint4*>(&sw[warp][0][0])[i] = reinterpret_cast<const uint4*>(states)[base * 8 + i];
        for (int i = lane; i < SWORDS; i += 32) { my_stream[i] = 0; my_pstream[i] = 0; }
        __syncwarp();
        // channel masks (hex order) and the 42 global features of each record
        for (int item = lane; item < G * 80; item += 32) {
            int s = item / 80, c = item - 80 * s;
            const uint32_t* w = sw[warp][s];
            if (c < 38) smask[warp][s][c] = channel_mask(w, c);
            else glob[(base + s) * 42 + (c - 38)] = cvt<T>(global_feature(w, c - 38));
        }
        if (lane < G) {
            uint32_t ph = (sw[warp][lane][HZ_W_BAG1META] >> 25) & 7u;
            float pv = ph == 1 ? (float)(1.0 / 3.0) : ph == 2 ? (float)(2.0 / 3.0) : ph == 3 ? 1.0f : 0.0f;
            sphase_bits[warp][lane] = VEC == 4 ? __float_as_uint(pv) : (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(pv));
        }
        __syncwarp();
        // bit streams in output order
        for (int item = lane; item < G * ROWS; item += 32) {
            int s = item / ROWS, row = item - ROWS * s;
            uint64_t bits = 0;
            if (NHWC) {
                int hx = scell[row];
                if (hx < 23) {
                    const uint32_t* w = sw[warp][s];
#pragma unroll
                    for (int pl = 0; pl < 6; pl++) {
                        uint32_t code = ((w[pl * 3] >> hx) & 1u) | (((w[pl * 3 + 1] >> hx) & 1u) << 1) | (((w[pl * 3 + 2] >> hx) & 1u) << 2);
                        int p = pl / 3, l = pl - 3 * p;
                        if (code) bits |= 1ull << (p * 18 + ((int)code - 1) * 3 + l);
                    }
                    bits |= (uint64_t)((smask[warp][s][36] >> hx) & 1u) << 36;
                    bits |= (uint64_t)((smask[warp][s][37] >> hx) & 1u) << 37;
                }
            } else {
                uint32_t m = smask[warp][s][row];
                bits = perm[0][m & 255] | perm[1][(m >> 8) & 255] | perm[2][(m >> 16) & 127];
            }
            if (bits) {
                int off = s * 1330 + row * ROW_BITS, wd = off >> 5, sh = off & 31;
                uint64_t lo = bits << sh;
                uint32_t hi = sh ? (uint32_t)(bits >> (64 - sh)) : 0u;
                if ((uint32_t)lo) atomicOr(&my_stream[wd], (uint32_t)lo);
                if ((uint32_t)(lo >> 32)) atomicOr(&my_stream[wd + 1], (uint32_t)(lo >> 32));
                if (hi) atomicOr(&my_stream[wd + 2], hi);
                uint64_t pbits = NHWC ? (bits & (1ull << 37)) : (row == 37 ? bits : 0ull);
                if (pbits) {
                    uint64_t plo = pbits << sh;
                    uint32_t phi = sh ? (uint32_t)(pbits >> (64 - sh)) : 0u;
                    if ((uint32_t)plo) atomicOr(&my_pstream[wd], (uint32_t)plo);
                    if ((uint32_t)(plo >> 32)) atomicOr(&my_pstream[wd + 1], (uint32_t)(plo >> 32));
                    if (phi) atomicOr(&my_pstream[wd + 2], phi);
                }
            }
        }
        __syncwarp();
        uint4* out = reinterpret_cast<uint4*>(board + base * 1330);   // G*1330*sizeof(T) is a multiple of 16
        for (int v = lane; v < NV; v += 32) {
            int e0 = v * VEC;
            uint32_t nib = (my_stream[e0 >> 5] >> (e0 & 31)) & ((1u << VEC) - 1u);
            uint32_t pm = (my_pstream[e0 >> 5] >> (e0 & 31)) & ((1u << VEC) - 1u);
            uint4 o = vlut[nib];
            while (pm) {                                             // phase-plane elements hold k/3, not 1.0
                int j = __ffs(pm) - 1;
                pm &= pm - 1;
                uint32_t pb = sphase_bits[warp][(e0 + j) / 1330];
                int wi = VEC == 4 ? j : (j >> 1);
                uint32_t cur = wi == 0 ? o.x : wi == 1 ? o.y : wi == 2 ? o.z : o.w;
                uint32_t nw = VEC == 4 ? pb : ((j & 1) ? ((cur & 0x0000FFFFu) | (pb << 16)) : ((cur & 0xFFFF0000u) | pb));
                o.x = wi == 0 ? nw : o.x; o.y = wi == 1 ? nw : o.y; o.z = wi == 2 ? nw : o.z; o.w = wi == 3 ? nw : o.w;
            }
            out[v] = o;
        }
        __syncwarp();
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

int hz_end_turn(void* states, int64_t n, const uint16_t* draws, uint8_t* status, void* stream) {
    if (n == 0) return HZ_OK;
    if (!states || n < 0) return HZ_ERR_ARG;
    k_end_turn<<<blocks_for(n, TPB), TPB, 0, (cudaStream_t)stream>>>(states, n, draws, status);
    return hz_launched(1);
}

int hz_replenish_piles(void* states, int64_t n, void* stream) {
    if (n == 0) return HZ_OK;
    if (!states || n < 0) return HZ_ERR_ARG;
    k_replenish<<<blocks_for(n, TPB), TPB, 0, (cudaStream_t)stream>>>(states, n);
    return hz_launched(1);
}

int hz_draw_tiles(void* states, int64_t n, int count, uint8_t* tiles, void* stream) {
    if (n == 0) return HZ_OK;
    if (!states || !tiles || n < 0 || count < 0 || count > 15) return HZ_ERR_ARG;
    k_draw_tiles<<<blocks_for(n, TPB), TPB, 0, (cudaStream_t)stream>>>(states, n, count, tiles);
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
    cudaStream_t st = (cudaStream_t)stream;
    int vec_ok = ((uintptr_t)board & 15) == 0;   // 16-byte vector stores need an aligned base
    // fast path: whole groups of 2 (fp32) / 4 (bf16) records, one warp per group, no block barriers
    int G = dtype == HZ_DTYPE_F32 ? 2 : 4;
    int64_t n_groups = vec_ok ? n / G : 0, done = n_groups * G;
    int launches = 0;
    if (n_groups) {
        int64_t blocks = (n_groups + ENCW_WARPS - 1) / ENCW_WARPS;
        int grid = (int)(blocks < 148 * 6 ? blocks : 148 * 6);
        if (dtype == HZ_DTYPE_F32) {
            if (layout == HZ_LAYOUT_NCHW) k_encode_w<float, false><<<grid, ENCW_WARPS * 32, 0, st>>>(states, n_groups, (float*)board, (float*)glob);
            else k_encode_w<float, true><<<grid, ENCW_WARPS * 32, 0, st>>>(states, n_groups, (float*)board, (float*)glob);
        } else {
            if (layout == HZ_LAYOUT_NCHW) k_encode_w<__nv_bfloat16, false><<<grid, ENCW_WARPS * 32, 0, st>>>(states, n_groups, (__nv_bfloat16*)board, (__nv_bfloat16*)glob);
            else k_encode_w<__nv_bfloat16, true><<<grid, ENCW_WARPS * 32, 0, st>>>(states, n_groups, (__nv_bfloat16*)board, (__nv_bfloat16*)glob);
        }
        launches++;
    }
    if (done < n) {   // tail (or unaligned output): the block kernel, scalar stores
        int64_t m = n - done;
        const char* sp = (const char*)states + done * 128;
        size_t esz = dtype == HZ_DTYPE_F32 ? 4 : 2;
        char* bp = (char*)board + (size_t)done * 1330 * esz;
        char* gp = (char*)glob + (size_t)done * 42 * esz;
        int64_t chunks = (m + ENC_S - 1) / ENC_S;
        int grid = (int)(chunks < 148 * 8 ? chunks : 148 * 8);
        int tail_vec = ((uintptr_t)bp & 15) == 0;
        if (dtype == HZ_DTYPE_F32) {
            if (layout == HZ_LAYOUT_NCHW) k_encode<float, false><<<grid, ENC_TPB, 0, st>>>(sp, m, (float*)bp, (float*)gp, tail_vec);
            else k_encode<float, true><<<grid, ENC_TPB, 0, st>>>(sp, m, (float*)bp, (float*)gp, tail_vec);
        } else {
            if (layout == HZ_LAYOUT_NCHW) k_encode<__nv_bfloat16, false><<<grid, ENC_TPB, 0, st>>>(sp, m, (__nv_bfloat16*)bp, (__nv_bfloat16*)gp, tail_vec);
            else k_encode<__nv_bfloat16, true><<<grid, ENC_TPB, 0, st>>>(sp, m, (__nv_bfloat16*)bp, (__nv_bfloat16*)gp, tail_vec);
        }
        launches++;
    }
    return hz_launched(launches);
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

int hz_greedy_actions(const void* states, int64_t n, int16_t* actions, void* stream) {
    if (n == 0) return HZ_OK;
    if (!states || !actions || n < 0) return HZ_ERR_ARG;
    k_greedy<<<blocks_for(n, GWPB), GWPB * 32, 0, (cudaStream_t)stream>>>(states, n, actions);
    return hz_launched(1);
}

int hz_playout(void* states, int64_t n, int max_steps, uint32_t* steps, unsigned long long* total_steps, void* stream) {
    if (n == 0) return HZ_OK;   // empty batch: nothing to do, pointers may be null
    if (!states || n < 0 || max_steps < 0) return HZ_ERR_ARG;
    if (n == 0) return HZ_OK;
    k_playout<false><<<blocks_for(n, PTPB), PTPB, 0, (cudaStream_t)stream>>>(states, n, max_steps, steps, total_steps,
                                                                             nullptr, 0, 0, nullptr);
    return hz_launched(1);
}

int hz_playout_keys(const uint64_t* keys, int64_t n, uint64_t seed, uint64_t first_id, int max_steps,
                    uint32_t* results, unsigned long long* total_steps, void* stream) {
    if (n == 0) return HZ_OK;   // empty batch: nothing to do, pointers may be null
    if (!results || n < 0 || max_steps < 0) return HZ_ERR_ARG;
    k_playout<true><<<blocks_for(n, PTPB), PTPB, 0, (cudaStream_t)stream>>>(nullptr, n, max_steps, nullptr, total_steps,
                                                                            keys, seed, first_id, results);
    return hz_launched(1);
}

}  // extern "C"

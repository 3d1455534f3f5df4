// ovc_rng.cuh — counter-based random start states (host + device), see ovc_random_start_t in
// include/ovc_b200.h.  Philox4x32-10 (Salmon et al., SC'11) with key = the 64-bit run seed and counter
// = (env index low, env index high, episode counter, draw block); the draw plan and the integer
// comparisons are part of the format so a CPU mirror (the test suite has one) can reproduce them bit for bit.
// Semantics: get_random_start_state_fn, reference overcooked_mdp.py:1307-1369.
#pragma once
#include <stdint.h>

#include "../../include/ovc_b200.h"

namespace ovc {

struct Philox4 {
    uint32_t v[4];
};

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint64_t key, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0, c1 = n1, c2 = n2, c3 = n3;
        k0 += 0x9E3779B9u, k1 += 0xBB67AE85u;
    }
    return Philox4{{c0, c1, c2, c3}};
}

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t x, uint32_t n) { return (uint32_t)(((uint64_t)x * n) >> 32); }

// "u < p" on 32-bit draws: draw < threshold, with the saturated threshold 0xFFFFFFFF (p = 1.0) meaning always
__host__ __device__ __forceinline__ bool draw_below(uint32_t draw, uint32_t thr) { return draw < thr || thr == 0xFFFFFFFFu; }

// soup code with n onions followed by m tomatoes and the given tick+1 field
__host__ __device__ __forceinline__ uint32_t soup_code(int n, int m, uint32_t tick_plus_1) {
    return (uint32_t)OVC_O_SOUP | ((uint32_t)(n + m) << 3) | ((((1u << m) - 1u) << n) << 5) | (tick_plus_1 << 8);
}

// The object a player starts with (:1346-1366) or 0: draws = {holds?, which, n, m}
__host__ __device__ __forceinline__ uint32_t random_held(const int32_t *cook_time, uint32_t thr, uint32_t d_has, uint32_t d_obj,
                                                         uint32_t d_n, uint32_t d_m) {
    if (!draw_below(d_has, thr)) return 0;
    if (d_obj < 858993459u) return OVC_O_DISH;    // p = 0.2
    if (d_obj < 3435973836u) return OVC_O_ONION;  // p = 0.6
    const int n = 1 + (int)mulhi32(d_n, 3), m = (int)mulhi32(d_m, (uint32_t)(4 - n));
    return soup_code(n, m, (uint32_t)cook_time[n * 4 + m] + 1u);  // finished soup: tick == cook time (:565-569)
}

// Variable MDP: the layout of the episode that starts now (block 2, word 1 of that episode's draws).
__host__ __device__ __forceinline__ int random_layout_id(const ovc_random_start_t &rs, uint64_t env, uint32_t episode, int n_layouts) {
    const Philox4 Cc = philox4x32_10(rs.seed, (uint32_t)env, (uint32_t)(env >> 32), episode, 2);
    return (int)mulhi32(Cc.v[1], (uint32_t)n_layouts);
}

// Writes one freshly drawn start record (S words).  `start_rec` supplies the fixed start positions,
// `cook_time` / `free_pos` / `n_free` / `n_pots` come from the layout; `episode` is the NEW episode counter.
template <class Store>
__host__ __device__ __forceinline__ void random_start_record(Store &&store, int S, const int32_t *start_rec, const int32_t *cook_time,
                                                             const uint8_t *free_pos, int n_free, int n_pots, int layout_id,
                                                             const ovc_random_start_t &rs, uint64_t env, uint32_t episode) {
    const uint32_t e_lo = (uint32_t)env, e_hi = (uint32_t)(env >> 32);
    const Philox4 A = philox4x32_10(rs.seed, e_lo, e_hi, episode, 0), B = philox4x32_10(rs.seed, e_lo, e_hi, episode, 1),
                  Cc = philox4x32_10(rs.seed, e_lo, e_hi, episode, 2);
    uint32_t pos0 = (uint32_t)start_rec[1] & 0xFF, pos1 = (uint32_t)start_rec[2] & 0xFF;
    if (rs.random_start_pos && n_free >= 2) {  // ordered pairs of distinct floor cells, itertools.product order (:1736-1747)
        const uint32_t idx = mulhi32(A.v[0], (uint32_t)(n_free * (n_free - 1)));
        const uint32_t i = idx / (uint32_t)(n_free - 1);
        uint32_t j = idx % (uint32_t)(n_free - 1);
        j += j >= i;
        pos0 = free_pos[i], pos1 = free_pos[j];
    }
    const uint32_t thr = rs.obj_threshold;
    uint32_t h0 = 0, h1 = 0;
    if (thr) {
        h0 = random_held(cook_time, thr, A.v[1], A.v[2], A.v[3], B.v[0]);
        h1 = random_held(cook_time, thr, B.v[1], B.v[2], B.v[3], Cc.v[0]);
    }
    store(0, 0);
    store(1, (int32_t)(pos0 | (h0 << 10)));  // facing NORTH (orientation index 0), from_player_positions :940-950
    store(2, (int32_t)(pos1 | (h1 << 10)));
    store(3, (int32_t)(((uint32_t)layout_id & 0xFF) | (episode << 16)));
    for (int k = 0; k < S - 4; k++) {
        uint32_t code = 0;
        if (thr && k < n_pots) {  // :1331-1344
            const Philox4 Pk = philox4x32_10(rs.seed, e_lo, e_hi, episode, 3 + (uint32_t)k);
            if (draw_below(Pk.v[0], thr)) {
                const int n = 1 + (int)mulhi32(Pk.v[1], 3), m = (int)mulhi32(Pk.v[2], (uint32_t)(4 - n));
                code = soup_code(n, m, draw_below(Pk.v[3], thr) ? 1u : 0u);
            }
        }
        store(4 + k, (int32_t)code);
    }
}

}  // namespace ovc

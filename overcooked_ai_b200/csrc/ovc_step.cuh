// ovc_step.cuh — device code of one Overcooked joint transition on the packed int32 record.
//
// Restates OvercookedGridworld.get_state_transition (reference
// src/overcooked_ai_py/mdp/overcooked_mdp.py:1375-1430: resolve_interacts :1432-1579, then
// resolve_movement :1644-1727, then step_environment_effects :1691-1703) plus the reward / done
// part of OvercookedEnv.step (overcooked_env.py:244-274), directly on the bit fields of the record
// described in include/ovc_b200.h.  One thread owns one environment.  The record is reached through
// an accessor `R` (shared-memory tile with or without the TMA hardware swizzle, or global memory)
// that provides 16-byte chunk and single-word loads / stores; chunk 0 is the header
// (timestep, player 0, player 1, misc), chunk 1 holds words 4..7 = the pots (and the first counters).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ovc_b200.h"
#include "ovc_rng.cuh"

#include <stddef.h>

namespace ovc {

// Layout-table accessors.  TblG reads the ovc_layout_t records from global memory through the
// read-only path; TblS reads a copy staged in shared memory (the step kernel bulk-copies the table next
// to the record tile, so a launch — which always starts with a cold L1 — pays no serial L2 round trips).
struct TblG {
    const char *base;
    __device__ __forceinline__ TblG at(unsigned id) const { return TblG{base + (size_t)id * sizeof(ovc_layout_t)}; }
    __device__ __forceinline__ int i32(int off) const { return __ldg(reinterpret_cast<const int *>(base + off)); }
    __device__ __forceinline__ unsigned u16(int off) const { return __ldg(reinterpret_cast<const unsigned short *>(base + off)); }
    __device__ __forceinline__ unsigned u8(int off) const { return __ldg(reinterpret_cast<const unsigned char *>(base + off)); }
    __device__ __forceinline__ const char *ptr(int off) const { return base + off; }
};
struct TblS {
    const char *base;  // derived from the __shared__ array, so these compile to LDS
    __device__ __forceinline__ TblS at(unsigned id) const { return TblS{base + id * (uint32_t)sizeof(ovc_layout_t)}; }
    __device__ __forceinline__ int i32(int off) const { return *reinterpret_cast<const int *>(base + off); }
    __device__ __forceinline__ unsigned u16(int off) const { return *reinterpret_cast<const unsigned short *>(base + off); }
    __device__ __forceinline__ unsigned u8(int off) const { return *reinterpret_cast<const unsigned char *>(base + off); }
    __device__ __forceinline__ const char *ptr(int off) const { return base + off; }
};
#define OVC_OFF(field) ((int)offsetof(ovc_layout_t, field))

struct StepOut {
    int sparse;
    int shaped0, shaped1;
    int done;
    unsigned ev0, ev1;
};

__device__ __forceinline__ int comp(const int4 &v, int k) { return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w; }
__device__ __forceinline__ void set_comp(int4 &v, int k, int x) {
    if (k == 0) v.x = x;
    else if (k == 1) v.y = x;
    else if (k == 2) v.z = x;
    else v.w = x;
}

// pos byte is y<<4|x, so a unit move is a byte add: N -16, S +16, E +1, W -1 (actions.py:12-16)
__device__ __forceinline__ int dir_delta(int d) { return (int)(int8_t)((0xFF0110F0u >> (d * 8)) & 0xFF); }

// recipe table row of a soup code: n_onion*4 + n_tomato (ordered kinds bits 5-7, count bits 3-4)
__device__ __forceinline__ int recipe_row(unsigned code) {
    int n = (code >> 3) & 3;
    int nt = __popc((code >> 5) & ((1u << n) - 1u));
    return ((n - nt) << 2) | nt;
}

template <class TB>
__device__ __forceinline__ bool soup_ready(const TB &L, unsigned code) {
    unsigned tp1 = (code >> 8) & 0x3FFF;  // _cooking_tick + 1, 0 = idle
    return (code & 7) == OVC_O_SOUP && tp1 != 0 && (int)(tp1 - 1) >= L.i32(OVC_OFF(cook_time) + 4 * recipe_row(code));  // :537-540
}

// One player's INTERACT (the body of the loop at :1446-1577) on the live record.
template <class R, class TB>
__device__ __forceinline__ void interact_one(R &r, const TB &L, unsigned &p_me,
                                             const unsigned p_other, unsigned &misc, const int n_full,
                                             const int n_dishable, const bool all_full, int &sparse, int &shaped_me,
                                             unsigned &ev_me) {
    const unsigned me = p_me;
    unsigned held = me >> 10;
    const int held_t = held & 7;
    const int other_t = (p_other >> 10) & 7;
    const int fpos = ((int)(me & 0xFF) + dir_delta((me >> 8) & 3)) & 0xFF;
    const unsigned cell = L.u16(OVC_OFF(cell) + 2 * fpos);
    const int terr = cell & 7;
    const int w_idx = 4 + (int)(cell >> 8);
    unsigned cw = 0;
    const bool has_slot = terr == OVC_T_COUNTER || terr == OVC_T_POT;
    if (has_slot) cw = (unsigned)r.ldw(w_idx);
    unsigned e = 0;
    unsigned new_cw = cw;
    if (terr == OVC_T_COUNTER) {
        if (held_t != 0 && cw == 0) {  // :1459-1471 drop on the counter; logged before it happens
            const int base = (0x100C0207u >> ((held_t - 1) * 8)) & 0xFF;  // <obj>_drop of onion,tomato,dish,soup
            e = 1u << base;
            bool useful = false;
            if (held_t <= OVC_O_TOMATO) useful = all_full && other_t != OVC_O_DISH;          // :2239-2254
            else if (held_t == OVC_O_DISH) useful = n_full == 0 && other_t != OVC_O_ONION;  // :2206-2221
            if (useful) e |= 2u << base;
            new_cw = held;
            held = 0;
            if (held_t == OVC_O_DISH) misc += 1u << 8;
        } else if (held_t == 0 && cw != 0) {  // :1473-1485 pick up from the counter
            const int ct = cw & 7;
            const int base = (0x0E0A0005u >> ((ct - 1) * 8)) & 0xFF;  // <obj>_pickup
            e = 1u << base;
            bool useful = false;
            if (ct <= OVC_O_TOMATO) useful = !(all_full && other_t != OVC_O_DISH);  // :2223-2237
            else if (ct == OVC_O_DISH)  // :2180-2204 — this dish still counts as "on a counter"
                useful = ((misc >> 8) & 0xFF) == 0 && (other_t == OVC_O_DISH) < n_dishable;
            if (useful) e |= 2u << base;
            held = cw;
            new_cw = 0;
            if (ct == OVC_O_DISH) misc -= 1u << 8;
        }
    } else if (terr == OVC_T_POT) {
        if (held_t == 0) {  // :1515-1522 start cooking an idle, non-empty soup (new dynamics only)
            if (!(L.i32(OVC_OFF(flags)) & OVC_LAYOUT_OLD_DYNAMICS) && (cw & 7) == OVC_O_SOUP &&
                ((cw >> 8) & 0x3FFF) == 0 && ((cw >> 3) & 3) != 0)
                new_cw = cw | (1u << 8);  // begin_cooking: tick := 0
        } else if (held_t == OVC_O_DISH) {
            if (soup_ready(L, cw)) {  // :1525-1539 plate the soup
                e = 1u << OVC_EV_SOUP_PICKUP;
                held = cw;
                new_cw = 0;
                shaped_me += L.i32(OVC_OFF(rew_soup_pickup));
            }
        } else if (held_t <= OVC_O_TOMATO) {  // :1541-1568 add an ingredient
            unsigned soup = cw ? cw : (unsigned)OVC_O_SOUP;  // empty pot: a fresh soup first (:1544-1546)
            const int n = (soup >> 3) & 3;
            if (((soup >> 8) & 0x3FFF) == 0 && n < 3) {  // not is_full :547-551
                const bool tom = held_t == OVC_O_TOMATO;
                const int old_val = L.i32(OVC_OFF(best_value) + 4 * (n ? recipe_row(soup) : 0));
                soup = (soup & ~(3u << 3)) | ((unsigned)(n + 1) << 3) | ((unsigned)tom << (5 + n));
                const int new_val = L.i32(OVC_OFF(best_value) + 4 * recipe_row(soup));
                // log_object_potting :2121-2140 + is_potting_* :2256-2308 (+ potting_onion again at :1567)
                e = 1u << (tom ? OVC_EV_POTTING_TOMATO : OVC_EV_POTTING_ONION);
                const int sh = tom ? 1 : 0;  // <kind>_tomato_potting = <kind>_onion_potting + 1
                if (old_val == new_val) e |= 1u << (OVC_EV_OPTIMAL_ONION_POTTING + sh);
                if (new_val > 0) e |= 1u << (OVC_EV_VIABLE_ONION_POTTING + sh);
                if (old_val > 0 && new_val == 0) e |= 1u << (OVC_EV_CATASTROPHIC_ONION_POTTING + sh);
                if (old_val == 0) e |= 1u << (OVC_EV_USELESS_ONION_POTTING + sh);
                shaped_me += L.i32(OVC_OFF(rew_placement_in_pot));
                new_cw = soup;
                held = 0;
            }
        }
    } else if (held_t == 0) {
        if (terr == OVC_T_ONION) {  // :1487-1494
            e = 1u << OVC_EV_ONION_PICKUP;
            if (!(all_full && other_t != OVC_O_DISH)) e |= 1u << OVC_EV_USEFUL_ONION_PICKUP;
            held = OVC_O_ONION;
        } else if (terr == OVC_T_TOMATO) {  // :1496-1498 — logs nothing (quirk Q5)
            held = OVC_O_TOMATO;
        } else if (terr == OVC_T_DISH) {  // :1500-1513
            e = 1u << OVC_EV_DISH_PICKUP;
            if (((misc >> 8) & 0xFF) == 0 && (other_t == OVC_O_DISH) < n_dishable) {
                e |= 1u << OVC_EV_USEFUL_DISH_PICKUP;
                shaped_me += L.i32(OVC_OFF(rew_dish_pickup));
            }
            held = OVC_O_DISH;
        }
    } else if (terr == OVC_T_SERVE && held_t == OVC_O_SOUP) {  // :1570-1577, deliver_soup :1631-1642
        const int row = recipe_row(held);
        sparse += L.i32(OVC_OFF(deliver_value) + 4 * row);
        e = (1u << OVC_EV_SOUP_DELIVERY) | ((unsigned)row << OVC_EV_RECIPE_SHIFT);
        held = 0;
    }
    if (has_slot && new_cw != cw) r.stw(w_idx, (int)new_cw);
    p_me = (me & 0x3FFu) | (held << 10);
    ev_me = e;
}

// RS: compiled with the random-start path (get_random_start_state_fn :1307-1369).  It is a separate
// instantiation because the generator's registers would otherwise weigh on every transition.
template <bool RS, class R, class TB>
__device__ __forceinline__ void step_core(R &r, const TB &layouts,
                                          const int32_t *__restrict__ start_records, int S, int a0, int a1,
                                          int horizon, int flags, const ovc_random_start_t *rs, long long env_index,
                                          int n_layouts, StepOut &o) {
    int4 h = r.ld4(0);
    const int t = h.x;
    if (horizon > 0 && t >= horizon) {  // stepping a finished env: untouched + flagged (overcooked_env.py:255)
        o.sparse = 0, o.shaped0 = 0, o.shaped1 = 0, o.done = 1;
        o.ev0 = OVC_EVF_STEPPED_DONE, o.ev1 = OVC_EVF_STEPPED_DONE;
        return;
    }
    unsigned p[2] = {(unsigned)h.y, (unsigned)h.z};
    unsigned misc = (unsigned)h.w;
    const TB L = layouts.at(misc & 0xFF);
    const int n_pots = L.i32(OVC_OFF(n_pots));
    const int lflags = L.i32(OVC_OFF(flags));

    // ---- pot snapshot, taken once before either player acts (get_pot_states :1809-1838 at :1439,
    //      quirk Q3).  Only two aggregates are ever consumed:
    //      n_full     = cooking + ready + idle-with-3      (get_full_pots :1875-1880)
    //      n_dishable = ready + cooking + idle-with-1-or-2 (is_dish_pickup_useful :2199-2203)
    int n_full = 0, n_dishable = 0;
    int4 pw = r.ld4(1);
#pragma unroll 1
    for (int k = 0; k < n_pots; k++) {  // n_pots is uniform across a layout segment: 1 or 2 trips, not 4
        const unsigned w = (unsigned)comp(pw, k);
        const bool soup = (w & 7) == OVC_O_SOUP;
        const int n = (w >> 3) & 3;
        const bool idle = (w & (0x3FFFu << 8)) == 0;
        n_full += soup && (!idle || n == 3);
        n_dishable += soup && (!idle || n == 1 || n == 2);
    }
    const bool all_full = n_full == n_pots;

    int sparse = 0;
    int shaped[2] = {0, 0};
    unsigned ev[2] = {0u, 0u};

    // terrain of each player's move target, fetched early (see resolve_movement below)
    unsigned move_cell[2];
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const int a = i == 0 ? a0 : a1;
        move_cell[i] = L.u16(OVC_OFF(cell) + 2 * ((int)((p[i] & 0xFF) + dir_delta(a & 3)) & 0xFF));
    }

    // ---- resolve_interacts :1446-1577: player 0 then player 1 on the same live record.  The body is
    //      emitted once and run as a loop of up to two trips: trip 0 handles, per environment, the first
    //      interacting player (player 0 if it interacts, else player 1); trip 1 handles player 1 where
    //      BOTH interact (1 environment in 36 under a uniform policy), so most warps skip it. ----
    {
        const bool i0 = a0 == OVC_A_INTERACT, i1 = a1 == OVC_A_INTERACT;
#pragma unroll 1
        for (int trip = 0; trip < 2; trip++) {
            const bool now = trip == 0 ? (i0 || i1) : (i0 && i1);
            if (now) {
                const bool second = trip == 1 || !i0;  // the acting player is player 1
                unsigned pa = second ? p[1] : p[0];
                const unsigned pb = second ? p[0] : p[1];
                int sa = 0;
                unsigned ea = 0u;
                interact_one(r, L, pa, pb, misc, n_full, n_dishable, all_full, sparse, sa, ea);
                if (second) p[1] = pa, shaped[1] = sa, ev[1] = ea;
                else p[0] = pa, shaped[0] = sa, ev[0] = ea;
            }
        }
    }

    // ---- resolve_movement :1644-1727 from the pre-step positions; a blocked or collided player
    //      still turns (quirk Q8).  (The two terrain lookups were issued before the interact section:
    //      positions do not change there, and the loads then overlap the interact latency.) ----
    int npos[2];
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const int pos = p[i] & 0xFF;
        npos[i] = pos;
        const int a = i == 0 ? a0 : a1;
        if (a < 4) {
            if ((move_cell[i] & 7) == OVC_T_FLOOR) npos[i] = (pos + dir_delta(a)) & 0xFF;
            p[i] = (p[i] & ~0x300u) | ((unsigned)a << 8);
        }
    }
    {
        const int o0 = p[0] & 0xFF, o1 = p[1] & 0xFF;
        const bool collide = npos[0] == npos[1] || (npos[0] == o1 && npos[1] == o0);  // :1673-1683
        if (!collide) {
            p[0] = (p[0] & ~0xFFu) | (unsigned)npos[0];
            p[1] = (p[1] & ~0xFFu) | (unsigned)npos[1];
        }
    }

    // ---- step_environment_effects :1691-1703 (soups outside pots are always finished: no tick) ----
    const int tn = t + 1;
    {
        pw = r.ld4(1);
        bool changed = false;
        const bool old_dyn = lflags & OVC_LAYOUT_OLD_DYNAMICS;
#pragma unroll 1
        for (int k = 0; k < n_pots; k++) {
            const unsigned w = (unsigned)comp(pw, k);
            if ((w & 7) == OVC_O_SOUP) {
                unsigned tp1 = (w >> 8) & 0x3FFF;
                if (old_dyn && tp1 == 0 && ((w >> 3) & 3) == 3) tp1 = 1;  // auto start :1696-1701
                if (tp1 != 0 && (int)(tp1 - 1) < L.i32(OVC_OFF(cook_time) + 4 * recipe_row(w))) tp1 += 1;  // cook :601-606
                const unsigned nw = (w & 0xFFu) | (tp1 << 8);
                changed |= nw != w;
                set_comp(pw, k, (int)nw);
            }
        }
        if (changed) r.st4(1, pw);
    }

    o.sparse = sparse;  // OvercookedEnv.step returns the sum over agents (overcooked_env.py:273)
    o.shaped0 = shaped[0], o.shaped1 = shaped[1];
    o.ev0 = ev[0], o.ev1 = ev[1];
    o.done = horizon > 0 && tn >= horizon;  // is_done overcooked_env.py:321-325
    if (o.done && (flags & OVC_F_AUTO_RESET)) {
        const int32_t *__restrict__ start = start_records + (size_t)(misc & 0xFF) * S;
        if (RS && rs) {
            const unsigned episode = ((misc >> 16) + 1u) & 0xFFFFu;
            int lid = (int)(misc & 0xFF);
            if (rs->random_layout) lid = random_layout_id(*rs, (uint64_t)env_index, episode, n_layouts);  // variable MDP
            const TB Ln = layouts.at((unsigned)lid);
            random_start_record([&](int w, int32_t v) { r.stw(w, v); }, S, start_records + (size_t)lid * S,
                                reinterpret_cast<const int32_t *>(Ln.ptr(OVC_OFF(cook_time))),
                                reinterpret_cast<const uint8_t *>(Ln.ptr(OVC_OFF(free_pos))), Ln.i32(OVC_OFF(n_free)),
                                Ln.i32(OVC_OFF(n_pots)), lid, *rs, (uint64_t)env_index, episode);
        } else {
            const int4 *__restrict__ src = reinterpret_cast<const int4 *>(start);
#pragma unroll 4
            for (int c = 0; c < S / 4; c++) r.st4(c, __ldg(src + c));
        }
    } else {
        r.st4(0, make_int4(tn, (int)p[0], (int)p[1], (int)misc));
    }
}

}  // namespace ovc

// ovc_potential.cuh — K6 potential_kernel: the reference's shaped-reward potential phi(s)
// (overcooked_mdp.py:2920-3250), one thread per environment, double precision.
//
// The reference evaluates gamma ** (small integer) in Python floats and combines the factors with
// plain double multiplications and additions in a fixed order.  The kernel takes the powers from a
// host-built table (gpow[k] = gamma ** k as Python computes it) and performs every multiplication and
// addition with __dmul_rn / __dadd_rn in exactly that order (no FMA contraction), so phi is
// reproduced bit for bit, not merely within a tolerance.  Player "lists" are 2-bit masks walked in
// player order; soups are visited in the orders the reference's Python containers would yield
// (pot order, dict insertion order, CPython's set order for the partially full pots — from the table).
#pragma once

namespace ovc {

struct PotArgs {
    const ovc_layout_t *layouts;
    const ovc_potential_t *pt;
    const ovc_cost_lut_entry_t *cost;
    const double *gpow;
    const int32_t *state;
    double *out;
    long long n_envs;
    int S, n_pow;
};

#define OVC_BIG 1000000

__device__ __forceinline__ double gpw(const PotArgs &a, int k) { return __ldg(a.gpow + (k < a.n_pow ? k : a.n_pow - 1)); }

__global__ void __launch_bounds__(128) potential_kernel(const PotArgs a) {
    const long long env = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= a.n_envs) return;
    const int32_t *__restrict__ rec = a.state + env * a.S;
    const int4 h = __ldg(reinterpret_cast<const int4 *>(rec));
    const int lid = h.w & 0xFF;
    const ovc_layout_t *__restrict__ L = a.layouts + lid;
    const ovc_potential_t *__restrict__ P = a.pt + lid;
    const ovc_cost_lut_entry_t *__restrict__ C = a.cost + (size_t)lid * 1024;
    const int n_pots = __ldg(&L->n_pots);
    const int max_del = __ldg(&P->max_delivery_steps), max_pick = __ldg(&P->max_pickup_steps);
    const unsigned pl[2] = {(unsigned)h.y, (unsigned)h.z};

    // planner costs of both players: one 8-byte entry each (serve, pot[0..3])
    unsigned long long ce[2];
#pragma unroll
    for (int i = 0; i < 2; i++)
        ce[i] = __ldg(reinterpret_cast<const unsigned long long *>(C + (((pl[i] & 0xFF) << 2) | ((pl[i] >> 8) & 3))));
    auto cost_of = [&](int player, int which) {  // which 0 = serve, 1 + k = pot k; OVC_BIG if unreachable
        const int c = (int)((ce[player] >> (8 * which)) & 0xFF);
        return c == OVC_COST_INF ? OVC_BIG : c;
    };

    // pot words and their classes (get_pot_states :1809-1838)
    unsigned pw[OVC_MAX_POTS];
    int cls[OVC_MAX_POTS];  // 0 empty, 1..3 idle with that many items, 4 cooking, 5 ready
    int remaining[OVC_MAX_POTS], row[OVC_MAX_POTS];
#pragma unroll
    for (int k = 0; k < OVC_MAX_POTS; k++) {
        pw[k] = k < n_pots ? (unsigned)__ldg(rec + 4 + k) : 0u;
        cls[k] = 0, remaining[k] = 0, row[k] = 0;
        if ((pw[k] & 7) == OVC_O_SOUP) {
            const int n = (pw[k] >> 3) & 3;
            const int tp1 = (pw[k] >> 8) & 0x3FFF;
            row[k] = recipe_row(pw[k]);
            if (tp1 == 0) cls[k] = n;  // n == 0 cannot persist
            else {
                remaining[k] = __ldg(&L->cook_time[row[k]]) - (tp1 - 1);
                cls[k] = remaining[k] <= 0 ? 5 : 4;
            }
        }
    }

    double potential = __ldg(&P->steady);  // :2985-2999

    // player masks by held object (:3047-3070)
    unsigned m_soup = 0, m_dish = 0, m_tom = 0, m_oni = 0, m_none = 0;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const int t = (pl[i] >> 10) & 7;
        m_soup |= (unsigned)(t == OVC_O_SOUP) << i, m_dish |= (unsigned)(t == OVC_O_DISH) << i;
        m_tom |= (unsigned)(t == OVC_O_TOMATO) << i, m_oni |= (unsigned)(t == OVC_O_ONION) << i;
        m_none |= (unsigned)(t == OVC_O_NONE) << i;
    }

    // ---- step 4: players holding a soup (:3075-3084) ----
#pragma unroll
    for (int i = 0; i < 2; i++)
        if ((m_soup >> i) & 1) {
            const int v = max(__ldg(&L->deliver_value[recipe_row(pl[i] >> 10)]), 1);
            potential = __dadd_rn(potential, __dmul_rn(gpw(a, min(cost_of(i, 0), max_del)), (double)v));
        }

    // ---- step 3: non-idle soups, cooking ones first then ready ones, each in pot order (:3026-3043) ----
    int non[OVC_MAX_POTS], n_non = 0;
    double val[OVC_MAX_POTS];
    for (int pass = 4; pass <= 5; pass++)
        for (int k = 0; k < n_pots; k++)
            if (cls[k] == pass) {
                const int v = max(__ldg(&L->deliver_value[row[k]]), 1);
                val[n_non] = __dmul_rn(gpw(a, max_del + max(max_pick, remaining[k])), (double)v);
                non[n_non++] = k;
            }
#pragma unroll
    for (int i = 0; i < 2; i++)
        if ((m_dish >> i) & 1) {  // :3089-3128
            int best = -1;
            double best_value = 0.0;
            for (int j = 0; j < n_non; j++) {
                const int k = non[j];
                const int pd = cost_of(i, 1 + k);
                if (pd >= OVC_BIG) continue;  // is_useful == 0: value 0, never selected
                const int v = max(__ldg(&L->deliver_value[row[k]]), 1);
                const double psv = __dmul_rn(gpw(a, max_del), (double)v);
                const double disc = gpw(a, max(remaining[k], min(pd, max_pick)));
                const double pv = __dmul_rn(__dmul_rn(disc, psv), 1.0);
                if (pv > best_value) best = j, best_value = pv;
            }
            if (best >= 0 && best_value > val[best]) val[best] = best_value;
        }
    for (int j = 0; j < n_non; j++) potential = __dadd_rn(potential, val[j]);

    // ---- step 2: idle soups (:3002-3022, :3137-3210) ----
    int idle[OVC_MAX_POTS], n_idle = 0, code = 0, p3 = 1;
    for (int k = 0; k < n_pots; k++) {
        if (cls[k] == 3) idle[n_idle++] = k;
        code += (cls[k] == 1 ? 1 : cls[k] == 2 ? 2 : 0) * p3;
        p3 *= 3;
    }
    {
        const unsigned ord = __ldg(reinterpret_cast<const unsigned *>(&P->partial_order[code][0]));
        for (int j = 0; j < 4; j++) {
            const int s = (ord >> (8 * j)) & 0xFF;
            if (s == OVC_NO_SLOT) break;
            idle[n_idle++] = s;
        }
    }
    for (int i = 1; i < n_idle; i++)  // stable, descending by the discounted value of the best reachable recipe
        for (int j = i; j > 0 && __ldg(&P->disc_value[row[idle[j - 1]]]) < __ldg(&P->disc_value[row[idle[j]]]); j--) {
            const int t = idle[j];
            idle[j] = idle[j - 1], idle[j - 1] = t;
        }
    for (int q = 0; q < n_idle; q++) {
        const int k = idle[q];
        const int cur = row[k], opt = __ldg(&P->opt_recipe[cur]);
        const int miss_on = (opt >> 2) - (cur >> 2), miss_to = (opt & 3) - (cur & 3);
        double disc = gpw(a, max(max_pick, __ldg(&L->cook_time[opt])) + max_del);
        for (int m = 0; m < miss_on + miss_to; m++) {  // sorted ingredient tuple: onions, then tomatoes
            const bool tom = m >= miss_on;
            unsigned &mask = tom ? m_tom : m_oni;
            int dist = OVC_BIG, who = -1;
#pragma unroll
            for (int i = 0; i < 2; i++)
                if ((mask >> i) & 1) {
                    const int cd = cost_of(i, 1 + k);
                    if (cd < dist) dist = cd, who = i;
                }
            disc = __dmul_rn(disc, gpw(a, min(dist, tom ? __ldg(&P->pot_tomato_steps) : __ldg(&P->pot_onion_steps))));
            if (who >= 0) mask &= ~(1u << who);  // that player's ingredient is spoken for
        }
        if (miss_on + miss_to > 0) {
            disc = __dmul_rn(disc, gpw(a, 1));
        } else {
            int cook_dist = OVC_BIG;
#pragma unroll
            for (int i = 0; i < 2; i++)
                if ((m_none >> i) & 1) cook_dist = min(cook_dist, cost_of(i, 1 + k));
            disc = __dmul_rn(disc, gpw(a, min(cook_dist, max_pick)));
        }
        potential = __dadd_rn(potential, __dmul_rn(disc, (double)max(__ldg(&L->deliver_value[opt]), 1)));
    }

    // ---- step 1: left-over ingredients and the closest EMPTY pot (:3215-3247) ----
    for (int pass = 0; pass < 2; pass++) {
        const unsigned mask = pass == 0 ? m_tom : m_oni;
#pragma unroll
        for (int i = 0; i < 2; i++)
            if ((mask >> i) & 1) {
                int dist = OVC_BIG;
                for (int k = 0; k < n_pots; k++)
                    if (cls[k] == 0) dist = min(dist, cost_of(i, 1 + k));
                const int steps = pass == 0 ? __ldg(&P->pot_tomato_steps) : __ldg(&P->pot_onion_steps);
                const double useful = dist < OVC_BIG ? 1.0 : 0.0;
                const double disc = __dmul_rn(gpw(a, min(steps, dist) + max_pick + max_del), useful);
                const int value = pass == 0 ? __ldg(&P->tomato_value) : __ldg(&P->onion_value);
                potential = __dadd_rn(potential, __dmul_rn(disc, (double)value));
            }
    }
    a.out[env] = potential;
}

static int potential_impl(const ovc_layout_t *layouts, const ovc_potential_t *pt, const ovc_cost_lut_entry_t *cost,
                          const double *gpow, int n_pow, const int32_t *state, double *out, long long n_envs, int S,
                          cudaStream_t st) {
    if (!pt || !cost || !gpow || !out) return fail(OVC_E_BADARG, "null pointer argument");
    if (n_pow < 2) return fail(OVC_E_BADARG, "gamma power table too short");
    if (n_envs == 0) return OVC_OK;
    PotArgs a{layouts, pt, cost, gpow, state, out, n_envs, S, n_pow};
    potential_kernel<<<(unsigned)((n_envs + 127) / 128), 128, 0, st>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "potential kernel launch");
    return OVC_OK;
}

}  // namespace ovc

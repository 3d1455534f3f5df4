"""CPU tests: the C oracle (oracle/ovc_oracle.c) against fixtures generated from the reference.

These pin the oracle.  The fixtures under tests/golden/ were produced by tools/make_golden.py,
which runs the unmodified reference (and asserts the reference's own golden vectors on the way).
"""
import json

import numpy as np
import pytest

from helpers import EVENT_MASK, GOLD, TRACE_FILES, TRACE_IDS, Trace, lut_bytes
from oracle import cpu
from overcooked_ai_b200 import layout as L
from overcooked_ai_b200.state import OvercookedState


def _check_transitions(tr):
    s0, a, s1, sparse, shaped, events = tr.flat()
    st = s0.copy()
    o_sp, o_sh, o_dn, o_ev = cpu.step(tr.tables, tr.starts, st, a, horizon=0)
    assert np.array_equal(st, s1)
    assert np.array_equal(o_sp, sparse)
    assert np.array_equal(o_sh, shaped)
    assert np.array_equal(o_ev & EVENT_MASK, events)
    # engine extra: recipe index of a delivered soup sits in bits 25-28 and maps to the reward
    deliv = (o_ev >> 15) & 1
    rec = (o_ev >> L.EV_RECIPE_SHIFT) & 15
    assert np.array_equal(rec > 0, deliv == 1)
    per_agent = tr.layout.deliver_value[rec] * deliv
    assert np.array_equal(per_agent.reshape(tr.sparse2.shape), tr.sparse2)


def test_reference_golden_trajectory_mdp_test():
    """testing/overcooked_test.py:516-525 (test_mdp_dynamics) — the reference's own 1500-step vector."""
    tr = Trace(GOLD + "/dynamics_mdp_test.npz")
    assert tr.T == 1500 and tr.sparse.sum() == 10
    _check_transitions(tr)


@pytest.mark.parametrize("path", TRACE_FILES, ids=TRACE_IDS)
def test_transitions(path):
    _check_transitions(Trace(path))


def test_greedy_rollouts_cramped_room():
    """5 seeded GreedyHumanModel games (180 sparse reward each): the soup pickup / delivery branches."""
    tr = Trace(GOLD + "/greedy_cramped_room.npz")
    assert tr.sparse.sum(1).tolist() == [180] * 5
    # the fixture stores s_0..s_{T-1}; check the T-1 transitions between stored states
    S = tr.S
    s0 = np.ascontiguousarray(tr.states[:, :-1].reshape(-1, S)).copy()
    a = np.ascontiguousarray(tr.actions[:, :-1].reshape(-1, 2))
    o_sp, o_sh, o_dn, o_ev = cpu.step(tr.tables, tr.starts, s0, a, horizon=400)
    assert np.array_equal(s0, tr.states[:, 1:].reshape(-1, S))
    assert np.array_equal(o_sp, tr.sparse[:, :-1].reshape(-1))
    assert np.array_equal(o_sh, tr.shaped[:, :-1].reshape(-1, 2))
    assert np.array_equal(o_ev & EVENT_MASK, tr.events[:, :-1].reshape(-1, 2))
    assert not o_dn.any()


def test_rollout_equals_repeated_step_and_done():
    tr = Trace(GOLD + "/greedy_cramped_room.npz")
    st = np.ascontiguousarray(tr.states[:, 0]).copy()
    acts = np.ascontiguousarray(tr.actions.transpose(1, 0, 2))  # [T,E,2]
    sp, sh, dn, ev = cpu.rollout(tr.tables, tr.starts, st, acts, horizon=400, flags=0, n_threads=2)
    assert np.array_equal(sp.T, tr.sparse) and np.array_equal(sh.transpose(1, 0, 2), tr.shaped)
    assert dn[-1].all() and not dn[:-1].any() and (st[:, 0] == 400).all()
    # stepping a finished env: untouched + flagged (the reference asserts, overcooked_env.py:255)
    before = st.copy()
    sp2, sh2, dn2, ev2 = cpu.step(tr.tables, tr.starts, st, acts[0], horizon=400)
    assert np.array_equal(before, st) and dn2.all() and (ev2 == L.EVF_STEPPED_DONE).all()
    # auto reset: the record becomes the layout's start record when the horizon is reached
    st = np.ascontiguousarray(tr.states[:, 0]).copy()
    sp, sh, dn, ev = cpu.rollout(tr.tables, tr.starts, st, acts, horizon=400, flags=1, n_threads=1)
    assert dn[-1].all() and np.array_equal(st, np.repeat(tr.starts, st.shape[0], 0))


def test_lossless_reference_golden_pickle():
    """testing/overcooked_test.py:1050-1067: (5,400,2,5,4,26) expected.pickle, bit for bit."""
    d = np.load(GOLD + "/greedy_cramped_room.npz")
    cl = L.compile_layout("cramped_room")
    tab, starts, S = L.build_tables([cl])
    enc = cpu.encode_lossless(tab, d["states"].reshape(-1, S), 5, 4, horizon=400)
    assert np.array_equal(enc.reshape(5, 400, 2, 5, 4, 26), d["lossless"].astype(np.int32))


@pytest.mark.parametrize("num_pots", [0, 1, 2])
def test_featurize_reference_golden_pickles(num_pots):
    """testing/overcooked_test.py:1069-1093: expected_{0,1,2}.pickle."""
    d = np.load(GOLD + "/greedy_cramped_room.npz")
    cl = L.compile_layout("cramped_room")
    tab, starts, S = L.build_tables([cl])
    f = cpu.featurize(tab, lut_bytes([cl]), d["states"].reshape(-1, S), num_pots)
    exp = d["feat_%d" % num_pots]
    assert np.array_equal(f.reshape(exp.shape), exp.astype(np.float64))


@pytest.mark.parametrize("path", TRACE_FILES, ids=TRACE_IDS)
def test_observations(path):
    tr = Trace(path)
    d = tr.data
    cl = tr.layout
    enc = cpu.encode_lossless(tr.tables, d["obs_states"], cl.width, cl.height, horizon=400)
    assert np.array_equal(enc, d["obs_lossless"].astype(np.int32))
    lut = lut_bytes([cl])
    for num_pots in (0, 1, 2, 3):
        f = cpu.featurize(tr.tables, lut, d["obs_states"], num_pots)
        assert np.array_equal(f, d["obs_feat_%d" % num_pots].astype(np.float64)), num_pots


def test_feature_lut_matches_reference_planner():
    """CompiledLayout.feature_lut() (own BFS) == argmins of the reference MotionPlanner
    (planning/planners.py:391-423) on every bundled 2-player layout the fixture covers."""
    d = np.load(GOLD + "/planner_luts.npz")
    assert len(d.files) >= 40
    for name in d.files:
        cl = L.compile_layout(name)
        mine = cl.feature_lut().view(np.uint8).reshape(-1)
        assert np.array_equal(mine, d[name]), name


@pytest.mark.parametrize("path", TRACE_FILES + [GOLD + "/dynamics_mdp_test.npz"])
def test_pack_unpack_roundtrip_against_reference_dicts(path):
    """unpack(pack(state)).to_dict() == the reference's own to_dict() output (wire format parity)."""
    tr = Trace(path)
    sample = json.loads(str(tr.data["to_dict_sample"]))
    assert sample
    for key, ref_dict in sample.items():
        st = OvercookedState.from_dict(ref_dict)
        rec = L.pack_state(tr.layout, st, 0, tr.S)
        back = L.unpack_state(tr.layout, rec)
        assert back == st
        # the reference lists objects in dict-insertion (history) order; the record is canonical
        got = json.loads(json.dumps(back.to_dict()))
        got["objects"].sort(key=lambda o: o["position"])
        ref_dict["objects"].sort(key=lambda o: o["position"])
        assert got == ref_dict
    # and every stored record survives unpack -> pack unchanged
    flat = tr.states.reshape(-1, tr.S)[:: max(1, tr.states.size // tr.S // 300)]
    for rec in flat:
        assert np.array_equal(L.pack_state(tr.layout, L.unpack_state(tr.layout, rec), 0, tr.S), rec)


def _potential_cases():
    g = np.load(GOLD + "/potential.npz")
    for path in TRACE_FILES:
        name = path.split("trace_")[-1][:-4]
        if name + "__phi" in g.files:
            yield path, name, g


@pytest.mark.parametrize("gamma_idx,gamma", [(0, 0.99), (1, 0.9)])
def test_potential_function_bit_exact(gamma_idx, gamma):
    """potential_function (overcooked_mdp.py:2920-3250): the C oracle reproduces the reference's Python floats
    exactly (==, not allclose) on every fixture state; the planner costs it uses equal the reference
    MotionPlanner's min_cost_to_feature."""
    n_states = 0
    for path, name, g in _potential_cases():
        tr = Trace(path)
        cl = tr.layout
        cost = cl.cost_lut()
        assert np.array_equal(cost["serve"], g[name + "__cost"][..., 0]), name
        assert np.array_equal(cost["pot"][..., :cl.n_pots], g[name + "__cost"][..., 1:1 + cl.n_pots]), name
        pt, cst, gpow = L.build_potential_tables([cl], gamma)
        phi = cpu.potential(tr.tables, pt, cst, gpow, tr.data["obs_states"])
        assert np.array_equal(phi, g[name + "__phi"][:, gamma_idx]), name
        n_states += len(phi)
    assert n_states > 10000


def test_random_start_states_are_valid_and_follow_the_reference_distribution():
    """The engine's counter-based get_random_start_state_fn (overcooked_mdp.py:1307-1369): every drawn state
    passes the reference's validity rules (pack_state re-checks them), and the marginals match what the
    reference's procedure implies: P(player holds something) = t with dish / onion / soup split 0.2 / 0.6 / 0.2,
    P(pot filled) = t, P(filled pot already cooking) = t, ingredient counts n uniform on 1..3 and m on 0..3-n,
    joint positions uniform over ordered pairs of distinct floor cells."""
    t = 0.6
    cl = L.compile_layout("counter_circuit")
    tab, starts, S = L.build_tables([cl])
    n = 200000
    state = np.zeros((n, S), np.int32)
    cpu.reset_random(tab, starts, state, cpu.random_start(5, t, True))
    held = (state[:, 1:3].astype(np.uint32) >> 10) & 7
    assert abs((held != 0).mean() - t) < 0.01
    for typ, p in ((L.O_DISH, 0.2), (L.O_ONION, 0.6), (L.O_SOUP, 0.2)):
        assert abs((held == typ).sum() / (held != 0).sum() - p) < 0.01
    pots = state[:, 4:4 + cl.n_pots].astype(np.uint32)
    filled = (pots & 7) == L.O_SOUP
    assert abs(filled.mean() - t) < 0.01
    assert abs((((pots >> 8) & 0x3FFF) == 1)[filled].mean() - t) < 0.01
    n_ing = ((pots >> 3) & 3)[filled]
    n_tom = np.array([bin(int(x)).count("1") for x in ((pots >> 5) & 7)[filled][:20000]])
    n_on = n_ing[:20000] - n_tom
    for k in (1, 2, 3):
        assert abs((n_on == k).mean() - 1 / 3) < 0.02
    pos = state[:, 1:3] & 0xFF
    assert (pos[:, 0] != pos[:, 1]).all()
    F = len(cl.terrain_pos_dict[" "])
    pairs, counts = np.unique(pos, axis=0, return_counts=True)
    assert len(pairs) == F * (F - 1) and counts.min() > 0.8 * n / (F * (F - 1))
    assert not state[:, 4 + cl.n_pots:].any() and (state[:, 0] == 0).all()
    # validity: every 400th state round-trips through pack_state (which asserts the reference's rules)
    for rec in state[::400]:
        st = L.unpack_state(cl, rec)
        back = L.pack_state(cl, st, 0, S)
        assert np.array_equal(back[:3], rec[:3]) and np.array_equal(back[4:], rec[4:])
        for p in st.players:
            if p.held_object is not None and p.held_object.name == "soup":
                assert p.held_object.is_ready
    # no randomisation requested -> the standard start, like the reference (:1316-1326)
    state2 = np.zeros((8, S), np.int32)
    cpu.reset_random(tab, starts, state2, cpu.random_start(5, 0.0, False))
    assert np.array_equal(state2[:, :3], np.repeat(starts[:, :3], 8, 0))


def _philox4x32_10(key, c0, c1, c2, c3):
    k0, k1 = key & 0xFFFFFFFF, (key >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c0, 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c3 ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF
        k0, k1 = (k0 + 0x9E3779B9) & 0xFFFFFFFF, (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return c0, c1, c2, c3


def test_baseline_config_1_trace_and_episode_info():
    """BASELINE config 1: the fixture's actions are the documented counter-based stream (Philox4x32-10, key = seed,
    counter (env id, t), word = agent, scaled by multiply-high), and the oracle's outputs over the 400 transitions
    add up to the episode info the reference's OvercookedEnv reports (overcooked_env.py:363-401)."""
    tr = Trace(GOLD + "/trace_config1_cramped_room.npz")
    infos = json.loads(str(tr.data["episode_info"]))
    assert tr.E == 8 and tr.T == 400 and (tr.states[:, 0] == tr.starts[0]).all() and (tr.states[:, -1, 0] == 400).all()
    for seed in range(8):
        for t in (0, 1, 57, 399):
            v = _philox4x32_10(seed, 0, t, 0, 0)
            assert tr.actions[seed, t].tolist() == [(v[0] * 6) >> 32, (v[1] * 6) >> 32]
    st = np.ascontiguousarray(tr.states[:, 0])
    acts = np.ascontiguousarray(tr.actions.transpose(1, 0, 2))
    sparse, shaped, done, events = cpu.rollout(tr.tables, tr.starts, st, acts, horizon=400, flags=0)
    assert np.array_equal(st, tr.states[:, -1]) and done[-1].all() and not done[:-1].any()
    val = tr.layout.deliver_value
    for seed, info in enumerate(infos):
        ev = events[:, seed]
        by_agent = (val[(ev >> 25) & 15] * ((ev >> 15) & 1)).sum(0)
        assert by_agent.tolist() == info["ep_sparse_r_by_agent"] and int(sparse[:, seed].sum()) == info["ep_sparse_r"]
        assert shaped[:, seed].sum(0).tolist() == info["ep_shaped_r_by_agent"] and info["ep_length"] == 400
        for i, name in enumerate(L.EVENT_TYPES):
            for agent in range(2):
                assert np.nonzero((ev[:, agent] >> i) & 1)[0].tolist() == info["game_stats"][name][agent], (seed, name)

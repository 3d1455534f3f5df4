"""Variable-MDP support on the host: the procedural layout generator against the reference's outputs
(tests/golden/layout_generator.npz, written by tools/make_golden.py from the live reference under
np.random.seed), the drop-in OvercookedEnv over a generator function, and the CPU mirror of the
engine's documented layout redraw (ovc_random_start_t.random_layout).  No GPU needed."""
import copy
import json
import os

import numpy as np
import pytest

from oracle import cpu
from overcooked_ai_b200 import layout as L
from overcooked_ai_b200 import layout_generator as LG
from overcooked_ai_b200.env import OvercookedEnv

GOLD = os.path.join(os.path.dirname(__file__), "golden", "layout_generator.npz")


def _cases():
    g = np.load(GOLD)
    names = sorted(k[: -len("__case")] for k in g.files if k.endswith("__case"))
    return g, names


@pytest.mark.parametrize("name", _cases()[1])
def test_generator_reproduces_the_reference_layouts(name):
    """Same numpy seed -> same terrain and start cells as the reference's LayoutGenerator (layout_generator.py:144-405)."""
    g, _ = _cases()
    case = json.loads(str(g[name + "__case"]))
    for i, k in enumerate(case["seeds"]):
        np.random.seed(k)
        gen = LG.LayoutGenerator(LG.MDPParamsGenerator.from_fixed_param(copy.deepcopy(case["params"])),
                                 outer_shape=tuple(case["outer_shape"]))
        m = gen.generate_padded_mdp()
        assert ["".join(r) for r in m.terrain_mtx] == list(g[name + "__terrain"][i]), (name, k)
        assert [tuple(p) for p in m.start_player_positions] == [tuple(int(v) for v in p) for p in g[name + "__starts"][i]]


def _floor_connected(rows):
    floor = {(x, y) for y, r in enumerate(rows) for x, c in enumerate(r) if c == " "}
    seen, todo = set(), [next(iter(floor))]
    while todo:
        c = todo.pop()
        if c in seen:
            continue
        seen.add(c)
        todo += [n for n in ((c[0] + 1, c[1]), (c[0] - 1, c[1]), (c[0], c[1] + 1), (c[0], c[1] - 1)) if n in floor]
    return seen == floor


def test_generated_layouts_have_every_feature_and_one_connected_floor():
    """What the reference's own tests assert (overcooked_test.py:1359-1395) plus the dig invariant (:341-357)."""
    np.random.seed(5)
    for feats in (["P", "D", "S", "O", "T"], ["P", "D", "S", "O"], ["P", "D", "S", "T"]):
        params = {"prop_feats": 0.5, "feature_types": feats, "prop_empty": 0.5, "inner_shape": (6, 5), "display": False,
                  "start_all_orders": [{"ingredients": ["onion", "onion", "onion"]}]}
        fn = LG.LayoutGenerator.mdp_gen_fn_from_dict(params, outer_shape=(6, 5))
        for _ in range(10):
            m = fn({})
            rows = ["".join(r) for r in m.terrain_mtx]
            chars = set("".join(rows))
            assert all(f in chars for f in feats) and not (set("OT") - set(feats)) & chars
            assert _floor_connected(rows)
            p0, p1 = m.start_player_positions
            assert p0 != p1 and rows[p0[1]][p0[0]] == " " and rows[p1[1]][p1[0]] == " "
            assert m.start_all_orders == [{"ingredients": ["onion", "onion", "onion"]}]


def test_generated_orders():
    """generate_all_orders / generate_bonus_orders (layout_generator.py:217-254, overcooked_test.py:1397-1483)."""
    np.random.seed(1)
    only_onions = [{"ingredients": ("onion", "onion")}, {"ingredients": ("onion", "onion", "onion")}]
    params = {"generate_all_orders": {"n": 2, "ingredients": ["onion"], "min_size": 2, "max_size": 3},
              "generate_bonus_orders": {"n": 1, "min_size": 2, "max_size": 3},
              "prop_feats": 0.9, "prop_empty": 0.1, "inner_shape": (6, 5), "display": False}
    fn = LG.LayoutGenerator.mdp_gen_fn_from_dict(params, outer_shape=(6, 5))
    key = lambda o: tuple(o["ingredients"])
    seen = set()
    for _ in range(10):
        try:
            m = fn({})
        except ValueError:  # a draw with more pots than the record format holds
            continue
        assert sorted(map(key, m.start_all_orders)) == sorted(map(key, only_onions))
        assert len(m.start_bonus_orders) == 1 and key(m.start_bonus_orders[0]) in set(map(key, only_onions))
        seen.add(key(m.start_bonus_orders[0]))
    assert len(seen) == 2
    with pytest.raises(AssertionError):
        LG.LayoutGenerator.mdp_gen_fn_from_dict({"inner_shape": (5, 4), "prop_empty": 0.9, "prop_feats": 0.1, "display": False},
                                                outer_shape=(5, 4))({})  # no orders at all
    with pytest.raises(TypeError):
        LG.LayoutGenerator.mdp_gen_fn_from_dict(**{"None": None})  # overcooked_test.py:1311-1316


def test_dropin_env_draws_a_new_mdp_at_every_reset():
    """OvercookedEnv(mdp_generator_fn): reset(regen_mdp=True) calls the generator (overcooked_env.py:299-302)."""
    np.random.seed(0)
    params = dict(LG.DEFAULT_MDP_GEN_PARAMS, prop_empty=0.8, prop_feats=0.2)
    env = OvercookedEnv(LG.LayoutGenerator.mdp_gen_fn_from_dict(params, outer_shape=(5, 4)), horizon=400)
    seen = {tuple("".join(r) for r in env.mdp.terrain_mtx)}
    for _ in range(6):
        env.reset()
        seen.add(tuple("".join(r) for r in env.mdp.terrain_mtx))
        assert env.state.timestep == 0 and env.state.player_positions == tuple(env.mdp.start_player_positions)
    assert len(seen) > 3
    fixed = OvercookedEnv(LG.LayoutGenerator.mdp_gen_fn_from_dict({"layout_name": "cramped_room"}), horizon=400)
    first = fixed.mdp.terrain_mtx
    fixed.reset()
    assert fixed.mdp.terrain_mtx == first


def test_layout_redraw_mirror_is_uniform_deterministic_and_per_episode():
    """ovc_random_start_t.random_layout through the CPU mirror: the id is a function of (seed, env, episode),
    uniform over the pool; the record is the drawn layout's standard start state."""
    np.random.seed(3)
    pool = LG.generate_layout_pool(6, dict(LG.DEFAULT_MDP_GEN_PARAMS, prop_empty=0.7, prop_feats=0.3), outer_shape=(6, 5))
    tab, starts, S = L.build_tables(pool, None)
    n = 6000
    rs = cpu.random_start(11, random_layout=True)
    a = np.zeros((n, S), np.int32)
    cpu.reset_random(tab, starts, a, rs)
    b = np.zeros((n, S), np.int32)
    cpu.reset_random(tab, starts, b, rs)
    assert np.array_equal(a, b)
    ids = a[:, 3] & 0xFF
    counts = np.bincount(ids, minlength=6)
    assert counts.min() > 800 and counts.max() < 1200
    want = starts[ids].copy()
    want[:, 3] |= 1 << 16
    assert np.array_equal(a, want)
    cpu.reset_random(tab, starts, b, rs)  # episode 2: a different assignment
    assert ((b[:, 3] >> 16) == 2).all() and (b[:, 3] & 0xFF != ids).mean() > 0.7
    c = np.zeros((n, S), np.int32)
    cpu.reset_random(tab, starts, c, cpu.random_start(12, random_layout=True))
    assert (c[:, 3] & 0xFF != ids).mean() > 0.7
    # auto-reset inside a rollout redraws too, and the next episode runs on the new layout's tables
    acts = np.random.RandomState(0).randint(0, 6, size=(25, n, 2)).astype(np.int32)
    st = a.copy()
    sparse, shaped, done, events = cpu.rollout(tab, starts, st, acts, horizon=10, flags=1, n_threads=2, rs=rs)
    assert ((st[:, 3] >> 16) == 3).all() and (st[:, 0] == 5).all()
    chk = np.zeros((n, S), np.int32)
    for _ in range(3):
        cpu.reset_random(tab, starts, chk, rs)
    assert np.array_equal(st[:, 3] & ~0xFF00, chk[:, 3])  # bits 8-15 count dishes dropped on counters since


def test_host_random_start_states_reproduce_the_reference_under_a_numpy_seed():
    """The drop-in mdp's get_random_start_state_fn (overcooked_mdp.py:1307-1369) against draws of the reference stored by
    tools/make_golden.py: joint positions, pot soups (idle / cooking), held dishes / onions / finished soups."""
    from overcooked_ai_b200.mdp import OvercookedGridworld

    with open(os.path.join(os.path.dirname(GOLD), "random_start_host.json")) as fh:
        gold = json.load(fh)
    mdps = {}
    kinds = set()
    for key, want in gold.items():
        name, pos, thr, seed = key.split("|")
        m = mdps.setdefault(name, OvercookedGridworld.from_layout_name(name))
        np.random.seed(int(seed))
        fn = m.get_random_start_state_fn(random_start_pos=bool(int(pos)), rnd_obj_prob_thresh=float(thr))
        for w in want:
            st = fn()
            got = json.loads(json.dumps(st.to_dict()))
            got["objects"].sort(key=lambda o: o["position"]), w["objects"].sort(key=lambda o: o["position"])
            assert got == w, key
            L.pack_state(m.compiled, st, 0, m.compiled.state_words)  # and the engine can hold it
            kinds |= {p["held_object"]["name"] for p in got["players"] if p["held_object"]} | {"pot:%s" % o["is_cooking"] for o in got["objects"]}
    assert {"dish", "onion", "soup", "pot:True", "pot:False"} <= kinds
    env = OvercookedEnv.from_mdp(mdps["cramped_room"], start_state_fn=mdps["cramped_room"].get_random_start_state_fn(True, 0.8), horizon=400)
    first = env.state
    assert any(env.reset() or env.state != first for _ in range(5))  # overcooked_test.py:1288-1309

"""GPU tests of the drop-in single-environment adapters (reference call surface, N=1 launches).
They read like the reference's own tests (testing/overcooked_test.py): build an mdp from a layout
name, wrap it in an env, step joint actions given as Action values, compare states and infos."""
import json

import numpy as np
import pytest

from helpers import GOLD, Trace
from overcooked_ai_b200 import layout as L
from overcooked_ai_b200.actions import Action, Direction
from overcooked_ai_b200.env import OvercookedEnv
from overcooked_ai_b200.mdp import OvercookedGridworld
from overcooked_ai_b200.state import ObjectState, OvercookedState, PlayerState, SoupState

pytestmark = pytest.mark.gpu

n, s, e, w = Direction.NORTH, Direction.SOUTH, Direction.EAST, Direction.WEST
stay, interact = Action.STAY, Action.INTERACT


def test_start_state_and_first_transition():
    """testing/overcooked_test.py:398-414, 468-514: start positions, one [n, e] transition, env/mdp agreement."""
    mdp = OvercookedGridworld.from_layout_name("mdp_test")
    start = mdp.get_standard_start_state()
    assert start.player_positions == ((1, 2), (3, 1)) and start.player_orientations == (n, n)
    new_state, infos = mdp.get_state_transition(start, (n, e))
    assert new_state.player_positions == ((1, 1), (3, 1))
    assert new_state.player_orientations == (n, e) and new_state.timestep == 1
    assert infos["sparse_reward_by_agent"] == [0, 0] and infos["shaped_reward_by_agent"] == [0, 0]
    assert set(infos["event_infos"]) == set(L.EVENT_TYPES) and not any(any(v) for v in infos["event_infos"].values())
    env = OvercookedEnv.from_mdp(mdp, horizon=10, info_level=0)
    s2, r, done, info = env.step((n, e))
    assert s2 == new_state and r == 0 and not done and env.state == new_state
    assert info["agent_infos"] == [{}, {}] and info["phi_s"] is None


def test_env_replays_reference_golden_trajectory_prefix():
    """The reference's mdp_test golden trajectory through OvercookedEnv.step with Action values."""
    tr = Trace(GOLD + "/dynamics_mdp_test.npz")
    mdp = OvercookedGridworld.from_layout_name("mdp_test")
    env = OvercookedEnv.from_mdp(mdp, horizon=1500, info_level=0)
    assert np.array_equal(L.pack_state(mdp.compiled, env.state), tr.states[0, 0])
    total = 0
    for t in range(200):  # covers the delivery at t=142
        ja = tuple(Action.INDEX_TO_ACTION[a] for a in tr.actions[0, t])
        st, r, done, info = env.step(ja)
        assert np.array_equal(L.pack_state(mdp.compiled, st), tr.states[0, t + 1]), t
        assert r == tr.sparse[0, t] and info["sparse_r_by_agent"] == tr.sparse2[0, t].tolist()
        assert info["shaped_r_by_agent"] == tr.shaped[0, t].tolist()
        total += r
    assert total == 10
    assert env.game_stats["cumulative_sparse_rewards_by_agent"].sum() == 10
    assert sum(len(x) for x in env.game_stats["soup_delivery"]) == 1


def test_episode_info_and_done():
    mdp = OvercookedGridworld.from_layout_name("cramped_room")
    env = OvercookedEnv.from_mdp(mdp, horizon=5, info_level=0)
    for t in range(5):
        st, r, done, info = env.step((stay, interact))
        assert done == (t == 4)
    assert info["episode"]["ep_length"] == 5 and info["episode"]["ep_sparse_r"] == 0
    assert set(info["episode"]) == {"ep_game_stats", "ep_sparse_r", "ep_shaped_r", "ep_sparse_r_by_agent", "ep_shaped_r_by_agent", "ep_length"}
    with pytest.raises(AssertionError):  # overcooked_env.py:255
        env.step((stay, stay))
    env.reset()
    assert env.state.timestep == 0 and not env.is_done()


def test_error_behaviour_matches_reference():
    mdp = OvercookedGridworld.from_layout_name("cramped_room")
    start = mdp.get_standard_start_state()
    with pytest.raises(ValueError):  # overcooked_mdp.py:1394-1398
        mdp.get_state_transition(start, (n, "jump"))
    with pytest.raises(ValueError):
        mdp.get_state_transition(start, ((2, 0), stay))
    bad = OvercookedState([PlayerState((0, 0), n), PlayerState((3, 1), n)], {}, all_orders=mdp.start_all_orders)
    with pytest.raises(AssertionError):  # player on a counter: _check_valid_state :1924
        mdp.get_state_transition(bad, (stay, stay))
    overlap = OvercookedState([PlayerState((1, 1), n), PlayerState((1, 1), n)], {}, all_orders=mdp.start_all_orders)
    with pytest.raises(AssertionError):
        mdp.get_state_transition(overlap, (stay, stay))
    floor_obj = OvercookedState([PlayerState((1, 1), n), PlayerState((3, 1), n)], {(2, 1): ObjectState("onion", (2, 1))},
                                all_orders=mdp.start_all_orders)
    with pytest.raises(AssertionError):  # loose object on the floor :1938
        mdp.get_state_transition(floor_obj, (stay, stay))


def test_scripted_soup_cycle_cramped_room():
    """Known-answer script: 3 onions -> cook 20 -> dish -> plate -> deliver = 20, with the shaped rewards
    3+3+3 (potting), 3 (useful dish pickup), 5 (soup pickup) (BASE_REW_SHAPING_PARAMS, quirk Q9)."""
    mdp = OvercookedGridworld.from_layout_name("cramped_room")
    env = OvercookedEnv.from_mdp(mdp, horizon=400, info_level=0)
    # player 0 starts at (1,2) facing north; onion dispenser at (0,1), pot at (2,0), dish at (1,3), serve at (3,3)
    def p0(*acts):
        out = []
        for a in acts:
            out.append(env.step((a, stay)))
        return out
    shaped = 0
    for _ in range(3):
        p0(n, w, interact)             # to (1,1), face the dispenser, take an onion
        res = p0(e, n, interact)       # to (2,1), face the pot, put it in
        shaped += res[-1][3]["shaped_r_by_agent"][0]
        assert res[-1][3]["shaped_r_by_agent"] == [3, 0]
        p0(w)                          # back to (1,1)
    assert shaped == 9
    p0(e, n, interact)                 # at (2,1) facing the pot: start cooking (tick 0 -> 1 this step, quirk Q4)
    soup = env.state.get_object((2, 0))
    assert soup.ingredients == ["onion"] * 3 and soup._cooking_tick == 1 and soup.is_cooking
    res = p0(w, s, s, interact)        # (1,1) -> (1,2), face the dish dispenser at (1,3), take a dish
    assert res[-1][3]["shaped_r_by_agent"] == [3, 0] and env.state.players[0].held_object.name == "dish"
    p0(n, e, n)                        # to (2,1), facing the pot
    while not env.state.get_object((2, 0)).is_ready:
        p0(stay)
    assert env.state.get_object((2, 0))._cooking_tick == 20
    res = p0(interact)
    assert res[-1][3]["shaped_r_by_agent"] == [5, 0] and env.state.players[0].held_object.name == "soup"
    assert not env.state.has_object((2, 0))
    res = p0(s, e, s, interact)        # (2,1) -> (2,2) -> (3,2), face the serving cell (3,3), deliver (P1 sits at (3,1))
    st, r, done, info = res[-1]
    assert r == 20 and info["sparse_r_by_agent"] == [20, 0] and st.players[0].held_object is None
    assert env.game_stats["soup_delivery"][0] == [st.timestep - 1]
    assert env.game_stats["cumulative_shaped_rewards_by_agent"].tolist() == [17, 0]


def test_adapter_encodings_match_fixture():
    tr = Trace(GOLD + "/trace_counter_circuit.npz")
    d = tr.data
    mdp = OvercookedGridworld.from_layout_name("counter_circuit")
    for k in range(0, len(d["obs_states"]), 97):
        st = L.unpack_state(mdp.compiled, d["obs_states"][k])
        enc = mdp.lossless_state_encoding(st, horizon=400)
        assert enc[0].dtype == np.int64 and enc[0].shape == (9, 5, 26)
        assert np.array_equal(np.stack(enc), d["obs_lossless"][k])
        f = mdp.featurize_state(st, None, num_pots=2)
        assert f[0].dtype == np.float64 and np.array_equal(np.stack(f), d["obs_feat_2"][k])


def test_featurization_symmetry():
    """testing/overcooked_test.py:1095-1128: swapping the players swaps the two observations."""
    tr = Trace(GOLD + "/trace_cramped_room.npz")
    mdp = OvercookedGridworld.from_layout_name("cramped_room")
    for k in range(0, len(tr.data["obs_states"]), 211):
        st = L.unpack_state(mdp.compiled, tr.data["obs_states"][k])
        sw = OvercookedState(list(reversed(st.players)), st.objects, bonus_orders=mdp.start_bonus_orders,
                             all_orders=mdp.start_all_orders, timestep=st.timestep)
        a0, a1 = mdp.lossless_state_encoding(st)
        b0, b1 = mdp.lossless_state_encoding(sw)
        assert np.array_equal(a0, b1) and np.array_equal(a1, b0)
        f0, f1 = mdp.featurize_state(st, None)
        g0, g1 = mdp.featurize_state(sw, None)
        assert np.array_equal(f0, g1) and np.array_equal(f1, g0)


def test_potential_function_and_display_phi():
    """testing/overcooked_test.py:607-999 style: phi through the drop-in mdp / env, exact reference values."""
    g = np.load(GOLD + "/potential.npz")
    tr = Trace(GOLD + "/trace_mdp_test.npz")
    mdp = OvercookedGridworld.from_layout_name("mdp_test")
    for k in range(0, len(tr.data["obs_states"]), 301):
        st = L.unpack_state(mdp.compiled, tr.data["obs_states"][k])
        assert mdp.potential_function(st, None, gamma=0.99) == g["mdp_test__phi"][k, 0]
        assert mdp.potential_function(st, None, gamma=0.9) == g["mdp_test__phi"][k, 1]
    env = OvercookedEnv.from_mdp(mdp, horizon=20, info_level=0)
    phi0 = env.potential()
    s1, r, done, info = env.step((n, interact), display_phi=True)
    assert info["phi_s"] == phi0 and info["phi_s_prime"] == env.potential()


def test_gym_wrapper_single_env():
    """overcooked_env.py:782-909 through the drop-in classes: seeded primary-agent draw, (primary, other) ordering."""
    from overcooked_ai_b200.env import Overcooked

    mdp = OvercookedGridworld.from_layout_name("cramped_room")
    base = OvercookedEnv.from_mdp(mdp, horizon=6, info_level=0)
    np.random.seed(3)
    gym = Overcooked(base, base.lossless_state_encoding_mdp)
    np.random.seed(3)
    expect_idx = [int(np.random.choice([0, 1])) for _ in range(3)]
    np.random.seed(3)
    for k in range(3):
        obs = gym.reset()
        assert gym.agent_idx == expect_idx[k] and obs["other_agent_env_idx"] == 1 - gym.agent_idx
        p0, p1 = base.lossless_state_encoding_mdp(base.state)
        want = (p0, p1) if gym.agent_idx == 0 else (p1, p0)
        assert np.array_equal(obs["both_agent_obs"][0], want[0]) and np.array_equal(obs["both_agent_obs"][1], want[1])
        # the primary agent walks west, the other stays: it must be player `agent_idx` that moved
        before = base.state.players[gym.agent_idx].position
        obs, r, done, info = gym.step((3, 4))
        moved = base.state.players[gym.agent_idx]
        assert moved.orientation == Direction.WEST and info["policy_agent_idx"] == gym.agent_idx
        assert base.state.players[1 - gym.agent_idx].orientation == Direction.NORTH
    with pytest.raises(AssertionError):
        gym.step((7, 0))


@pytest.mark.parametrize("seed", [0, 2, 7])
def test_baseline_config_1_through_the_dropin_env(seed):
    """BASELINE config 1: cramped_room, ONE environment, horizon 400, standard start, uniform random joint
    actions from the documented counter-based stream — the reference's OvercookedEnv trace (fixture) against the
    drop-in OvercookedEnv: every state, reward, done flag, and the episode info handed out with the last step."""
    import json

    from overcooked_ai_b200.actions import Action

    tr = Trace(GOLD + "/trace_config1_cramped_room.npz")
    want_info = json.loads(str(tr.data["episode_info"]))[seed]
    mdp = OvercookedGridworld.from_layout_name("cramped_room")
    env = OvercookedEnv.from_mdp(mdp, horizon=400, info_level=0)
    assert np.array_equal(L.pack_state(mdp.compiled, env.state, 0, tr.S), tr.states[seed, 0])
    for t in range(400):
        ja = tuple(Action.INDEX_TO_ACTION[int(a)] for a in tr.actions[seed, t])
        nxt, r, done, info = env.step(ja)
        assert np.array_equal(L.pack_state(mdp.compiled, nxt, 0, tr.S), tr.states[seed, t + 1]), t
        assert r == int(tr.sparse2[seed, t].sum()) and list(info["sparse_r_by_agent"]) == tr.sparse2[seed, t].tolist()
        assert list(info["shaped_r_by_agent"]) == tr.shaped[seed, t].tolist()
        assert done == (t == 399) and ("episode" in info) == done
    ep = info["episode"]
    assert ep["ep_length"] == want_info["ep_length"] == 400
    assert int(ep["ep_sparse_r"]) == want_info["ep_sparse_r"] and int(ep["ep_shaped_r"]) == want_info["ep_shaped_r"]
    assert [int(v) for v in ep["ep_sparse_r_by_agent"]] == want_info["ep_sparse_r_by_agent"]
    assert [int(v) for v in ep["ep_shaped_r_by_agent"]] == want_info["ep_shaped_r_by_agent"]
    for name, lists in want_info["game_stats"].items():
        assert [list(map(int, l)) for l in ep["ep_game_stats"][name]] == lists, name
    with pytest.raises(AssertionError):  # overcooked_env.py:255
        env.step(ja)

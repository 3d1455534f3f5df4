"""GPU tests of the vector-env front ends (gym wrapper / RLlib multi-agent wrapper, batched)."""
import numpy as np
import pytest
import torch

from helpers import lut_bytes
from oracle import cpu
from overcooked_ai_b200.batched import BatchedOvercookedEnv
from overcooked_ai_b200.vecenv import BatchedOvercookedGym, BatchedOvercookedMultiAgent

pytestmark = pytest.mark.gpu


def _np(t):
    return t.cpu().numpy()


@pytest.mark.parametrize("featurize", ["lossless", "features"])
def test_batched_gym_wrapper_against_oracle(featurize):
    """overcooked_env.py:842-909 semantics per environment: (primary, other) action order, observations in
    (primary, other) order, primary index redrawn exactly at episode boundaries."""
    n, horizon, T = 2003, 30, 70
    env = BatchedOvercookedEnv("coordination_ring", n, horizon=horizon, auto_reset=True)
    gym = BatchedOvercookedGym(env, featurize=featurize, seed=123)
    l = env.layouts[0]
    lut = lut_bytes([l])
    ref = _np(env.state).copy()
    rng = np.random.RandomState(1)

    def expected_obs(state, idx):
        if featurize == "lossless":
            e = cpu.encode_lossless(env._tab_host, state, l.width, l.height, horizon).astype(np.float32)
        else:
            e = cpu.featurize(env._tab_host, lut, state, 2).astype(np.float32)
        ar = np.arange(n)
        return np.stack([e[ar, idx], e[ar, 1 - idx]], 1)

    obs = gym.reset()
    idx = _np(gym.agent_idx).copy()
    assert set(np.unique(idx)) == {0, 1}
    assert np.array_equal(_np(obs["both_agent_obs"]), expected_obs(ref, idx))
    assert np.array_equal(_np(obs["other_agent_env_idx"]), 1 - idx)
    for t in range(T):
        a = rng.randint(0, 6, size=(n, 2)).astype(np.int32)
        a[rng.rand(n, 2) < 0.3] = 5
        joint = np.where(idx[:, None] == 0, a, a[:, ::-1]).astype(np.int32)
        sp, sh, dn, ev = cpu.step(env._tab_host, env._starts_host, ref, joint, horizon=horizon, flags=1)
        obs, reward, done, info = gym.step(torch.from_numpy(a).cuda())
        assert np.array_equal(_np(reward), sp) and np.array_equal(_np(done), dn)
        assert np.array_equal(_np(info["policy_agent_idx"]), idx) and np.array_equal(_np(info["shaped_r_by_agent"]), sh)
        new_idx = _np(gym.agent_idx).copy()
        assert np.array_equal(new_idx[dn == 0], idx[dn == 0])  # unchanged inside an episode
        idx = new_idx
        assert np.array_equal(_np(obs["both_agent_obs"]), expected_obs(ref, idx)), t
    assert np.array_equal(_np(env.state), ref)


def test_baselines_reproducible_draws_one_index_for_all():
    env = BatchedOvercookedEnv("cramped_room", 64, horizon=5, auto_reset=True)
    gym = BatchedOvercookedGym(env, seed=0, baselines_reproducible=True)
    for _ in range(8):
        gym.reset()
        assert len(torch.unique(gym.agent_idx)) == 1


def test_multi_agent_wrapper_rewards_and_annealing():
    n = 1500
    env = BatchedOvercookedEnv("cramped_room", n, horizon=400, auto_reset=True)
    ma = BatchedOvercookedMultiAgent(env, reward_shaping_factor=1.0, reward_shaping_horizon=1000)
    assert ma._anneal(1.0, 500, 1000) == 0.5 and ma._anneal(1.0, 2000, 1000) == 0 and ma._anneal(0.7, 123, 0) == 0.7
    ma.anneal_reward_shaping_factor(250)
    assert abs(ma.reward_shaping_factor - 0.75) < 1e-12
    obs = ma.reset()
    ref = _np(env.state).copy()
    rng = np.random.RandomState(2)
    l = env.layouts[0]
    for t in range(40):
        a = rng.randint(0, 6, size=(n, 2)).astype(np.int32)
        a[rng.rand(n, 2) < 0.4] = 5
        sp, sh, dn, ev = cpu.step(env._tab_host, env._starts_host, ref, a, horizon=400, flags=1)
        obs, rew, dones, infos = ma.step({"ppo_0": torch.from_numpy(a[:, 0].copy()).cuda(), "ppo_1": torch.from_numpy(a[:, 1].copy()).cuda()})
        for i, ag in enumerate(("ppo_0", "ppo_1")):
            assert np.allclose(_np(rew[ag]), sp + 0.75 * sh[:, i])
        assert not _np(dones["__all__"]).any()
    enc = cpu.encode_lossless(env._tab_host, ref, l.width, l.height, 400).astype(np.float32)
    assert obs["ppo_0"].dtype == torch.float32
    assert np.array_equal(_np(obs["ppo_0"]), enc[:, 0]) and np.array_equal(_np(obs["ppo_1"]), enc[:, 1])


def test_multi_agent_wrapper_use_phi_dense_reward():
    """rllib.py:314-329 with use_phi=True: reward_i = sparse + factor * (phi(s') - phi(s)), phi(s') on the
    terminal state of a finishing episode."""
    from overcooked_ai_b200 import layout as L

    n, horizon = 700, 12
    env = BatchedOvercookedEnv("cramped_room", n, horizon=horizon, auto_reset=False)
    ma = BatchedOvercookedMultiAgent(env, reward_shaping_factor=0.5, use_phi=True)
    ma.reset()
    pt, cst, gpow = L.build_potential_tables(env.layouts, 0.99)
    ref = _np(env.state).copy()
    rng = np.random.RandomState(4)
    for t in range(30):
        a = rng.randint(0, 6, size=(n, 2)).astype(np.int32)
        a[rng.rand(n, 2) < 0.4] = 5
        phi_s = cpu.potential(env._tab_host, pt, cst, gpow, ref)
        sp, sh, dn, ev = cpu.step(env._tab_host, env._starts_host, ref, a, horizon=horizon, flags=0)
        phi_n = cpu.potential(env._tab_host, pt, cst, gpow, ref)
        obs, rew, dones, infos = ma.step({"ppo_0": torch.from_numpy(a[:, 0].copy()).cuda(), "ppo_1": torch.from_numpy(a[:, 1].copy()).cuda()})
        want = (sp.astype(np.float32) + np.float32(0.5) * (phi_n - phi_s).astype(np.float32)).astype(np.float32)
        assert np.allclose(_np(rew["ppo_0"]), want, rtol=0, atol=1e-5) and np.array_equal(_np(rew["ppo_0"]), _np(rew["ppo_1"]))
        assert np.array_equal(_np(dones["__all__"]), dn != 0)
        ref[dn != 0] = env._starts_host[0]
        assert np.array_equal(_np(env.state), ref)


def test_episode_stats_match_reference_game_stats_on_greedy_games():
    """EpisodeStats == len() of the reference env's game_stats lists + cumulative rewards (overcooked_env.py:382-401),
    on the 5 GreedyHumanModel games of the fixture (9 deliveries each)."""
    from helpers import GOLD, Trace
    from overcooked_ai_b200.batched import EpisodeStats

    tr = Trace(GOLD + "/greedy_cramped_room.npz")
    env = BatchedOvercookedEnv("cramped_room", tr.E, horizon=400, auto_reset=True)
    stats = EpisodeStats(env)
    acts = torch.from_numpy(np.ascontiguousarray(tr.actions.transpose(1, 0, 2))).cuda()
    fin = None
    for t in range(tr.T):
        out = env.step(acts[t])
        f = stats.update(*out)
        assert (f is None) == (t < tr.T - 1)
        fin = f or fin
    assert fin["ep_length"].tolist() == [400] * 5 and fin["ep_sparse_r"].tolist() == [180] * 5
    want_counts = np.stack([((tr.events[e][:, :, None] >> np.arange(25)) & 1).sum(0) for e in range(tr.E)])
    assert np.array_equal(_np(fin["ep_game_stats"]), want_counts)
    assert np.array_equal(_np(fin["ep_shaped_r_by_agent"]), tr.shaped.sum(1))
    assert np.array_equal(_np(fin["ep_sparse_r_by_agent"]), tr.sparse2.sum(1))
    assert int(fin["ep_game_stats"][:, :, 15].sum()) == 45  # soup_delivery
    assert not stats.event_counts.any() and not stats.ep_length.any()

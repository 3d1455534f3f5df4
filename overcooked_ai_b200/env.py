"""``OvercookedEnv`` with the reference's surface (src/overcooked_ai_py/mdp/overcooked_env.py:33-483),
stepping through the CUDA engine.

Kept: ``from_mdp`` :117-141, ``step`` :244-274 (returns ``(next_state, sparse_reward, done, env_info)``),
``reset`` :288-319, ``is_done`` :321-325, ``lossless_state_encoding_mdp`` / ``featurize_state_mdp``
:276-286, ``game_stats`` bookkeeping :382-401, episode info :363-380, ``execute_plan`` :407-423.
Not kept: planners (``mlam`` / ``mp``), rendering, agent rollouts (``run_agents`` / ``get_rollouts``) —
host drivers outside the hot path.
"""
import numpy as np

from overcooked_ai_b200.layout import EVENT_TYPES
from overcooked_ai_b200.mdp import OvercookedGridworld

DEFAULT_ENV_PARAMS = {"horizon": 400}
MAX_HORIZON = 1e10


class OvercookedEnv(object):
    def __init__(self, mdp_generator_fn, start_state_fn=None, horizon=MAX_HORIZON, mlam_params=None, info_level=0,
                 num_mdp=1, initial_info={}, **kwargs):
        assert callable(mdp_generator_fn), (
            "OvercookedEnv takes in a OvercookedGridworld generator function. "
            "If trying to instantiate directly from a OvercookedGridworld instance, use the OvercookedEnv.from_mdp method"
        )
        # mlam_params: the reference's planner parameters (default NO_COUNTERS_PARAMS, overcooked_env.py:92-100).  The
        # planners themselves are outside this engine; featurize_state bakes exactly those defaults into its table.
        self.mlam_params = mlam_params
        self.num_mdp = num_mdp
        self.variable_mdp = num_mdp > 1
        self.mdp_generator_fn = mdp_generator_fn
        self.horizon = horizon
        self.start_state_fn = start_state_fn
        self.info_level = info_level
        self.reset(outside_info=initial_info)

    @staticmethod
    def from_mdp(mdp, start_state_fn=None, horizon=MAX_HORIZON, mlam_params=None, info_level=1, num_mdp=None, **kwargs):
        assert isinstance(mdp, OvercookedGridworld)
        if num_mdp is not None:
            assert num_mdp == 1
        return OvercookedEnv(lambda _ignored: mdp, start_state_fn=start_state_fn, horizon=horizon,
                             mlam_params=mlam_params, info_level=info_level, num_mdp=1)

    @property
    def env_params(self):
        return {"start_state_fn": self.start_state_fn, "horizon": self.horizon, "mlam_params": self.mlam_params,
                "info_level": self.info_level, "num_mdp": self.num_mdp}

    def step(self, joint_action, joint_agent_action_info=None, display_phi=False):
        assert not self.is_done()
        if joint_agent_action_info is None:
            joint_agent_action_info = [{}, {}]
        next_state, mdp_infos = self.mdp.get_state_transition(self.state, joint_action, display_phi)
        self._update_game_stats(mdp_infos)
        self.state = next_state
        done = self.is_done()
        env_info = self._prepare_info_dict(joint_agent_action_info, mdp_infos)
        if done:
            self._add_episode_info(env_info)
        return (next_state, sum(mdp_infos["sparse_reward_by_agent"]), done, env_info)

    def lossless_state_encoding_mdp(self, state):
        return self.mdp.lossless_state_encoding(state, self.horizon)

    def featurize_state_mdp(self, state, num_pots=2):
        p = self.mlam_params
        if p is not None and (p.get("counter_goals") or p.get("counter_drop") or p.get("counter_pickup")):
            raise NotImplementedError("featurize_state is built for the env's default planner parameters "
                                      "(NO_COUNTERS_PARAMS: counters are never motion goals)")
        return self.mdp.featurize_state(state, None, num_pots=num_pots)

    def potential(self, mlam=None, state=None, gamma=0.99):
        """overcooked_env.py:327-337"""
        return self.mdp.potential_function(state if state else self.state, gamma=gamma)

    def reset(self, regen_mdp=True, outside_info={}):
        if regen_mdp:
            self.mdp = self.mdp_generator_fn(outside_info)
        if self.start_state_fn is None:
            self.state = self.mdp.get_standard_start_state()
        else:
            self.state = self.start_state_fn()
        events_dict = {k: [[] for _ in range(self.mdp.num_players)] for k in EVENT_TYPES}
        rewards_dict = {
            "cumulative_sparse_rewards_by_agent": np.array([0] * self.mdp.num_players),
            "cumulative_shaped_rewards_by_agent": np.array([0] * self.mdp.num_players),
        }
        self.game_stats = {**events_dict, **rewards_dict}

    def is_done(self):
        return self.state.timestep >= self.horizon or self.mdp.is_terminal(self.state)

    def _prepare_info_dict(self, joint_agent_action_info, mdp_infos):
        env_info = {"agent_infos": [joint_agent_action_info[i] for i in range(self.mdp.num_players)]}
        env_info["sparse_r_by_agent"] = mdp_infos["sparse_reward_by_agent"]
        env_info["shaped_r_by_agent"] = mdp_infos["shaped_reward_by_agent"]
        env_info["phi_s"] = mdp_infos.get("phi_s", None)
        env_info["phi_s_prime"] = mdp_infos.get("phi_s_prime", None)
        return env_info

    def _add_episode_info(self, env_info):
        gs = self.game_stats
        env_info["episode"] = {
            "ep_game_stats": gs,
            "ep_sparse_r": sum(gs["cumulative_sparse_rewards_by_agent"]),
            "ep_shaped_r": sum(gs["cumulative_shaped_rewards_by_agent"]),
            "ep_sparse_r_by_agent": gs["cumulative_sparse_rewards_by_agent"],
            "ep_shaped_r_by_agent": gs["cumulative_shaped_rewards_by_agent"],
            "ep_length": self.state.timestep,
        }
        return env_info

    def _update_game_stats(self, infos):
        self.game_stats["cumulative_sparse_rewards_by_agent"] += np.array(infos["sparse_reward_by_agent"])
        self.game_stats["cumulative_shaped_rewards_by_agent"] += np.array(infos["shaped_reward_by_agent"])
        for event_type, by_agent in infos["event_infos"].items():
            for idx, happened in enumerate(by_agent):
                if happened:  # timestep is logged before the tick, so events carry 0..horizon-1
                    self.game_stats[event_type][idx].append(self.state.timestep)

    def execute_plan(self, start_state, joint_action_plan, display=False):
        self.state = start_state
        done = False
        for joint_action in joint_action_plan:
            self.step(joint_action)
            done = self.is_done()
            if done:
                break
        successor_state = self.state
        self.reset(False)
        return successor_state, done


class Overcooked(object):
    """The reference's gym wrapper (overcooked_env.py:782-909) over the drop-in env, single environment:
    the primary agent's index is redrawn with ``np.random.choice([0, 1])`` at every reset exactly as the
    reference does (so a seeded run assigns the same indices), actions arrive as (primary, other) action
    INDICES and observations leave as (primary, other).  No gymnasium dependency: ``observation_space`` /
    ``action_space`` are plain descriptions.  For many environments use vecenv.BatchedOvercookedGym."""

    env_name = "Overcooked-v0"

    def __init__(self, base_env, featurize_fn, baselines_reproducible=False):
        if baselines_reproducible:
            np.random.seed(0)  # overcooked_env.py:821-832
        self.base_env = base_env
        self.featurize_fn = featurize_fn
        dummy = self.featurize_fn(self.base_env.mdp.get_standard_start_state())[0]
        self.observation_space = {"shape": tuple(dummy.shape), "low": 0.0, "high": float("inf"), "dtype": np.float32}
        self.action_space = {"n": 6}
        self.reset()

    def _both(self, state):
        ob_p0, ob_p1 = self.featurize_fn(state)
        return (ob_p0, ob_p1) if self.agent_idx == 0 else (ob_p1, ob_p0)

    def step(self, action):
        assert all(isinstance(a, (int, np.integer)) and 0 <= a < 6 for a in action), "%r (%s) invalid" % (action, type(action))
        from overcooked_ai_b200.actions import Action

        agent_action, other_agent_action = [Action.INDEX_TO_ACTION[a] for a in action]
        joint_action = (agent_action, other_agent_action) if self.agent_idx == 0 else (other_agent_action, agent_action)
        next_state, reward, done, env_info = self.base_env.step(joint_action)
        env_info["policy_agent_idx"] = self.agent_idx
        if "episode" in env_info:
            env_info["episode"]["policy_agent_idx"] = self.agent_idx
        obs = {"both_agent_obs": self._both(next_state), "overcooked_state": next_state, "other_agent_env_idx": 1 - self.agent_idx}
        return obs, reward, done, env_info

    def reset(self):
        self.base_env.reset()
        self.mdp = self.base_env.mdp
        self.agent_idx = np.random.choice([0, 1])
        return {"both_agent_obs": self._both(self.base_env.state), "overcooked_state": self.base_env.state,
                "other_agent_env_idx": 1 - self.agent_idx}

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
        """(next_state, summed sparse reward, done, env_info) — overcooked_env.py:244-274."""
        assert not self.is_done()
        agent_infos = [{}, {}] if joint_agent_action_info is None else joint_agent_action_info
        t = self.state.timestep  # events are logged with the pre-transition timestep: 0 .. horizon - 1
        self.state, infos = self.mdp.get_state_transition(self.state, joint_action, display_phi)
        sparse, shaped = infos["sparse_reward_by_agent"], infos["shaped_reward_by_agent"]
        self._returns += np.array([sparse, shaped])
        self._event_log.extend((t, agent, name) for name, flags in infos["event_infos"].items() for agent, hit in enumerate(flags) if hit)
        done = self.is_done()
        env_info = {
            "agent_infos": [agent_infos[i] for i in range(self.mdp.num_players)],
            "sparse_r_by_agent": sparse, "shaped_r_by_agent": shaped,
            "phi_s": infos.get("phi_s"), "phi_s_prime": infos.get("phi_s_prime"),
        }
        if done:
            env_info["episode"] = self._episode_summary()
        return (self.state, sum(sparse), done, env_info)

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
        """overcooked_env.py:288-319: new MDP (unless told not to), start state, empty episode statistics."""
        if regen_mdp:
            self.mdp = self.mdp_generator_fn(outside_info)
        self.state = self.mdp.get_standard_start_state() if self.start_state_fn is None else self.start_state_fn()
        # episode statistics live in the engine's terms — a log of (timestep, agent, event) and a [sparse|shaped, agent]
        # return array; `game_stats` renders them in the reference's shape on demand
        self._event_log = []
        self._returns = np.zeros((2, self.mdp.num_players), dtype=np.int64)

    @property
    def game_stats(self):
        """The reference's dict (:308-319, 382-401): per event name a list per agent of the timesteps it fired at, plus
        the two cumulative reward arrays."""
        stats = {name: [[] for _ in range(self.mdp.num_players)] for name in EVENT_TYPES}
        for t, agent, name in self._event_log:
            stats[name][agent].append(t)
        stats["cumulative_sparse_rewards_by_agent"] = self._returns[0].copy()
        stats["cumulative_shaped_rewards_by_agent"] = self._returns[1].copy()
        return stats

    def is_done(self):
        return self.state.timestep >= self.horizon or self.mdp.is_terminal(self.state)

    def _episode_summary(self):
        """The ``episode`` entry of the last transition's info dict (:363-380)."""
        gs = self.game_stats
        sparse, shaped = gs["cumulative_sparse_rewards_by_agent"], gs["cumulative_shaped_rewards_by_agent"]
        return {"ep_game_stats": gs, "ep_sparse_r": sum(sparse), "ep_shaped_r": sum(shaped),
                "ep_sparse_r_by_agent": sparse, "ep_shaped_r_by_agent": shaped, "ep_length": self.state.timestep}

    def execute_plan(self, start_state, joint_action_plan, display=False):
        """Runs the plan from ``start_state`` until it ends or the episode does; returns (state reached, done) and
        leaves the env reset on the same MDP (:407-423)."""
        self.state = start_state
        for joint_action in joint_action_plan:
            if self.step(joint_action)[2]:
                break
        reached, done = self.state, self.is_done() if len(joint_action_plan) else False
        self.reset(False)
        return reached, done


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

"""Vector-env front ends over BatchedOvercookedEnv: the callers either side of the hot path.

``BatchedOvercookedGym``         the reference's gym wrapper ``Overcooked`` (overcooked_env.py:782-909),
                                 N environments at once: a primary-agent index drawn per environment at
                                 every reset, actions given as (primary, other), observations returned
                                 as (primary, other) — the swap happens inside the observation kernels
                                 (``view_swap``), not as a second pass.
``BatchedOvercookedMultiAgent``  the self-play path of ``OvercookedMultiAgent.step / reset``
                                 (human_aware_rl/rllib/rllib.py:293-368) with ``use_phi=False``:
                                 per-agent float32 observations, reward = sparse + factor * shaped_i
                                 with the reference's linear annealing of the factor (:283-291, :358-368).

Randomness: the reference draws agent indices from ``np.random`` global state (overcooked_env.py:898),
which a device engine cannot reproduce bit for bit; here they come from a ``torch.Generator`` on the
env's device, seeded explicitly.  Everything downstream of the drawn indices is exact.
Episode boundaries follow the vector-env convention: with ``auto_reset`` the observation returned
for a finished environment is the first observation of its next episode (and its agent index has
been redrawn), ``done`` marks the boundary.
"""
import torch

from overcooked_ai_b200.batched import BatchedOvercookedEnv


class BatchedOvercookedGym(object):
    def __init__(self, env, featurize="lossless", num_pots=2, obs_dtype=torch.float32, seed=0, baselines_reproducible=False):
        assert isinstance(env, BatchedOvercookedEnv)
        assert featurize in ("lossless", "features")
        assert env.auto_reset, "the vector gym front end needs auto_reset=True"
        self.env, self.featurize, self.num_pots, self.obs_dtype = env, featurize, num_pots, obs_dtype
        self.gen = torch.Generator(device=env.device)
        self.gen.manual_seed(seed)
        # baselines_reproducible (overcooked_env.py:821-832): all envs share one index per reset
        self.sync = bool(baselines_reproducible)
        N = env.n_envs
        self.agent_idx = torch.zeros(N, dtype=torch.int32, device=env.device)
        self._joint = torch.zeros((N, 2), dtype=torch.int32, device=env.device)
        self._obs = None

    def _draw(self, mask=None):
        N = self.env.n_envs
        if self.sync:
            new = torch.randint(0, 2, (1,), device=self.env.device, generator=self.gen, dtype=torch.int32).expand(N)
        else:
            new = torch.randint(0, 2, (N,), device=self.env.device, generator=self.gen, dtype=torch.int32)
        if mask is None:
            self.agent_idx.copy_(new)
        else:
            self.agent_idx.copy_(torch.where(mask != 0, new, self.agent_idx))

    def _observe(self):
        if self.featurize == "lossless":
            self._obs = self.env.lossless_state_encoding(out=self._obs, dtype=self.obs_dtype, view_swap=self.agent_idx)
        else:
            self._obs = self.env.featurize_state(self.num_pots, out=self._obs, view_swap=self.agent_idx)
        return {"both_agent_obs": self._obs, "overcooked_state": self.env.state, "other_agent_env_idx": 1 - self.agent_idx}

    def reset(self):
        """overcooked_env.py:885-909"""
        self.env.reset()
        self._draw()
        return self._observe()

    def step(self, action):
        """action int32 [N, 2] = (primary agent's action, other agent's action), overcooked_env.py:842-883."""
        assert action.dtype == torch.int32 and tuple(action.shape) == (self.env.n_envs, 2)
        swap = (self.agent_idx != 0).unsqueeze(1)
        self._joint.copy_(torch.where(swap, action.flip(1), action))
        sparse, shaped, done, events = self.env.step(self._joint)
        info = {
            "policy_agent_idx": self.agent_idx.clone(),
            "shaped_r_by_agent": shaped,
            "events": events,
        }
        self._draw(done)  # finished envs were auto-reset: new episode, new primary agent
        return self._observe(), sparse, done, info


class BatchedOvercookedMultiAgent(object):
    AGENTS = ("ppo_0", "ppo_1")

    def __init__(self, env, reward_shaping_factor=0.0, reward_shaping_horizon=0, obs_dtype=torch.float32,
                 use_phi=False, gamma=0.99):
        """use_phi: dense reward = phi(s') - phi(s) for both agents (rllib.py:314-319) instead of the
        per-agent shaped reward; the potential of s' is taken BEFORE a finished environment is reset,
        so the env must be built with auto_reset=False (this wrapper resets finished envs itself)."""
        assert isinstance(env, BatchedOvercookedEnv)
        assert not (use_phi and env.auto_reset), "use_phi needs auto_reset=False (phi(s') is evaluated on the terminal state)"
        self.env = env
        self.use_phi, self.gamma = bool(use_phi), gamma
        self._phi = None
        self._initial_reward_shaping_factor = reward_shaping_factor
        self.reward_shaping_factor = reward_shaping_factor
        self.reward_shaping_horizon = reward_shaping_horizon
        self.obs_dtype = obs_dtype
        self._obs = None
        self._joint = torch.zeros((env.n_envs, 2), dtype=torch.int32, device=env.device)

    @staticmethod
    def _anneal(start_v, curr_t, end_t, end_v=0, start_t=0):
        """rllib.py:283-291"""
        if end_t == 0:
            return start_v
        fraction = max(1 - float(curr_t - start_t) / (end_t - start_t), 0)
        return fraction * start_v + (1 - fraction) * end_v

    def anneal_reward_shaping_factor(self, timesteps):
        """rllib.py:358-368"""
        self.reward_shaping_factor = self._anneal(self._initial_reward_shaping_factor, timesteps, self.reward_shaping_horizon)

    def _get_obs(self):
        """rllib.py:257-260: agent i gets view i of the lossless encoding, as float32."""
        self._obs = self.env.lossless_state_encoding(out=self._obs, dtype=self.obs_dtype)
        return {self.AGENTS[0]: self._obs[:, 0], self.AGENTS[1]: self._obs[:, 1]}

    def reset(self, regen_mdp=True):
        self.env.reset()
        if self.use_phi:
            self._phi = self.env.potential(self.gamma)
        return self._get_obs()

    def step(self, action_dict):
        """action_dict: {"ppo_0": int32[N], "ppo_1": int32[N]} -> (obs, rewards, dones, infos), rllib.py:293-342."""
        self._joint[:, 0].copy_(action_dict[self.AGENTS[0]])
        self._joint[:, 1].copy_(action_dict[self.AGENTS[1]])
        sparse, shaped, done, events = self.env.step(self._joint)
        sp = sparse.to(torch.float32)
        if self.use_phi:
            phi_next = self.env.potential(self.gamma)
            dense = (phi_next - self._phi).to(torch.float32)
            rewards = {a: sp + self.reward_shaping_factor * dense for a in self.AGENTS}
            self.env.reset(done)  # finished envs start their next episode now
            self._phi = torch.where(done != 0, self.env.potential(self.gamma), phi_next)
        else:
            rewards = {a: sp + self.reward_shaping_factor * shaped[:, i].to(torch.float32) for i, a in enumerate(self.AGENTS)}
            if not self.env.auto_reset:
                self.env.reset(done)  # finished envs start their next episode now (with auto_reset the kernel already did)
        d = done != 0
        dones = {self.AGENTS[0]: d, self.AGENTS[1]: d, "__all__": d}
        info = {"sparse_r": sparse, "shaped_r_by_agent": shaped, "events": events}
        return self._get_obs(), rewards, dones, {self.AGENTS[0]: info, self.AGENTS[1]: info}

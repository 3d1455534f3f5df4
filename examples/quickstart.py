#!/usr/bin/env python
"""Quick tour of the engine on one GPU:  python examples/quickstart.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from overcooked_ai_b200.actions import Action, Direction  # noqa: E402
from overcooked_ai_b200.batched import BatchedOvercookedEnv, EpisodeStats  # noqa: E402
from overcooked_ai_b200.env import OvercookedEnv  # noqa: E402
from overcooked_ai_b200.mdp import OvercookedGridworld  # noqa: E402
from overcooked_ai_b200.vecenv import BatchedOvercookedGym  # noqa: E402

# 1. the reference's own call surface, one environment (N = 1 launches underneath)
mdp = OvercookedGridworld.from_layout_name("cramped_room")
env = OvercookedEnv.from_mdp(mdp, horizon=400, info_level=0)
state, reward, done, info = env.step((Direction.NORTH, Action.INTERACT))
print("drop-in step:", state.players[0], "| reward", reward, "| phi", env.potential())
print("lossless obs shape per player:", env.lossless_state_encoding_mdp(state)[0].shape)

# 2. the batched tensor API: 65 536 environments, mixed layouts, one launch per transition
benv = BatchedOvercookedEnv(["cramped_room", "coordination_ring"], 65536, horizon=400, auto_reset=True)
stats = EpisodeStats(benv)
for t in range(400):
    actions = torch.randint(0, 6, (benv.n_envs, 2), dtype=torch.int32, device="cuda")
    finished = stats.update(*benv.step(actions))
print("batched: episodes finished at t=400:", len(finished["env_index"]), "| mean shaped reward", float(finished["ep_shaped_r"].float().mean()))

# 3. a whole horizon in ONE launch, observations for a CNN and for an MLP, the shaping potential
actions = torch.randint(0, 6, (400, benv.n_envs, 2), dtype=torch.int32, device="cuda")
sparse, shaped, done, events = benv.rollout(actions)
obs = benv.lossless_state_encoding(dtype=torch.float32)  # list: one [n, 2, W, H, 26] tensor per grid shape
feat = benv.featurize_state(num_pots=2)                  # [N, 2, 96]
phi = benv.potential(gamma=0.99)                         # [N] float64, bit-identical to the reference
print("rollout:", tuple(sparse.shape), "| obs", [tuple(o.shape) for o in obs], "| feat", tuple(feat.shape), "| phi", float(phi.mean()))

# 4. the gym wrapper, vectorised: primary agent index drawn per environment at every reset
gym = BatchedOvercookedGym(BatchedOvercookedEnv("cramped_room", 1024, horizon=400, auto_reset=True), featurize="lossless")
o = gym.reset()
o, r, d, info = gym.step(torch.randint(0, 6, (1024, 2), dtype=torch.int32, device="cuda"))
print("gym:", tuple(o["both_agent_obs"].shape), "| primary agent of env 0:", int(info["policy_agent_idx"][0]))

# 5. variable MDP: a pool of procedurally generated layouts (the reference's LayoutGenerator; a numpy seed gives
#    the reference's grids), each environment redrawing its layout from the pool at every (auto-)reset on the device
import numpy as np  # noqa: E402

from overcooked_ai_b200.layout_generator import generate_layout_pool  # noqa: E402

np.random.seed(0)
pool = generate_layout_pool(32, {"inner_shape": (6, 5), "prop_empty": 0.6, "prop_feats": 0.3, "display": False,
                                 "start_all_orders": [{"ingredients": ["onion", "onion", "onion"]}]}, outer_shape=(7, 6),
                            skip_unsupported=True)
venv = BatchedOvercookedEnv(pool, 16384, horizon=100, auto_reset=True, random_layout=True, random_start_pos=True, seed=3)
before = venv.layout_ids().clone()
venv.rollout(torch.randint(0, 6, (100, venv.n_envs, 2), dtype=torch.int32, device="cuda"))
print("variable MDP: %d layouts in the pool, e.g.\n  %s\n  %.0f %% of the environments moved to another layout after one episode"
      % (len(pool), "\n  ".join("".join(r) for r in pool[0].terrain_mtx), 100 * float((venv.layout_ids() != before).float().mean())))

# 6. host buffers in, host buffers out: one byte of joint action in, one int16 word out per env-step over PCIe,
#    driven natively (ovc_pipeline_*), expanded to dense arrays on the host cores
from overcooked_ai_b200 import wire  # noqa: E402
from overcooked_ai_b200.batched import HostRolloutPipeline  # noqa: E402

henv = BatchedOvercookedEnv("cramped_room", 65536, horizon=400, auto_reset=True)
pipe = HostRolloutPipeline(henv, 400, chunk=200, codes=True)
h_actions = torch.from_numpy(wire.pack_actions(np.random.randint(0, 6, size=(400, henv.n_envs, 2)))).pin_memory()
words = pipe.run(h_actions)[3]
torch.cuda.synchronize()
dense = henv.expand_codes(words, events=True)
print("host pipeline:", tuple(words.shape), words.dtype, "->", {k: (tuple(v.shape), str(v.dtype)) for k, v in dense.items()},
      "| shaped reward collected:", int(dense["shaped"].sum(dtype=torch.int64)))

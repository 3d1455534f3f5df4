#!/usr/bin/env python
"""Every kernel family at smoke sizes, for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):
K1 with each record-I/O strategy (2-D TMA tile, 1-D bulk TMA, direct), K5 (the rollout kernel, int32 and
host-transfer formats, standard and random-start auto-resets, 16- / 32- / 64-word records, a partial last tile),
the round-1 fused path, K4 reset (copy + random), K2, K3, K6 and the host-buffer pipeline.  Results are checked
against the CPU oracle on the way, so a run under the sanitizer is also a parity run.

    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch  # noqa: E402

from oracle import cpu  # noqa: E402
from overcooked_ai_b200 import _native, wire  # noqa: E402
from overcooked_ai_b200.batched import BatchedOvercookedEnv, HostRolloutPipeline  # noqa: E402

rng = np.random.RandomState(0)


def acts_for(T, n):
    a = rng.randint(0, 6, size=(T, n, 2)).astype(np.int32)
    a[rng.rand(T, n, 2) < 0.3] = 5
    return a


def check(env, acts, split, rs=None):
    ref_state = env.state.cpu().numpy().copy()
    ref = cpu.rollout(env._tab_host, env._starts_host, ref_state, acts, horizon=env.horizon, flags=int(env.auto_reset), n_threads=2, rs=rs)
    d = torch.from_numpy(acts).cuda()
    for t in range(split):
        out = env.step(d[t])
        for g, w in zip(out, ref):
            assert np.array_equal(g.cpu().numpy(), w[t])
    out = env.rollout(d[split:].contiguous())
    for g, w in zip(out, ref):
        assert np.array_equal(g.cpu().numpy(), w[split:])
    assert np.array_equal(env.state.cpu().numpy(), ref_state)


n = 333  # partial last tile for every tile size
for io in (_native.IO_TMA_TENSOR, _native.IO_TMA_BULK, _native.IO_DIRECT):
    env = BatchedOvercookedEnv(["cramped_room"], n, horizon=25, io=io, auto_reset=True)
    check(env, acts_for(40, n), 8)
print("K1 x3 I/O strategies + K5 (S=16) ok", flush=True)
env = BatchedOvercookedEnv(["cramped_room", "counter_circuit", "asymmetric_advantages"], n, horizon=30, auto_reset=True)
check(env, acts_for(70, n), 5)
env = BatchedOvercookedEnv(["cramped_room", "counter_circuit"], n, horizon=30, auto_reset=False)
check(env, acts_for(50, n), 5)
print("K5 mixed layouts (S=32), with and without auto-reset ok", flush=True)
env = BatchedOvercookedEnv(["marshmallow_experiment"], 100, horizon=30, auto_reset=True)
check(env, acts_for(45, 100), 3)
print("K5 S=64 ok", flush=True)
env = BatchedOvercookedEnv(["cramped_room", "coordination_ring"], n, horizon=20, auto_reset=True, random_start_pos=True, rnd_obj_prob_thresh=0.6, seed=9)
rs = cpu.random_start(9, 0.6, True)
ref = env.state.cpu().numpy().copy()
cpu.reset_random(env._tab_host, env._starts_host, ref, rs, env_layout=env.env_layout_host)
ref[:, 3] = env.state.cpu().numpy()[:, 3]
check(env, acts_for(50, n), 4, rs=rs)
print("K4 random reset + K5 random-start auto-reset ok", flush=True)
# host-transfer formats through K5 and the native pipeline
env = BatchedOvercookedEnv(["cramped_room", "counter_circuit"], n, horizon=30, auto_reset=True)
a = acts_for(48, n)
ref_state = env.state.cpu().numpy().copy()
ref = cpu.rollout(env._tab_host, env._starts_host, ref_state, a, horizon=30, flags=1, n_threads=2)
for kw in ({"codes": True}, {"packed": True}, {"narrow": True}, {}):
    env.reset()
    pipe = HostRolloutPipeline(env, 48, chunk=16, **kw)
    ha = torch.from_numpy(wire.pack_actions(a)).pin_memory() if kw.get("codes") else torch.from_numpy(a.astype(np.uint8) if kw else a).pin_memory()
    out = pipe.run(ha)
    torch.cuda.synchronize()
    if kw.get("codes"):
        dense = env.expand_codes(out[3], events=True)
        for k, w in zip(("sparse", "shaped", "done", "events"), ref):
            assert np.array_equal(dense[k].numpy(), w)
    else:
        assert np.array_equal(out[0].numpy().astype(np.int32), ref[0])
    assert np.array_equal(env.state.cpu().numpy(), ref_state)
    pipe.close()
print("host-buffer pipeline, 4 transfer formats ok", flush=True)
# round-1 fused path (step_kernel with n_steps > 1) stays reachable: more than 8 layouts keep the tables in global memory
names = ["cramped_room", "coordination_ring", "forced_coordination", "five_by_five", "centre_pots", "centre_objects", "bottleneck",
         "simple_o", "scenario2", "scenario3"]
env = BatchedOvercookedEnv(names, n, horizon=30, auto_reset=True)
check(env, acts_for(40, n), 3)
print("fused step_kernel path (global tables) ok", flush=True)
# observation kernels
env = BatchedOvercookedEnv(["cramped_room"], n, horizon=400, auto_reset=True)
env.rollout(torch.from_numpy(acts_for(120, n)).cuda())
st = env.state.cpu().numpy()
for dt in (torch.float32, torch.uint8, torch.bfloat16, torch.int32):
    enc = env.lossless_state_encoding(dtype=dt)
    want = cpu.encode_lossless(env._tab_host, st, 5, 4, horizon=400)
    assert np.array_equal(enc.float().cpu().numpy(), want.astype(np.float32))
feat = env.featurize_state(num_pots=2).cpu().numpy()
lut = np.stack([l.feature_lut() for l in env.layouts]).view(np.uint8).reshape(1, -1)
assert np.array_equal(feat.astype(np.float64), cpu.featurize(env._tab_host, lut, st, 2))
phi = env.potential(0.99).cpu().numpy()
assert phi.shape == (n,)
env.reset(torch.from_numpy((rng.rand(n) < 0.5).astype(np.int32)).cuda())
torch.cuda.synchronize()
print("K2 x4 dtypes, K3, K6, K4 masked reset ok", flush=True)
# K7 (first policy layer from the record), K8 (dense tail + draw), the draw / return kernels, and the self-play transition on them
from overcooked_ai_b200.selfplay import SelfPlayRollout  # noqa: E402

env = BatchedOvercookedEnv(["cramped_room"], n, horizon=40, auto_reset=True)
env.rollout(torch.from_numpy(acts_for(25, n)).cuda())
wt = ((torch.rand((520, 128), device="cuda") - 0.5) * 0.1).to(torch.bfloat16)
bias = (torch.rand(128, device="cuda") - 0.5) * 0.2
obs = env.lossless_state_encoding(dtype=torch.float32).view(2 * n, 520)
want = torch.nn.functional.leaky_relu(obs @ wt.float() + bias, 0.2)
got = env.encoded_linear(wt, bias, neg_slope=0.2).float()
assert ((got - want).abs() <= want.abs() * 2.0 ** -8 + 1e-4).all()  # bf16 output of a float32 accumulation
sp = SelfPlayRollout(env, use_graph=False, seed=3)
assert sp.fused_first_layer and sp.fused_tail
ref_state = env.state.cpu().numpy().copy()
for t in range(6):
    sp.run(1)
    cpu.step(env._tab_host, env._starts_host, ref_state, sp.actions.cpu().numpy(), horizon=40, flags=1)
    assert np.array_equal(env.state.cpu().numpy(), ref_state)
sp2 = SelfPlayRollout(env, model=sp.model, use_graph=False, fused_tail=False)
sp2.run(3)
torch.cuda.synchronize()
print("K7, K8, sample / accumulate kernels, self-play transitions ok", flush=True)
print("sanitize_smoke: all ok")

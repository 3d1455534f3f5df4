"""Worker for test_two_ranks_mixed_batch_on_gpus (torch.distributed.run, NCCL, one rank per GPU).

BASELINE config 3's shape at test size: a mixed batch of the 5 classic layouts stored as 5 GLOBAL segments, sharded by
environment index with dist.shard_segments; each rank advances its shard with the fused rollout kernel, checks it
against the CPU oracle, and rank 0 checks the all-reduced counters against a single-process oracle run of the whole
batch.  No data-path collective: NCCL carries the seed and the counters only."""
import numpy as np
import torch

from oracle import cpu
from overcooked_ai_b200 import dist as D
from overcooked_ai_b200.batched import BatchedOvercookedEnv

CLASSIC5 = ["cramped_room", "asymmetric_advantages", "coordination_ring", "forced_coordination", "counter_circuit"]
rank, ws, local = D.init(backend="nccl")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
seed = D.broadcast_seed(424242 if rank == 0 else 0, device=dev)
assert seed == 424242
n_total, T, horizon = 20011, 90, 40
b, e = D.shard_range(n_total, rank, ws)
lay = D.shard_segments(n_total, 5, rank, ws)
assert len(lay) == e - b and (np.diff(lay) >= 0).all()
acts_all = np.random.RandomState(seed).randint(0, 6, size=(T, n_total, 2)).astype(np.int32)  # every rank draws the same trace
acts_all[np.random.RandomState(seed + 1).rand(T, n_total, 2) < 0.3] = 5
acts = np.ascontiguousarray(acts_all[:, b:e])
env = BatchedOvercookedEnv(CLASSIC5, e - b, horizon=horizon, device=dev, auto_reset=True, env_layout=lay)
ref_state = env.state.cpu().numpy().copy()
want = cpu.rollout(env._tab_host, env._starts_host, ref_state, acts, horizon=horizon, flags=1, n_threads=4)
got = env.rollout(torch.from_numpy(acts).to(dev))
for g, w in zip(got, want):
    assert np.array_equal(g.cpu().numpy(), w), "rank %d: shard mismatch" % rank
assert np.array_equal(env.state.cpu().numpy(), ref_state)
steps, ms, rew = D.reduce_counters(float((e - b) * T), 1.0 + rank, float(got[0].sum().item()), device=dev)
D.barrier()
if rank == 0:
    # the whole batch in one process on the CPU: what the two shards together must add up to
    from overcooked_ai_b200 import layout as L

    full_lay = D.shard_segments(n_total, 5, 0, 1)
    state = np.ascontiguousarray(env._starts_host[full_lay])
    whole = cpu.rollout(env._tab_host, env._starts_host, state, acts_all, horizon=horizon, flags=1, n_threads=8)
    assert steps == n_total * T and ms == float(ws) and rew == float(whole[0].sum()), (steps, ms, rew, whole[0].sum())
    print("DIST_GPU_OK")

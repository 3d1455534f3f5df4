"""Worker for test_sharding_world_size_2_gloo (launched by torch.distributed.run, gloo, CPU)."""
import numpy as np

from overcooked_ai_b200 import dist as D

rank, ws, local = D.init(backend="gloo")
assert ws == 2
n_total = 262144 + 3
b, e = D.shard_range(n_total, rank, ws)
sizes = [D.shard_range(n_total, r, ws) for r in range(ws)]
assert sizes[0][0] == 0 and sizes[-1][1] == n_total and all(sizes[i][1] == sizes[i + 1][0] for i in range(ws - 1))
seg = D.shard_segments(n_total, 5, rank, ws)
assert len(seg) == e - b and (np.diff(seg) >= 0).all() and seg.min() >= 0 and seg.max() <= 4
seed = D.broadcast_seed(1234567890123 if rank == 0 else 0)
assert seed == 1234567890123
steps, ms, rew = D.reduce_counters(100 * (rank + 1), 5.0 + rank, 7)
assert steps == 300 and ms == 6.0 and rew == 14
D.barrier()
if rank == 0:
    print("DIST_OK")

#!/usr/bin/env python
"""Back-to-back launches of the fused rollout kernel, as bench.py's timed region runs them (K launches, one event pair
around all of them), against single launches with a synchronise in between — by programmatic dependent launch on / off
and CTA tile.  One JSON line per measurement.    python tools/k5_backtoback.py [--n 65536] [--layouts cramped_room]"""
import argparse
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CHILD = r'''
import json, os, sys, time, torch
sys.path.insert(0, os.path.join(%(here)r, ".."))
from overcooked_ai_b200.batched import BatchedOvercookedEnv
layouts, n, T, pdl, K = %(layouts)r, %(n)d, 400, %(pdl)d, 10
env = BatchedOvercookedEnv(layouts, n, horizon=400, auto_reset=True, pdl=bool(pdl))
g = torch.Generator(device="cuda"); g.manual_seed(1)
acts = torch.randint(0, 6, (T, n, 2), dtype=torch.int32, device="cuda", generator=g)
out = env.alloc_rollout_out(T)
for _ in range(3):
    env.rollout(acts, out=out)
torch.cuda.synchronize()
single = []
for _ in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); env.rollout(acts, out=out); e1.record(); torch.cuda.synchronize()
    single.append(e0.elapsed_time(e1))
burst = []
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(K):
        env.rollout(acts, out=out)
    e1.record()
    t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    burst.append((e0.elapsed_time(e1) / K, t_enq / K * 1e3))
print(json.dumps({"tile": os.environ.get("OVC_K5_TILE", "default"), "lib": os.path.basename(os.environ.get("OVC_B200_LIB", "default")), "pdl": pdl,
                  "n_envs": n, "layouts": layouts, "single_launch_ms_best": round(min(single), 4),
                  "back_to_back_ms_per_launch": [round(b[0], 4) for b in burst], "host_enqueue_ms_per_launch": [round(b[1], 4) for b in burst]}), flush=True)
'''
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=65536)
ap.add_argument("--layouts", default="cramped_room")
ap.add_argument("--tiles", default="32,64")
ap.add_argument("--libs", default="")
args = ap.parse_args()
for lib in [""] + [x for x in args.libs.split(",") if x]:
    for tile in args.tiles.split(","):
        for pdl in (1, 0):
            env = dict(os.environ, OVC_K5_TILE=tile)
            env.pop("OVC_B200_LIB", None)
            if lib:
                env["OVC_B200_LIB"] = os.path.join(HERE, "..", "overcooked_ai_b200", "csrc", "libovc_b200_%s.so" % lib)
            r = subprocess.run([sys.executable, "-c", CHILD % {"here": HERE, "layouts": args.layouts.split(","), "n": args.n, "pdl": pdl}],
                               env=env, capture_output=True, text=True)
            sys.stdout.write(r.stdout if r.returncode == 0 else json.dumps({"error": r.stderr[-300:]}) + "\n")
            sys.stdout.flush()

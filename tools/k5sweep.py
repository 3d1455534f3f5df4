#!/usr/bin/env python
"""K5 (fused rollout) sweep on one GPU: the round-1 fused path (step_kernel with n_steps > 1, OVC_K5_LEGACY=1)
against the rollout kernel at each CTA tile size (OVC_K5_TILE), per batch size / layout set.  Each configuration
runs in its own process because the library reads the tuning variables once.  One JSON line per measurement.
    python tools/k5sweep.py [--sizes 65536,262144,1048576] [--layouts cramped_room] [--T 400]
"""
import argparse
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

CHILD = r'''
import json, os, sys, torch
sys.path.insert(0, os.path.join(%(here)r, ".."))
from overcooked_ai_b200.batched import BatchedOvercookedEnv
layouts, n, T, fmt = %(layouts)r, %(n)d, %(T)d, %(fmt)r
env = BatchedOvercookedEnv(layouts, n, horizon=400, auto_reset=True)
g = torch.Generator(device="cuda"); g.manual_seed(1)
acts = torch.randint(0, 6, (T, n, 2), dtype=torch.int32, device="cuda", generator=g)
if fmt == "codes":
    acts = (acts[..., 0] | (acts[..., 1] << 4)).to(torch.uint8).contiguous()
    out = env.alloc_rollout_out(T, codes=True)
else:
    out = env.alloc_rollout_out(T)
for _ in range(3):
    env.rollout(acts, out=out)
torch.cuda.synchronize()
best, tot, reps = 1e30, 0.0, 8
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); env.rollout(acts, out=out); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1); best = min(best, ms); tot += ms
chk = int(out[0].sum().item()) if out[0] is not None else int(out[3].to(torch.int64).sum().item())
print(json.dumps({"variant": os.environ.get("OVC_K5_LEGACY", "0") == "1" and "legacy step_kernel T-loop" or "rollout_kernel tile=" + os.environ.get("OVC_K5_TILE", "default") + (" lib=" + os.path.basename(os.environ["OVC_B200_LIB"]) if os.environ.get("OVC_B200_LIB") else ""),
                  "layouts": layouts, "n_envs": n, "T": T, "format": fmt, "S": env.state_words, "ms_best": round(best, 4), "ms_mean": round(tot / reps, 4),
                  "env_steps_per_s": n * T / (best * 1e-3), "checksum": chk}), flush=True)
'''

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="65536,262144,1048576")
ap.add_argument("--layouts", default="cramped_room")
ap.add_argument("--T", type=int, default=400)
ap.add_argument("--tiles", default="32,64,128")
ap.add_argument("--formats", default="int32")
ap.add_argument("--libs", default="", help="comma-separated experiment builds to run as extra variants")
args = ap.parse_args()
for fmt in args.formats.split(","):
    for n in [int(x) for x in args.sizes.split(",")]:
        T = args.T if n * args.T <= 2 ** 28 else max(50, 2 ** 28 // n)
        variants = [{"OVC_K5_LEGACY": "1"}] + [{"OVC_K5_TILE": t} for t in args.tiles.split(",")]
        for lib in args.libs.split(",") if args.libs else []:  # experiment builds (python -m overcooked_ai_b200.build --variant=x -D...)
            variants += [{"OVC_K5_TILE": t, "OVC_B200_LIB": os.path.join(HERE, "..", "overcooked_ai_b200", "csrc", "libovc_b200_%s.so" % lib)}
                         for t in args.tiles.split(",")]
        for v in variants:
            env = dict(os.environ)
            env.pop("OVC_K5_LEGACY", None), env.pop("OVC_K5_TILE", None), env.pop("OVC_B200_LIB", None)
            env.update(v)
            code = CHILD % {"here": HERE, "layouts": args.layouts.split(","), "n": n, "T": T, "fmt": fmt}
            r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
            sys.stdout.write(r.stdout)
            if r.returncode:
                sys.stdout.write(json.dumps({"variant": v, "n_envs": n, "error": r.stderr[-400:]}) + "\n")
            sys.stdout.flush()

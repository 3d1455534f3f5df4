#!/usr/bin/env python
"""bench.py — env-steps/sec of the batched Overcooked step engine on N B200s (BASELINE.json metric).

One bench "step" = one pass of the hot path over one batch of synthetic input = one full
400-transition horizon of the workload's environments (BASELINE.json configs[1]: cramped_room,
65 536 environments per GPU, random joint actions, 400 steps), i.e. 26 214 400 joint transitions
per GPU per step.  Environments shard by index across GPUs with no data-path collective
(weak scaling: per-GPU work is fixed); NCCL carries the run seed and the final counters.

  value     whole-job env-steps/s with the action trace already resident in HBM
            (mode "step": 400 launches of the step kernel K1, each through the C ABI ovc_step;
             mode "graph": the same 400 launches replayed from one CUDA graph;
             mode "rollout": one launch of the fused T-step kernel through ovc_rollout)
  e2e       the same metric through the public host-buffer API (HostRolloutPipeline): actions start
            in pinned HOST memory, rewards / done / events end in pinned HOST memory, every byte
            copied inside the timed region
  roofline  dominant kernel of the timed region vs the measured HBM copy peak (MEASURED_PEAKS.json)
  cpu_baseline   the CPU oracle (C restatement of the reference's transition, oracle/) on the host cores

--impl reference times that CPU restatement with all host threads on the same workload
(the reference itself is Python and is not present on the GPU box; its own Python step was
measured at 25.4 k steps/s/core in the build container).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (layouts, envs per GPU, horizon)
    "config2": (["cramped_room"], 65536, 400),
    "config3": (["cramped_room", "asymmetric_advantages", "coordination_ring", "forced_coordination", "counter_circuit"], 262144, 400),
    "config4": (["asymmetric_advantages"], 131072, 400),
    # config 5: PPO-style self-play rollout (K2 encode -> torch CNN -> multinomial -> K1 step), 262 144 envs on 8 GPUs
    "config5": (["cramped_room"], 32768, 400),
}
METRIC = "env-steps/sec (joint transitions)"
_REAL_STDOUT = None


def emit(line):
    """The one JSON line of this run, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def algorithmic_bytes_per_env_step(S):
    """read + write the record, 8 B actions in, 4+8+4+8 B outputs."""
    return 2 * 4 * S + 32


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(object):
    """nvidia-smi clock / throttle sampling DURING the timed region (B200_PROFILING.md recipe)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])), mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def oracle_tables(layout_names, n_envs):
    from overcooked_ai_b200 import layout as L

    layouts = [L.compile_layout(n) for n in layout_names]
    tab, starts, S = L.build_tables(layouts)
    bounds = [n_envs * i // len(layouts) for i in range(len(layouts) + 1)]
    state = np.concatenate([np.repeat(starts[i:i + 1], bounds[i + 1] - bounds[i], 0) for i in range(len(layouts))])
    return tab, starts, S, np.ascontiguousarray(state)


def cpu_run(layout_names, n_envs, T, horizon, threads, seed=0):
    """One bounded CPU sample: n_envs environments x T transitions through the oracle. Returns (steps, seconds)."""
    from oracle import cpu as oracle_cpu

    tab, starts, S, state = oracle_tables(layout_names, n_envs)
    rng = np.random.RandomState(seed)
    acts = rng.randint(0, 6, size=(T, n_envs, 2)).astype(np.int32)
    out = oracle_cpu.alloc_rollout_out(T, n_envs)  # pre-touched: the timed region is the transitions only
    t0 = time.perf_counter()
    oracle_cpu.rollout(tab, starts, state, acts, horizon=horizon, flags=1, n_threads=threads, out=out)
    return n_envs * T, time.perf_counter() - t0


def cpu_baseline(layout_names, horizon, budget_s=12.0):
    from oracle import cpu as oracle_cpu

    threads = oracle_cpu.max_threads()
    n_envs, T = 8192, 100
    cpu_run(layout_names, 1024, 20, horizon, threads)  # warm
    steps, sec, reps = 0, 0.0, 0
    t_start = time.perf_counter()
    while time.perf_counter() - t_start < budget_s and reps < 64:
        s, dt = cpu_run(layout_names, n_envs, T, horizon, threads, seed=reps)
        steps, sec, reps = steps + s, sec + dt, reps + 1
    return {
        "value": steps / sec, "unit": "env-steps/s", "cores": threads, "kind": "port",
        "sample": "%d x (%d envs x %d transitions) of %s through oracle/ovc_oracle.c, %d threads, rollout time only"
                  % (reps, n_envs, T, "+".join(layout_names), threads),
    }


def run_reference(args, rank, world):
    """--impl reference: the CPU restatement on all host threads, same workload, bounded sample per step."""
    if rank != 0:
        return
    from oracle import cpu as oracle_cpu

    layouts, n_envs, horizon = WORKLOADS[args.workload]
    threads = oracle_cpu.max_threads()
    sample_envs = min(n_envs, 65536 if threads >= 32 else 16384)
    T = horizon
    for _ in range(args.warmup):
        cpu_run(layouts, min(sample_envs, 2048), 50, horizon, threads)
    tot_steps, tot_sec = 0, 0.0
    for k in range(args.steps):
        s, dt = cpu_run(layouts, sample_envs, T, horizon, threads, seed=k)
        tot_steps, tot_sec = tot_steps + s, tot_sec + dt
    value = tot_steps / tot_sec
    S = oracle_tables(layouts, 8)[2]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_sec / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": "%s: %s, %d envs/GPU, %d-step horizon, uniform random joint actions" % (args.workload, "+".join(layouts), n_envs, horizon),
                   "state_words": S, "bounded_sample": "%d envs x %d transitions per step (CPU)" % (sample_envs, T)},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": threads, "kind": "port",
                         "sample": "%d steps x %d envs x %d transitions, oracle/ovc_oracle.c, %d threads" % (args.steps, sample_envs, T, threads)},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_policy_pipeline(args, rank, world, local):
    """BASELINE config 5: step + lossless_state_encoding feeding a random-init torch CNN policy."""
    import torch

    from overcooked_ai_b200 import dist as D
    from overcooked_ai_b200.batched import BatchedOvercookedEnv
    from overcooked_ai_b200.selfplay import SelfPlayRollout

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    D.init("nccl")
    seed = D.broadcast_seed(20260922, device=dev)
    torch.manual_seed(seed + rank)
    layouts, n_envs, horizon = WORKLOADS[args.workload]
    if args.envs:
        n_envs = args.envs
    T = horizon
    torch.backends.cudnn.benchmark = True
    env = BatchedOvercookedEnv(layouts, n_envs, horizon=horizon, device=dev, auto_reset=True)
    sp = SelfPlayRollout(env, use_graph=True)

    def timed(fn, k, w):
        for _ in range(w):
            fn(T // 4)
        D.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn(T)
        e1.record()
        torch.cuda.synchronize(dev)
        D.barrier()
        return e0.elapsed_time(e1)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(sp.run, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    ms_env = timed(sp.env_only, max(2, args.steps // 2), 1)
    steps_local = float(n_envs) * T * args.steps
    tot_steps, max_ms, tot_reward = D.reduce_counters(steps_local, ms, float(sp.ret_sparse.sum().item()), device=dev)
    _, max_ms_env, _ = D.reduce_counters(0, ms_env / max(2, args.steps // 2), 0, device=dev)
    if rank != 0:
        return
    S = env.state_words
    l = env.layouts[0]
    enc_bytes = 4 * S + 2 * l.width * l.height * 26 * sp.obs.element_size()
    line = {
        "metric": METRIC, "value": tot_steps / (max_ms * 1e-3), "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32 env / bf16 observations / bf16-autocast policy", "data": "synthetic",
        "config": {"workload": "config5: %s, %d envs/GPU, self-play: K2 lossless encode bf16 -> torch CNN (RllibPPOModel-shaped, random init, shared) -> multinomial -> K1 step, whole transition in one CUDA graph"
                               % ("+".join(layouts), n_envs), "state_words": S, "parallelism": "env-index sharding x%d" % world},
        "clocks": clocks, "gpu_launches": 2 * T * args.steps,
        "env_only": {"ms_per_400_transitions": max_ms_env, "env_steps_per_s_per_gpu": n_envs * T / (max_ms_env * 1e-3),
                     "share_of_pipeline_time": max_ms_env / (max_ms / args.steps),
                     "algorithmic_bytes_per_env_step": algorithmic_bytes_per_env_step(S) + enc_bytes},
        "sparse_reward_sum": tot_reward,
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="rollout", choices=["step", "graph", "rollout"])
    ap.add_argument("--io", type=int, default=0, help="record I/O strategy of K1 (0 default, 1 TMA tensor, 2 TMA bulk, 3 direct)")
    ap.add_argument("--envs", type=int, default=0, help="override environments per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--extra", action="store_true", help="also time the remaining modes")
    ap.add_argument("--no-pdl", action="store_true", help="disable programmatic dependent launch between K1 launches")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    # stdout must carry exactly one JSON line, but NCCL printf()s its version banner to fd 1 at communicator
    # creation: park the real stdout, point fd 1 at stderr for the run, emit the JSON line on the real one.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload == "config5":
        run_policy_pipeline(args, rank, world, local)
        return

    import torch

    from overcooked_ai_b200 import dist as D
    from overcooked_ai_b200.batched import BatchedOvercookedEnv, HostRolloutPipeline

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    D.init("nccl")
    seed = D.broadcast_seed(20260922, device=dev)

    layouts, n_envs, horizon = WORKLOADS[args.workload]
    if args.envs:
        n_envs = args.envs
    T = horizon
    env = BatchedOvercookedEnv(layouts, n_envs, horizon=horizon, device=dev, auto_reset=True, io=args.io, pdl=not args.no_pdl)
    S = env.state_words
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed + rank)
    actions = torch.randint(0, 6, (T, n_envs, 2), dtype=torch.int32, device=dev, generator=gen)
    out = (torch.empty((T, n_envs), dtype=torch.int32, device=dev), torch.empty((T, n_envs, 2), dtype=torch.int32, device=dev),
           torch.empty((T, n_envs), dtype=torch.int32, device=dev), torch.empty((T, n_envs, 2), dtype=torch.int32, device=dev))
    out_t = [tuple(o[t] for o in out) for t in range(T)]

    def pass_step():
        for t in range(T):
            env.step(actions[t], out=out_t[t])
        return T

    graph = None

    def pass_graph():
        graph.replay()
        return T

    def pass_rollout():
        env.rollout(actions, out=out)
        return 1

    def make_graph():
        s = torch.cuda.Stream(dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            pass_step()
        torch.cuda.current_stream(dev).wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            pass_step()
        return g

    graph = make_graph()  # always: the per-transition kernel K1 is reported next to the headline mode
    passes = {"step": pass_step, "graph": pass_graph, "rollout": pass_rollout}

    def timed(fn, k, w):
        env.reset()
        for _ in range(w):
            fn()
        D.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launches = 0
        for _ in range(k):
            launches += fn()
        e1.record()
        torch.cuda.synchronize(dev)
        D.barrier()
        return e0.elapsed_time(e1), launches

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, launches = timed(passes[args.mode], args.steps, args.warmup)

    steps_local = float(n_envs) * T * args.steps
    reward_local = float(out[0].sum().item())
    tot_steps, max_ms, tot_reward = D.reduce_counters(steps_local, ms, reward_local, device=dev)
    value = tot_steps / (max_ms * 1e-3)

    # ---- the per-transition kernel K1 (400 launches from one CUDA graph), measured in the same run ----
    k1_steps = max(2, min(args.steps, 5))
    if args.mode == "graph":
        k1_ms, k1_launches = ms, launches
        k1_steps = args.steps
    else:
        k1_ms, k1_launches = timed(pass_graph, k1_steps, 3)
    clocks = sampler.stop() if rank == 0 else None  # sampled (20 ms period) over the headline region and the K1 region

    # ---- e2e: the same workload through the public host-buffer API ----
    def run_e2e(fmt, expand=False):
        """One HostRolloutPipeline format.  expand: additionally rebuild the dense sparse / shaped / done arrays
        from the code words on the host cores inside the timed region (a consumer that wants arrays, not words)."""
        narrow = fmt != "int32"
        codes = fmt == "codes"
        # measured over several boxes (tools/e2e_probe.py): with 2-byte words the copies are short, and fewer, larger
        # chunks ride out the host's PCIe / memory noise best; the wider formats are plainly D2H bound at any size
        chunk = int(os.environ.get("OVC_E2E_CHUNK", "200" if codes else "50"))
        pipe = HostRolloutPipeline(env, T, chunk=chunk, narrow=narrow, packed=fmt == "packed", codes=codes, host_buffers=2)
        if codes:
            from overcooked_ai_b200 import wire
            h_actions = torch.from_numpy(wire.pack_actions(actions.cpu().numpy())).pin_memory()
            dense = {"sparse": torch.empty((T, n_envs), dtype=torch.int16), "shaped": torch.empty((T, n_envs, 2), dtype=torch.int8),
                     "done": torch.empty((T, n_envs), dtype=torch.uint8)}
        else:
            h_actions = torch.empty((T, n_envs, 2), dtype=pipe.act_dtype, pin_memory=True)
            h_actions.copy_(actions)
        env.reset()

        n_thr = max(1, min(64, (os.cpu_count() or 1) // world))  # 64 threads saturate the host expansion (tools/expand_bench.py)

        def passes_e2e(k):
            """k passes back to back, as a collection loop runs them: pass i+1 is submitted before pass i has
            drained (two pinned output sets), and with `expand` the host rebuilds pass i's arrays meanwhile."""
            prev = None
            for _ in range(k):
                cur_ = pipe.run(h_actions, wait=False)
                if expand and prev is not None:
                    prev[1].synchronize()
                    env.expand_codes(prev[0][3], out=dense, n_threads=n_thr)
                prev = cur_
            prev[1].synchronize()
            if expand:
                env.expand_codes(prev[0][3], out=dense, n_threads=n_thr)
            pipe.join()
            return prev[0]

        passes_e2e(2)
        torch.cuda.synchronize(dev)
        D.barrier()
        k_e2e = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        h_out = passes_e2e(k_e2e)
        torch.cuda.synchronize(dev)
        e2e_ms = (time.perf_counter() - t0) * 1e3
        D.barrier()
        _, e2e_max_ms, _ = D.reduce_counters(0, e2e_ms, 0, device=dev)
        sparse_host = env.expand_codes(h_out[3], shaped=False, done=False)["sparse"] if codes else h_out[0]
        what = {"codes": "both agents' event codes + done + reward-grant bits in ONE int16 per env-step (lossless: rewards are "
                         "table lookups of the codes, env.expand_codes / ovc_expand_codes_host)%s"
                         % (", expanded to dense int16 sparse / int8x2 shaped / uint8 done arrays on the host cores inside the timed region" if expand else ""),
                "packed": "sparse int16 + shaped int8x2 + both agents' event codes and done in one int16 (lossless, wire.decode_event_codes)",
                "narrow": "sparse int16 / shaped int8 / done uint8 / events int32", "int32": "sparse/shaped/done/events int32"}[fmt]
        return {"value": float(n_envs) * T * k_e2e * world / (e2e_max_ms * 1e-3), "unit": "env-steps/s",
                "h2d_bytes_per_step": pipe.h2d_bytes_per_step * T, "d2h_bytes_per_step": pipe.d2h_bytes_per_step * T,
                "steps": k_e2e, "ms_per_step": e2e_max_ms / k_e2e,
                "api": "overcooked_ai_b200.batched.HostRolloutPipeline(%s).run: pinned host actions (%s) in, pinned host %s out, "
                       "%d-transition chunks, H2D / fused rollout kernel / D2H on three streams, successive passes submitted "
                       "back to back (two pinned output sets; every pass's copies and its completion are inside the timed region)"
                       % (fmt, "one uint8 per joint action" if codes else "uint8" if narrow else "int32", what, chunk),
                "checksum_sparse": int(sparse_host.sum(dtype=torch.int64).item())}

    e2e = None
    if not args.no_e2e:
        if env.narrow_ok():
            e2e = run_e2e("codes")                       # 1 B in, 2 B out per env-step
            e2e["codes_expanded"] = run_e2e("codes", expand=True)  # + dense reward / done arrays rebuilt on the host
            e2e["packed_formats"] = run_e2e("packed")    # rewards as arrays, events as codes (2 B in, 6 B out)
            e2e["narrow_formats"] = run_e2e("narrow")    # the same pipeline with int32 event masks (13 B out)
            e2e["int32_formats"] = run_e2e("int32")      # and with 32-bit-everything formats (8 B in, 24 B out)
        else:
            e2e = run_e2e("int32")

    extra = {}
    if args.extra:
        for m in ("step", "rollout"):
            if m == args.mode:
                continue
            ms_m, l_m = timed(passes[m], max(3, args.steps // 2), 3)
            extra[m] = {"env_steps_per_s_per_gpu": float(n_envs) * T * max(3, args.steps // 2) / (ms_m * 1e-3), "launches": l_m}

    if rank != 0:
        return

    peak, peak_src = load_peaks()
    bytes_per = algorithmic_bytes_per_env_step(S)

    def roofline_of(mode, ms_, launches_):
        fused = mode == "rollout"
        per_launch_env_steps = n_envs * (T if fused else 1)
        avg_launch_s = ms_ * 1e-3 / launches_
        achieved = per_launch_env_steps * bytes_per / avg_launch_s / 1e9
        r = {
            "bound": "hbm",
            "kernel": "ovc::step_kernel<S=%d> %s" % (S, "K5: T=%d transitions fused in one launch, record tile resident in shared memory" % T if fused else "K1: one transition per launch"),
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
            "algorithmic_bytes_per_env_step": bytes_per, "env_steps_per_launch": per_launch_env_steps,
            "avg_launch_us": avg_launch_s * 1e6, "traffic": None,
            "note": "achieved = env_steps_per_launch x %d B (SURVEY 8d) / avg launch duration; avg launch duration = CUDA-event time of the timed region / launches (includes launch gaps)" % bytes_per,
        }
        # secondary bound: warp-instruction issue.  Instructions per warp-transition come from the
        # ncu captures under profiles/ (418 for K5, 445 for K1 on cramped_room); peak = SMs x 4 schedulers x SM clock.
        wi = {16: (418, 445)}.get(S)
        if wi and clocks and clocks.get("sm_mhz"):
            per_wt = wi[0] if fused else wi[1]
            issued = per_launch_env_steps / 32.0 * per_wt / avg_launch_s
            peak_issue = 148 * 4 * clocks["sm_mhz"] * 1e6
            r["issue_bound"] = {"warp_instructions_per_warp_transition": per_wt, "source": "profiles/r1_final_kernels_ncu_full.md",
                                "achieved_warp_inst_per_s": issued, "peak_warp_inst_per_s": peak_issue, "frac": issued / peak_issue}
        # DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed `ncu --set full`
        # captures of exactly this launch shape; any other shape has no capture and stays null
        if S == 16 and n_envs == 65536 and args.workload == "config2":
            cap = (213.935104e6 + 577.443584e6, "profiles/r1_final_k5_ncu_raw.csv") if fused and T == 400 else \
                  (4.74e6, "profiles/r1_final_k1_ncu_raw.csv") if not fused else None
            if cap:
                r["traffic"], r["traffic_source"] = cap
        if fused:
            r["streamed_GBps"] = per_launch_env_steps * 32 / avg_launch_s / 1e9
            r["note"] += "; the fused kernel keeps the record on chip between transitions, so only actions + outputs (32 B per env-step, streamed_GBps) cross HBM: a frac near or above 1 is traffic avoided by fusion, not bandwidth"
        return r

    roofline = roofline_of(args.mode, ms, launches)
    roofline_k1 = roofline_of("graph", k1_ms, k1_launches)
    roofline_k1["env_steps_per_s_per_gpu"] = float(n_envs) * T * k1_steps / (k1_ms * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": {
            "workload": "%s: %s, %d envs/GPU, %d-step horizon with auto-reset, uniform random joint actions" % (args.workload, "+".join(layouts), n_envs, horizon),
            "state_words": S, "mode": args.mode, "io": args.io, "parallelism": "env-index sharding x%d, no data-path collective" % world,
            "l2": "per bench step the action trace + outputs (%.0f MB) stream through HBM and exceed the 126 MB L2; the %.1f MB state tensor is the carried value and stays L2 resident"
                  % (n_envs * T * 32 / 1e6, n_envs * S * 4 / 1e6),
        },
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "roofline_k1": roofline_k1,
        "episode_sparse_reward_sum": tot_reward,
    }
    if extra:
        line["other_modes"] = extra
    if not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(layouts, horizon)
    emit(line)


if __name__ == "__main__":
    try:
        main()
    finally:
        try:
            import torch.distributed as _d

            if _d.is_available() and _d.is_initialized():
                _d.destroy_process_group()
        except Exception:
            pass

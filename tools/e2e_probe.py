"""Where does the host-buffer path spend its time?  For a few chunk sizes: env-steps/s of the `codes` pipeline,
the host time spent ENQUEUEING a pass (ovc_pipeline_run returns before anything has run) and the device-side
time of the same pass with the copies removed (the fused rollout kernel alone on the staged chunk)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from overcooked_ai_b200 import wire  # noqa: E402
from overcooked_ai_b200.batched import BatchedOvercookedEnv, HostRolloutPipeline  # noqa: E402


def main():
    n, T = 65536, 400
    fmt = sys.argv[1] if len(sys.argv) > 1 else "codes"
    chunks = [int(c) for c in sys.argv[2:]] or [25, 50, 100, 200]
    env = BatchedOvercookedEnv("cramped_room", n, horizon=400, auto_reset=True)
    rng = np.random.RandomState(0)
    acts = rng.randint(0, 6, size=(T, n, 2)).astype(np.int32)
    for chunk in chunks:
        pipe = HostRolloutPipeline(env, T, chunk=chunk, codes=fmt == "codes", packed=fmt == "packed", host_buffers=2)
        h = torch.from_numpy(wire.pack_actions(acts) if fmt == "codes" else acts.astype(np.uint8)).pin_memory()
        for _ in range(3):
            pipe.run(h)
        torch.cuda.synchronize()
        k = 10
        t0 = time.perf_counter()
        enq = 0.0
        for _ in range(k):
            e0 = time.perf_counter()
            out, ticket = pipe.run(h, wait=False)
            enq += time.perf_counter() - e0
        ticket.synchronize()
        pipe.join()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / k
        # the kernel alone on device-resident chunks of the same size
        d_act = pipe.d_act[0]
        d_out = pipe.d_out[0]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(T // chunk):
            env.rollout(d_act, out=d_out)
        e1.record()
        torch.cuda.synchronize()
        print("%s chunk %3d: %.3e env-steps/s  pass %.3f ms  host enqueue %.3f ms/pass  kernels alone %.3f ms/pass  D2H %.1f GB/s"
              % (fmt, chunk, n * T / dt, dt * 1e3, enq / k * 1e3, e0.elapsed_time(e1), pipe.d2h_bytes_per_step * T / dt / 1e9))
        pipe.close()


if __name__ == "__main__":
    main()

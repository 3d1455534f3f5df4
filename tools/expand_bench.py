"""Host-side expansion of OVC_F_OUT_CODES words (ovc_expand_codes_host): words/s against the number of worker
threads, on the workload of bench.py's e2e leg (400 x 65 536 words).  No GPU work."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from overcooked_ai_b200 import _native, wire  # noqa: E402
from overcooked_ai_b200 import layout as L  # noqa: E402


def main():
    lib = _native.lib()
    tbl = wire.code_reward_table([L.compile_layout("cramped_room")])
    T, N = 400, 65536
    rng = np.random.RandomState(0)
    w = (rng.randint(0, 32, (T, N)) | (rng.randint(0, 32, (T, N)) << 5)).astype(np.int16)
    sp, sh, dn = np.zeros((T, N), np.int16), np.zeros((T, N, 2), np.int8), np.zeros((T, N), np.uint8)
    threads = [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8, 16, 32, 64, 128]
    for thr in threads:
        for _ in range(2):
            lib.ovc_expand_codes_host(w.ctypes.data, T, N, 0, tbl.ctypes.data, 1, sp.ctypes.data, sh.ctypes.data, dn.ctypes.data, 0, thr)
        t0 = time.perf_counter()
        for _ in range(5):
            lib.ovc_expand_codes_host(w.ctypes.data, T, N, 0, tbl.ctypes.data, 1, sp.ctypes.data, sh.ctypes.data, dn.ctypes.data, 0, thr)
        dt = (time.perf_counter() - t0) / 5
        print("threads %3d: %7.2f ms  %.2e words/s  (%d cores online)" % (thr, dt * 1e3, T * N / dt, os.cpu_count()))
    want = wire.decode_codes(w[:3], tbl)
    assert np.array_equal(sp[:3], want[0]) and np.array_equal(sh[:3], want[1])


if __name__ == "__main__":
    main()

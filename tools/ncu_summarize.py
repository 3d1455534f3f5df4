#!/usr/bin/env python
"""Read `ncu --set full` captures (gpurun_out/*.ncu-rep) here, without a GPU: export the raw page to profiles/ as CSV,
print a markdown table of the figures the design discusses, and update profiles/ncu_summary.json (what bench.py reads for
`roofline.traffic` and `issue_bound`).

    python tools/ncu_summarize.py --tag r2 config2=gpurun_out/r2_prof_k5_config2.ncu-rep:65536x400 config2_k1=...:65536x1
Each argument is key=report:ENVSxSTEPS (environments and transitions per captured launch).
"""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
METRICS = [
    ("duration under ncu (us)", "gpu__time_duration.sum", 1.0),
    ("DRAM read (MB)", "dram__bytes_read.sum", None),
    ("DRAM written (MB)", "dram__bytes_write.sum", None),
    ("grid", "launch__grid_size", 1.0), ("block", "launch__block_size", 1.0), ("registers", "launch__registers_per_thread", 1.0),
    ("warps active (% of peak)", "sm__warps_active.avg.pct_of_peak_sustained_active", 1.0),
    ("warp instructions", "smsp__inst_executed.sum", 1.0),
    ("threads per instruction", "smsp__thread_inst_executed_per_inst_executed.ratio", 1.0),
    ("issue active (% of active cycles)", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1.0),
    ("stall wait / issue", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", 1.0),
    ("stall short_scoreboard / issue", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", 1.0),
    ("stall long_scoreboard / issue", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", 1.0),
    ("stall branch_resolving / issue", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", 1.0),
    ("stall no_instruction / issue", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", 1.0),
    ("stall not_selected / issue", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", 1.0),
    ("stall math_pipe_throttle / issue", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", 1.0),
    ("shared bank conflicts (count)", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 1.0),
]
TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    if out.returncode != 0:
        sys.exit("ncu -i %s failed: %s" % (rep, out.stderr[-300:]))
    return out.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="r2")
    ap.add_argument("items", nargs="+")
    args = ap.parse_args()
    summary_path = os.path.join(ROOT, "profiles", "ncu_summary.json")
    summary = json.load(open(summary_path)) if os.path.exists(summary_path) else {}
    cols = []
    for item in args.items:
        key, rest = item.split("=", 1)
        rep, shape = rest.rsplit(":", 1)
        n_envs, n_steps = [int(x) for x in shape.split("x")]
        text = raw_page(rep)
        raw_name = "%s_%s_ncu_raw.csv" % (args.tag, key)
        open(os.path.join(ROOT, "profiles", raw_name), "w").write(text)
        rows = list(csv.reader(io.StringIO(text)))
        hdr, units, vals = rows[0], rows[1], rows[-1]  # the last captured launch
        get = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
        col = {"key": key, "kernel": get.get("Kernel Name", ("?", ""))[0][:60]}
        for label, m, _ in METRICS:
            if m not in get:
                col[label] = None
                continue
            v, u = get[m]
            v = float(v.replace(",", ""))
            if m.startswith("dram__bytes"):
                v = v * TO_BYTES.get(u, 1.0) / 1e6
            if m == "gpu__time_duration.sum":
                v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
            col[label] = v
        wi = col["warp instructions"] / (n_envs / 32.0 * n_steps)
        dram = (col["DRAM read (MB)"] + col["DRAM written (MB)"]) * 1e6
        col["warp instructions per warp-transition"] = wi
        col["launch shape"] = "%d envs x %d transitions" % (n_envs, n_steps)
        cols.append(col)
        summary[key] = {"warp_inst_per_warp_transition": round(wi, 1), "dram_bytes_per_launch": dram,
                        "threads_per_instruction": col["threads per instruction"], "warps_active_pct": col["warps active (% of peak)"],
                        "issue_active_pct": col["issue active (% of active cycles)"], "duration_under_ncu_us": col["duration under ncu (us)"],
                        "launch_shape": col["launch shape"], "kernel": col["kernel"], "source": "profiles/" + raw_name}
    json.dump(summary, open(summary_path, "w"), indent=1, sort_keys=True)
    labels = ["kernel", "launch shape"] + [l for l, _, _ in METRICS] + ["warp instructions per warp-transition"]
    print("| metric | " + " | ".join(c["key"] for c in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for l in labels:
        cells = []
        for c in cols:
            v = c.get(l)
            cells.append("-" if v is None else ("%.4g" % v if isinstance(v, float) else str(v)))
        print("| %s | %s |" % (l, " | ".join(cells)))


if __name__ == "__main__":
    main()

"""Summarise an .ncu-rep (one kernel launch captured with --set full) into JSON + a launch list CSV
into a table.  Usage:
    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01_x_summary.json [flops_per_launch]
    python profiles/summarize_ncu.py --launches gpurun_out/launches.csv profiles/r01_x_launches.md
"""
import csv
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "lts__t_bytes.sum": "l2_bytes",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active": "dmma_pipe_pct_of_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct_of_elapsed",
    "sm__pipe_tensor_cycles_active.max.pct_of_peak_sustained_elapsed": "tensor_pipe_pct_of_elapsed_max_sm",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dyn_smem_per_block",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe_throttle",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_wavefront_pct",
}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def summarize(rep, out, flops=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")]}
        for h, u, v in zip(hdr, units, vals):
            if h in KEYS:
                try:
                    x = float(v.replace(",", ""))
                except ValueError:
                    continue
                d[KEYS[h]] = x * UNIT.get(u, 1.0) if (KEYS[h] in ("duration_us", "dram_read", "dram_write", "l2_bytes", "dyn_smem_per_block")) else x
        d["dram_bytes_per_launch"] = d.get("dram_read", 0.0) + d.get("dram_write", 0.0)
        if flops:
            d["flops_per_launch"] = flops
            d["tflops_under_ncu"] = flops / d["duration_us"] * 1e-6
        res.append(d)
    json.dump(res if len(res) > 1 else res[0], open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


def launches(csv_path, out):
    rows = [r for r in csv.reader(open(csv_path)) if r and r[0].isdigit()]
    # columns: ID, Process ID, Process Name, Host Name, Kernel Name, Context, Stream, Block Size, Grid Size, Device, CC, Section, Metric Name, Unit, Value
    agg = {}
    for r in rows:
        name = r[4].split("(")[0]
        unit, val = r[-2], float(r[-1].replace(",", ""))
        us = val * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3}.get(unit, 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"| kernel | launches | total us | share |\n|---|---|---|---|\n")
        for name, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name[:90]}` | {c} | {us:.1f} | {100 * us / tot:.1f}% |\n")
    print(open(out).read())


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        summarize(sys.argv[1], sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else None)

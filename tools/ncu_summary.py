#!/usr/bin/env python
"""Summarise ncu captures into profiles/ (run here, no GPU needed).
  python tools/ncu_summary.py full  gpurun_out/prof3.ncu-rep profiles/r01_loglik_full.json
  python tools/ncu_summary.py list  gpurun_out/launches.csv  profiles/r01_launches_summary.json"""
import csv
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.per_cycle_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.sum",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    result = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        entry = {"kernel": d.get("Kernel Name"), "metrics": {}, "stalls_per_issue": {}}
        for k in KEEP:
            if k in d:
                entry["metrics"][k] = f"{d[k]} {units[hdr.index(k)]}".strip()
        for k in hdr:
            if k.startswith(STALL) and k.endswith("_per_issue_active.ratio"):
                try:
                    v = float(d[k])
                except ValueError:
                    continue
                if v >= 0.01:
                    entry["stalls_per_issue"][k[len(STALL):-len("_per_issue_active.ratio")]] = round(v, 3)
        result.append(entry)
    json.dump({"source": rep, "command": "ncu --set full --clock-control none --import-source on", "launches": result},
              open(out, "w"), indent=1)
    print(json.dumps(result, indent=1)[:3000])


def launch_list(path, out):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = {}
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        name = r[ik].split("(")[0]
        t = float(r[iv].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    total = sum(a[1] for a in agg.values())
    summ = {k: {"launches": v[0], "total_ns": v[1], "mean_ns": v[1] / v[0], "share": v[1] / total}
            for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])}
    json.dump({"source": path, "command": "ncu --metrics gpu__time_duration.sum --clock-control none",
               "note": "per-launch times under ncu are serialised/cold-cache: compare shares, not absolutes",
               "kernels": summ}, open(out, "w"), indent=1)
    print(json.dumps(summ, indent=1))


if __name__ == "__main__":
    {"full": full, "list": launch_list}[sys.argv[1]](sys.argv[2], sys.argv[3])

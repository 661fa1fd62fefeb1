#!/usr/bin/env python
"""Extracts the counters bench.py's `roofline` object cites from an `ncu --set full` report of ONE render-kernel launch
and files them under profiles/:

    python tools/ncu_extract.py report.ncu-rep scene:WxH:aa:world rays_in_launch [out.json]

writes the selected raw metrics to out.json (default profiles/<report>.counters.json) and updates
profiles/traffic.json[scene:WxH:aa:world] = {"bytes": dram read + write of the launch, "ncu": {...}}."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__thread_inst_executed_pred_on_per_inst_executed.ratio",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "sm__maximum_warps_per_active_cycle_pct"]
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}


def main():
    rep, key, rays = sys.argv[1], sys.argv[2], int(sys.argv[3])
    out = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "profiles", os.path.basename(rep).replace(".ncu-rep", "") + ".counters.json")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    sel = {}
    for k in KEYS + ["Kernel Name"]:
        if k in d:
            v, u = d[k]
            try:
                sel[k] = {"value": float(v.replace(",", "")), "unit": u}
            except ValueError:
                sel[k] = {"value": v, "unit": u}

    def num(k, default=None):
        return sel[k]["value"] if k in sel and isinstance(sel[k]["value"], float) else default

    def dram(k):
        return num(k, 0.0) * UNIT_SCALE.get(sel.get(k, {}).get("unit", "byte"), 1.0)
    summary = {"source": os.path.relpath(out, ROOT), "workload": key, "rays_in_launch": rays,
               "kernel_ms": num("gpu__time_duration.sum", 0) / {"ns": 1e6, "nsecond": 1e6, "us": 1e3, "usecond": 1e3, "s": 1e-3, "second": 1e-3}.get(
                   sel.get("gpu__time_duration.sum", {}).get("unit"), 1.0),
               "issue_slot_util": num("sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
               "lanes_active": num("smsp__thread_inst_executed_pred_on_per_inst_executed.ratio"),
               "warp_inst_per_ray": num("smsp__inst_executed.sum", 0) / rays,
               "l1_wavefront_util": num("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
               "l1_hit_rate": num("l1tex__t_sector_hit_rate.pct"), "l2_hit_rate": num("lts__t_sector_hit_rate.pct"),
               "dram_read_bytes": dram("dram__bytes_read.sum"), "dram_write_bytes": dram("dram__bytes_write.sum")}
    json.dump({"summary": summary, "metrics": sel}, open(out, "w"), indent=1)
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(tpath))
    except Exception:
        t = {}
    t[key] = {"bytes": summary["dram_read_bytes"] + summary["dram_write_bytes"], "from": summary["source"],
              "ncu": {k: summary[k] for k in ("issue_slot_util", "lanes_active", "warp_inst_per_ray", "l1_wavefront_util", "source")}}
    json.dump(t, open(tpath, "w"), indent=1)
    print(json.dumps(summary, indent=1))


if __name__ == "__main__":
    main()

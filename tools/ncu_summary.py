#!/usr/bin/env python
"""Summarise an .ncu-rep (captured on the B200 box with `ncu --set full`) into a small JSON that is
committed under profiles/.  Runs here (no GPU): `ncu -i <rep> --page raw --csv`.

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_xx_name.json [--note "..."]
"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.sum", "sm__inst_executed.sum.per_cycle_active", "sm__inst_executed.sum.per_cycle_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[4] if len(sys.argv) > 4 and sys.argv[3] == "--note" else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        rec = {"kernel": d.get("Kernel Name"), "grid": d.get("Grid Size"), "block": d.get("Block Size")}
        for k in KEEP:
            if k in d and d[k] != "":
                try:
                    rec[k] = float(d[k].replace(",", ""))
                except ValueError:
                    rec[k] = d[k]
                rec.setdefault("_units", {})[k] = units[hdr.index(k)]
        stalls = {}
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and d[h] != "":
                stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(d[h])
        rec["stall_warps_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
        if "dram__bytes_read.sum" in rec:
            u = rec["_units"]
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rec["traffic_bytes"] = rec["dram__bytes_read.sum"] * scale[u["dram__bytes_read.sum"]] + \
                rec["dram__bytes_write.sum"] * scale[u["dram__bytes_write.sum"]]
        launches.append(rec)
    json.dump({"source": rep, "note": note, "how": "ncu --set full --clock-control none; summarised by tools/ncu_summary.py",
               "launches": launches}, open(out, "w"), indent=1)
    print(f"{out}: {len(launches)} launch(es)")


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Summarise an ncu report (.ncu-rep, `ncu --set full`) into the short text/JSON kept under profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [--json out.json] [--note "..."] > profiles/x.txt

Runs here (no GPU needed): it only reads the report through `ncu -i ... --page raw --csv`.
"""
import argparse
import csv
import io
import json
import subprocess

KEEP = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum",
    "l1tex__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
    "launch__block_size",
    "launch__grid_size",
    "launch__registers_per_thread",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum",
    "sm__inst_executed_pipe_lsu.sum",
    "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--json")
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        rec = dict(zip(hdr, r))
        print("== %s  (grid %s x block %s)" % (rec.get("Kernel Name", "?")[:100], rec.get("launch__grid_size"), rec.get("launch__block_size")))
        js = {"kernel": rec.get("Kernel Name", "?")}
        for k in KEEP:
            if k in rec and rec[k] != "":
                u = units[hdr.index(k)]
                print("%-86s %-16s %s" % (k, u, rec[k]))
                try:
                    v = float(rec[k].replace(",", ""))
                    js[k] = v * SCALE.get(u, 1.0) if "bytes" in k else v
                except ValueError:
                    pass
        if "dram__bytes_read.sum" in js:
            js["dram_bytes_per_launch"] = js["dram__bytes_read.sum"] + js["dram__bytes_write.sum"]
        js["source"] = a.report
        if a.json:
            json.dump(js, open(a.json, "w"), indent=1)
    if a.note:
        print("\n" + a.note)


if __name__ == "__main__":
    main()

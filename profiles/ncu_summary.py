#!/usr/bin/env python
"""Selected raw counters of the first kernel in an .ncu-rep as JSON (profiles/*.json are made with this).

    python profiles/ncu_summary.py gpurun_out/x.ncu-rep "command that produced it" > profiles/r01_x_ncu_raw.json
"""
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
        "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]


UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}


def current(argv):
    """--current out.json raw.json:reads [raw.json:reads ...]: the per-launch counters bench.py turns into `roofline.traffic`
    and `issue_slots`, keyed by kernel name, from *_ncu_raw.json summaries (made by this script from captures of
    `tools/run_kernel.py`, which prints the reads per launch), stamped with the hash of the kernel sources of the tree
    they are written in.  Bytes are bytes, durations nanoseconds."""
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import source_hash
    out_path, specs = argv[0], argv[1:]
    res = {"source_hash": source_hash(), "kernels": {}}
    for spec in specs:
        raw, reads = spec.rsplit(":", 1)
        with open(raw) as fh:
            doc = json.load(fh)
        for k in doc["kernels"]:
            name = k["name"].replace("void ", "").split("<")[0].split("(")[0].replace("rk::", "").strip()
            if name in res["kernels"]:
                continue

            def num(key):
                v = k.get(key)
                if not v or not isinstance(v["value"], (int, float)):
                    return None
                return float(v["value"]) * UNIT_SCALE.get(v["unit"], 1.0)
            res["kernels"][name] = {
                "reads": int(reads), "capture": os.path.basename(raw), "command": doc.get("_command", ""), "full_name": k["name"],
                "dram_bytes": (num("dram__bytes_read.sum") or 0.0) + (num("dram__bytes_write.sum") or 0.0),
                "warp_instructions": num("smsp__inst_executed.sum"),
                "duration_ns_under_ncu": num("gpu__time_duration.sum"),
                "issue_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "alu_pct": num("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                "fma_pct": num("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                "lsu_pct": num("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
                "lts_pct": num("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                "lts_hit_pct": num("lts__t_sector_hit_rate.pct"),
                "dram_pct": num("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                "warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
                "smem_wavefronts": num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
                "smem_bank_conflicts": num("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
                "registers": num("launch__registers_per_thread")}
    with open(out_path, "w") as fh:
        json.dump(res, fh, indent=1)
        fh.write("\n")


def main():
    if sys.argv[1] == "--current":
        return current(sys.argv[2:])
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = {"_report": rep, "_command": sys.argv[2] if len(sys.argv) > 2 else "", "kernels": []}
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(units, vals)))
        k = {"name": d.get("Kernel Name", ("", ""))[1]}
        for key in KEYS:
            if key in d:
                try:
                    k[key] = {"value": float(d[key][1].replace(",", "")), "unit": d[key][0]}
                except ValueError:
                    k[key] = {"value": d[key][1], "unit": d[key][0]}
        res["kernels"].append(k)
    json.dump(res, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""tools/ncu_summary.py <raw.csv from `ncu -i rep --page raw --csv`> <workload> <in-samples per launch> <algorithmic bytes per in-sample> <out.json>
Reduces one captured launch to the JSON summary kept under profiles/ (bench.py reads roofline.traffic from it)."""
import csv
import json
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "lts__t_sector_hit_rate.pct"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    raw, workload, n_in, alg, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), float(sys.argv[4]), sys.argv[5]
    rows = list(csv.reader(open(raw)))
    hdr, units, r = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    m = {}
    for k in KEEP:
        if k in col:
            m[k] = {"value": r[col[k]], "unit": units[col[k]]}
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            v = float(r[col[h]])
            if v >= 0.1:
                m[h] = {"value": r[col[h]], "unit": units[col[h]]}
    dram = sum(float(m[k]["value"]) * UNIT.get(m[k]["unit"], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    s = {"round": 2, "kernel": r[col["Kernel Name"]] if "Kernel Name" in col else "", "workload": workload,
         "command": "ncu --set full --clock-control none --import-source on -k regex:<kernel> -s 4 -c 1 python bench.py --workload %s --steps 3 --warmup 3 --no-cpu --no-e2e (tools/r2_ncu_all.sh)" % workload,
         "in_samples_per_launch": n_in, "dram_bytes_per_launch": dram, "dram_bytes_per_in_sample": dram / n_in,
         "algorithmic_bytes_per_in_sample": alg, "metrics": m}
    json.dump(s, open(out, "w"), indent=1)
    print(out, s["kernel"][:60], "%.3f B/in-sample" % (dram / n_in), m["gpu__time_duration.sum"])


if __name__ == "__main__":
    main()

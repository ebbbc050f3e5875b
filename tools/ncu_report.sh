#!/bin/bash
# tools/ncu_report.sh <report.ncu-rep> <mangled kernel> <top-level .cu> [line:label ...]  -- headline metrics + per-phase attribution
set -e
REP=$1; KERN=$2; TOP=$3; shift 3
TMP=$(mktemp -d)
ncu -i $REP --page raw --csv > $TMP/raw.csv
python3 - $TMP/raw.csv <<'PY'
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr,units,r=rows[0],rows[1],rows[2]
want=['gpu__time_duration.sum','sm__warps_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
 'smsp__issue_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread',
 'launch__shared_mem_per_block_dynamic','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor.sum',
 'smsp__inst_executed.sum','launch__grid_size','launch__block_size']
for h,u,v in zip(hdr,units,r):
    if h in want or h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio'):
        print("%-90s %-12s %s"%(h,u,v))
PY
ncu -i $REP --page source --csv --print-source sass > $TMP/src.csv 2>/dev/null
LIB=$(cd $(dirname $0)/.. && pwd)/r8brain-free-src_b200/libr8bgpu.so
HERE=$(cd $(dirname $0) && pwd)
(cd $TMP && cuobjdump -xelf all $LIB >/dev/null && for f in *.cubin; do nvdisasm -gi $f >> all.sass 2>/dev/null || true; done)
python3 $HERE/ncu_phase_report.py $TMP/src.csv $TMP/all.sass "$KERN" "$TOP" "$@"
rm -rf $TMP

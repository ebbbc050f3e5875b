# tools/final_run.sh -- reproduces the round-2 numbers quoted in DESIGN.md section 8 / profiles/ on one B200:
#   GPU tests, smoke, the default bench line (with e2e, CPU baseline, verification), every chain, the ncu summaries.
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py --steps 20 --warmup 4 2>/dev/null | tail -1 > gpurun_out/final_default.json
bash tools/r2_all.sh          # every BASELINE chain + the two decimating extras, device-resident, verified
bash tools/r2_final1.sh       # ncu --set full captures reduced on the box (tools/ncu_report.sh, tools/ncu_summary.py)

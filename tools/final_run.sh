set -x
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for w in cfg2_1024ch_44100_96000_r24 cfg3_1024ch_48000_44100_r24 cfg5_512ch_48000_47999_r24 cfg4_128ch_44100_2822400_r24_extfft cfg3b_1024ch_192000_44100_r24; do
  extra="--no-cpu"; if [ $w = cfg2_1024ch_44100_96000_r24 ]; then extra=""; fi
  timeout 200 python bench.py --workload $w $extra 2>/dev/null | tail -1 > gpurun_out/final_$w.json
  python -c "import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], d['ms_per_step'], d['value'], d['e2e']['value'] if d['e2e'] else None, d['roofline']['frac'], d['roofline']['path']['frac'])" gpurun_out/final_$w.json
done

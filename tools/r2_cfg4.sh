timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for w in cfg4_128ch_44100_2822400_r24_extfft; do
  timeout 200 python bench.py --workload $w --no-cpu --no-e2e 2>gpurun_out/all_$w.err | tail -1 > gpurun_out/all_$w.json
  python -c "import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d['ms_per_step'],4), round(d['value']), d['verified'], d['verification'].get('max_err_eps'), d['verification'].get('rms_err_eps'), round(d['roofline']['path']['frac'],4), d['roofline']['stage_ms_per_step'])" gpurun_out/all_$w.json || tail -3 gpurun_out/all_$w.err
done

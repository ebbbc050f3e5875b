for w in cfg2_1024ch_44100_96000_r24 cfg3_1024ch_48000_44100_r24 cfg5_512ch_48000_47999_r24 cfg4_128ch_44100_2822400_r24_extfft cfg3b_1024ch_192000_44100_r24 cfg3c_1024ch_2822400_44100_r24; do
  timeout 200 python bench.py --workload $w --no-cpu --no-e2e 2>gpurun_out/all_$w.err | tail -1 > gpurun_out/all_$w.json
  python -c "import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d['ms_per_step'],4), round(d['value']), d['verified'], d['verification'].get('max_err_eps'), d['verification'].get('rms_err_eps'), round(d['roofline']['path']['frac'],4), d['roofline']['stage_ms_per_step'])" gpurun_out/all_$w.json || tail -3 gpurun_out/all_$w.err
done

timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_random.py -m gpu -x -q 2>&1 | tail -3
for w in cfg5_512ch_48000_47999_r24; do
  for e in A=1 R8BGPU_FUSED_V1=1; do
  env $e timeout 200 python bench.py --workload $w --no-cpu --no-e2e 2>gpurun_out/c5.err | tail -1 > gpurun_out/c5.json
  python -c "import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[2], round(d['ms_per_step'],4), round(d['value']), d['verified'], d['verification'].get('max_err_eps'), d['verification'].get('rms_err_eps'), d['roofline']['stage_ms_per_step'])" gpurun_out/c5.json $e || tail -3 gpurun_out/c5.err
  done
done
R8BGPU_FUSED_V1=1 timeout 200 python bench.py --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg2 on v1 kernel', round(d['ms_per_step'],4), d['verified'])"

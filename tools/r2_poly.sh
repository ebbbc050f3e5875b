timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_random.py -m gpu -x -q 2>&1 | tail -6
for e in A=1 R8BGPU_POLY_V1=1; do
env $e timeout 200 python bench.py --workload cfg5_512ch_48000_47999_r24 --no-cpu --no-e2e 2>gpurun_out/c5.err | tail -1 > gpurun_out/c5.json
python -c "import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[2], round(d['ms_per_step'],4), round(d['value']), d['verified'], d['verification'], d['roofline']['stage_ms_per_step'])" gpurun_out/c5.json $e || tail -3 gpurun_out/c5.err
done

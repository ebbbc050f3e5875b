set -x
nvidia-smi -L
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15
timeout 200 python bench.py --no-cpu --no-e2e 2>gpurun_out/b_v2.err | tail -1 > gpurun_out/b_v2.json
python -c "import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], d['ms_per_step'], d['value'], d['roofline']['kernel'], d['roofline']['frac'])" gpurun_out/b_v2.json
R8BGPU_F2_FLAGS=0 timeout 200 python bench.py --no-cpu --no-e2e 2>/dev/null | tail -1 > gpurun_out/b_v2_f0.json
python -c "import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], d['ms_per_step'], d['value'], d['roofline']['kernel'], d['roofline']['frac'])" gpurun_out/b_v2_f0.json
R8BGPU_F2_FLAGS=1 timeout 200 python bench.py --no-cpu --no-e2e 2>/dev/null | tail -1 > gpurun_out/b_v2_f1.json
python -c "import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], d['ms_per_step'], d['value'], d['roofline']['kernel'], d['roofline']['frac'])" gpurun_out/b_v2_f1.json
R8BGPU_F2_FLAGS=2 timeout 200 python bench.py --no-cpu --no-e2e 2>/dev/null | tail -1 > gpurun_out/b_v2_f2.json
python -c "import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], d['ms_per_step'], d['value'], d['roofline']['kernel'], d['roofline']['frac'])" gpurun_out/b_v2_f2.json
R8BGPU_FUSED_V1=1 timeout 200 python bench.py --no-cpu --no-e2e 2>/dev/null | tail -1 > gpurun_out/b_v1.json
python -c "import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], d['ms_per_step'], d['value'], d['roofline']['kernel'], d['roofline']['frac'])" gpurun_out/b_v1.json
tail -5 gpurun_out/b_v2.err

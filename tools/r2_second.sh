set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
run() { name=$1; shift; env "$@" timeout 200 python bench.py --no-cpu --no-e2e 2>gpurun_out/$name.err | tail -1 > gpurun_out/$name.json; python -c "import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], d['ms_per_step'], d['value'], d['roofline']['kernel'], d['roofline']['frac'], d.get('verified'), d['verification'].get('max_err_eps'), d['roofline'].get('fp64'))" gpurun_out/$name.json || tail -5 gpurun_out/$name.err; }
run b_tc A=1
run b_tc_pp R8BGPU_F2_FLAGS=7
run b_tc_notma R8BGPU_F2_FLAGS=4
run b_fma R8BGPU_F2_FLAGS=2

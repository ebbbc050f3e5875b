nvidia-smi -L | head -8
nvidia-smi topo -m 2>/dev/null | head -12
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -6
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/multi_n$N.err | tail -1 > gpurun_out/multi_n$N.json
python -c "import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], d['n_gpus'], round(d['ms_per_step'],4), round(d['value']), d['verified'], 'e2e', d['e2e'], 'scatter', d['scatter_gather'])" gpurun_out/multi_n$N.json || tail -20 gpurun_out/multi_n$N.err

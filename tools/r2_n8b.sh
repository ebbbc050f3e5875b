N=$(nvidia-smi -L | wc -l)
for w in cfg3_1024ch_48000_44100_r24 cfg5_512ch_48000_47999_r24; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --workload $w --steps 10 --warmup 3 --no-e2e --no-scatter --no-cpu 2>gpurun_out/n${N}_$w.err | tail -1 > gpurun_out/n${N}_$w.json
  python -c "import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], d['n_gpus'], round(d['ms_per_step'],4), round(d['value']), d['verified'])" gpurun_out/n${N}_$w.json || tail -5 gpurun_out/n${N}_$w.err
done
w=cfg4_128ch_44100_2822400_r24_extfft
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --workload $w --steps 10 --warmup 3 --no-e2e --no-scatter --no-cpu 2>gpurun_out/n4_$w.err | tail -1 > gpurun_out/n4_$w.json
python -c "import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], d['n_gpus'], round(d['ms_per_step'],4), round(d['value']), d['verified'])" gpurun_out/n4_$w.json || tail -5 gpurun_out/n4_$w.err
R8BGPU_NO_HUGEPAGES=1 timeout 200 python tools/front_bench.py > gpurun_out/front_n${N}_nohuge.json 2>/dev/null; tail -1 gpurun_out/front_n${N}_nohuge.json | cut -c1-600

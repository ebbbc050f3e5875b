# quick GPU check: parity file + default bench + variants passed as "name ENV=.." args
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run() { name=$1; shift; env "$@" timeout 200 python bench.py --no-cpu --no-e2e 2>gpurun_out/$name.err | tail -1 > gpurun_out/$name.json; python -c "import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d['ms_per_step'],4), round(d['value']), d['roofline']['kernel'], round(d['roofline']['frac'],4), d.get('verified'), d['verification'].get('max_err_eps'), d['verification'].get('rms_err_eps'))" gpurun_out/$name.json || tail -5 gpurun_out/$name.err; }
run b_def A=1
for v in "$@"; do n=$(echo $v | tr -c 'A-Za-z0-9\n' '_'); run b_$n $v; done

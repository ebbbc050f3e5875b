run() { name=$1; shift; env "$@" timeout 200 python bench.py --no-cpu --no-e2e 2>gpurun_out/$name.err | tail -1 > gpurun_out/$name.json; python -c "import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d['ms_per_step'],4), d['verified'])" gpurun_out/$name.json || tail -3 gpurun_out/$name.err; }
run v_base A=1
run v_kunroll R8BGPU_LIB_PATH=/root/repo/r8brain-free-src_b200/libalt_A.so
run v_pair R8BGPU_LIB_PATH=/root/repo/r8brain-free-src_b200/libalt_B.so
run v_base2 A=1

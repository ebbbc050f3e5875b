run() { name=$1; shift; env "$@" timeout 200 python bench.py --workload cfg4_128ch_44100_2822400_r24_extfft --no-cpu --no-e2e 2>gpurun_out/$name.err | tail -1 > gpurun_out/$name.json; python -c "import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d['ms_per_step'],4), d['verified'], d['roofline']['stage_ms_per_step'])" gpurun_out/$name.json || tail -3 gpurun_out/$name.err; }
run hb_base A=1
run hb_b7000 R8BGPU_HB_SMEM_DOUBLES=7000
run hb128_b7000 R8BGPU_LIB_PATH=/root/repo/r8brain-free-src_b200/libr8bgpu_hb128.so R8BGPU_HB_SMEM_DOUBLES=7000
run hb128_b5000 R8BGPU_LIB_PATH=/root/repo/r8brain-free-src_b200/libr8bgpu_hb128.so R8BGPU_HB_SMEM_DOUBLES=5000
run hb128_b14336 R8BGPU_LIB_PATH=/root/repo/r8brain-free-src_b200/libr8bgpu_hb128.so
run hb_nolast2 R8BGPU_HB_NO_LAST2=1

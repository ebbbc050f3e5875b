run() { name=$1; wl=$2; shift 2; env "$@" timeout 200 python bench.py --workload $wl --no-cpu --no-e2e 2>gpurun_out/$name.err | tail -1 > gpurun_out/$name.json; python -c "import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d['ms_per_step'],4), d['verified'], d['roofline']['stage_ms_per_step'])" gpurun_out/$name.json || tail -3 gpurun_out/$name.err; }
C2=cfg2_1024ch_44100_96000_r24; C3=cfg3_1024ch_48000_44100_r24; C4=cfg4_128ch_44100_2822400_r24_extfft; C3C=cfg3c_1024ch_2822400_44100_r24
run v_c2_base $C2 A=1
run v_c2_kunroll $C2 R8BGPU_LIB_PATH=/root/repo/r8brain-free-src_b200/libalt_A.so
run v_c3_kunroll $C3 R8BGPU_LIB_PATH=/root/repo/r8brain-free-src_b200/libalt_A.so
run v_c4_b9000 $C4 R8BGPU_HB_SMEM_DOUBLES=9000
run v_c4_b5500 $C4 R8BGPU_HB_SMEM_DOUBLES=5500
run v_c3c_6400 $C3C R8BGPU_HBD_SMEM_DOUBLES=6400
run v_c3c_3200 $C3C R8BGPU_HBD_SMEM_DOUBLES=3200
run v_c3c_25000 $C3C R8BGPU_HBD_SMEM_DOUBLES=25000

B="--steps 3 --warmup 3 --no-cpu --no-e2e"
ncu --set full --clock-control none --import-source on -k regex:k_up2_frac2 -s 4 -c 1 -o /tmp/c5 python bench.py --workload cfg5_512ch_48000_47999_r24 $B > /dev/null 2>&1
tools/ncu_report.sh /tmp/c5.ncu-rep _ZN6r8bgpu11k_up2_frac2ILi8ELb0ELi0ELb1ELi2ELb0ELb1EEEvNS_11FusedParamsENS_7SrcViewENS_7DstViewE r8b_fused2.cu > gpurun_out/r2_ncu_cfg5_poly.txt 2>&1
tail -70 gpurun_out/r2_ncu_cfg5_poly.txt | cut -c1-150

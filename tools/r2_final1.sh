timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
B="--steps 3 --warmup 3 --no-cpu --no-e2e"
cap() { name=$1; kre=$2; wl=$3; mang=$4; top=$5; shift 5
  ncu --set full --clock-control none --import-source on -k regex:$kre -s 4 -c 1 -o /tmp/$name python bench.py --workload $wl $B > /dev/null 2>&1
  tools/ncu_report.sh /tmp/$name.ncu-rep "$mang" "$top" "$@" > gpurun_out/r2_ncu_$name.txt 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/r2_ncu_$name.raw.csv
  rm -f /tmp/$name.ncu-rep; }
F2=_ZN6r8bgpu11k_up2_frac2ILi8ELb0ELi0ELb1ELi2ELb0EEEvNS_11FusedParamsENS_7SrcViewENS_7DstViewE
LBL="148:setup 200:loop_head 211:A_gather_pass1 225:A_prepare+bar 227:B_fwd_passes 235:bar_B 236:C+inv_pass1 252:D_inverse 277:bar_D 278:E_interp 369:bar_E+tail"
cap cfg2 k_up2_frac2 cfg2_1024ch_44100_96000_r24 $F2 r8b_fused2.cu $LBL
cap cfg3c_cascade k_hbdown_cascade cfg3c_1024ch_2822400_44100_r24 _ZN6r8bgpu16k_hbdown_cascadeENS_16HbDownCascParamsENS_7SrcViewENS_7DstViewE r8b_kernels.cu
cap cfg3b_hbdown k_hbdown cfg3b_1024ch_192000_44100_r24 _ZN6r8bgpu8k_hbdownENS_8HbParamsENS_7SrcViewENS_7DstViewE r8b_kernels.cu
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2_launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
bash tools/r2_all.sh

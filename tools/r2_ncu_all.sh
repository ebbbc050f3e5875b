# ncu captures of the main kernels; each report is reduced ON THE BOX to a raw-metrics CSV line set + a per-phase table
# (tools/ncu_report.sh) and deleted, so that gpurun_out/ stays small.
B="--steps 3 --warmup 3 --no-cpu --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_launches_cfg2.log 2>&1
cap() { # name kernel-regex workload mangled top-file labels...
  name=$1; kre=$2; wl=$3; mang=$4; top=$5; shift 5
  ncu --set full --clock-control none --import-source on -k regex:$kre -s 4 -c 1 -o /tmp/$name python bench.py --workload $wl $B > /dev/null 2>&1
  tools/ncu_report.sh /tmp/$name.ncu-rep "$mang" "$top" "$@" > gpurun_out/r2_ncu_$name.txt 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/r2_ncu_$name.raw.csv
  rm -f /tmp/$name.ncu-rep
}
F2=_ZN6r8bgpu11k_up2_frac2ILi8ELb0ELi0ELb1ELi2ELb0EEEvNS_11FusedParamsENS_7SrcViewENS_7DstViewE
F2P=_ZN6r8bgpu11k_up2_frac2ILi8ELb1ELi0ELb1ELi2ELb0EEEvNS_11FusedParamsENS_7SrcViewENS_7DstViewE
F1P=_ZN6r8bgpu11k_up2_frac2ILi8ELb1ELi0ELb1ELi1ELb0EEEvNS_11FusedParamsENS_7SrcViewENS_7DstViewE
LBL="140:setup 192:loop_head 203:A_gather_pass1 217:A_prepare+bar 219:B_fwd_passes 225:bar_B 226:C_split_mul 242:bar_C 243:D_inverse 269:bar_D 270:E_interp 338:bar_E+tail"
cap cfg2 k_up2_frac2 cfg2_1024ch_44100_96000_r24 $F2 r8b_fused2.cu $LBL
cap cfg3 k_up2_frac2 cfg3_1024ch_48000_44100_r24 $F2P r8b_fused2.cu $LBL
cap cfg3b_up1 k_up2_frac2 cfg3b_1024ch_192000_44100_r24 $F1P r8b_fused2.cu $LBL
cap cfg3b_hbdown k_hbdown cfg3b_1024ch_192000_44100_r24 _ZN6r8bgpu8k_hbdownENS_8HbParamsENS_7SrcViewENS_7DstViewE r8b_kernels.cu
cap cfg3c_cascade k_hbdown_cascade cfg3c_1024ch_2822400_44100_r24 _ZN6r8bgpu16k_hbdown_cascadeENS_16HbDownCascParamsENS_7SrcViewENS_7DstViewE r8b_kernels.cu
cap cfg4_cascade k_hbup_cascade cfg4_128ch_44100_2822400_r24_extfft _ZN6r8bgpu14k_hbup_cascadeENS_15HbCascadeParamsENS_7SrcViewENS_7DstViewE r8b_kernels.cu
cap cfg5 "k_up2_frac<" cfg5_512ch_48000_47999_r24 none r8b_fused.cu
ls -la gpurun_out/ | head -40
bash tools/r2_all.sh

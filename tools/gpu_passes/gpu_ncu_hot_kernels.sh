cd /root/repo
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none -k regex:'conv_nhwc_bf16_kernel|wgrad_kernel|split_bf16_cm_kernel|demod_act_bwd_kernel|modulate_bwd_kernel|upfirdn2d_cminor4_kernel|upfirdn2d_wminor4_kernel|fir4_tma_kernel|fir4_up2_tma_kernel|bias_act_vec_kernel|adam_ema_kernel|raymarch_fwd2_kernel|raymarch_bwd2_kernel' -c 32 -o /tmp/r1_hot_kernels -f python tools/ncu_kernels.py > gpurun_out/ncu_hot.log 2>&1
tail -2 gpurun_out/ncu_hot.log
ncu -i /tmp/r1_hot_kernels.ncu-rep --page raw --csv > gpurun_out/r1_hot_kernels_raw.csv 2>/dev/null
ncu -i /tmp/r1_hot_kernels.ncu-rep --page details > gpurun_out/r1_hot_kernels_details.txt 2>/dev/null
ls -la gpurun_out/

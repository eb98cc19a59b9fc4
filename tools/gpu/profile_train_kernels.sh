# ncu --set full of the round-2 training kernels at the config-3 sizes -> gpurun_out/prof_train_kernels.ncu-rep
set -x
timeout 800 ncu --set full --clock-control none -k regex:'pa_fwd_tc|pa_bwd|mha_bwd|mha_tc|groupnorm_stats|groupnorm_apply|conv_dgrad_stem|bilinear_bwd_gather|bilinear_nhwc_v4' --launch-skip 11 --launch-count 11 -o gpurun_out/prof_train_kernels -f python tools/diag_train_kernels_once.py > gpurun_out/prof_train_kernels.log 2>&1
echo "rc $?"; tail -3 gpurun_out/prof_train_kernels.log

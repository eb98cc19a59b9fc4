set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'chain_qkvattn2' --launch-skip 4 --launch-count 1 -o gpurun_out/r2n_qkvattn3 -f python tools/diag_decoder.py 249615 --once > gpurun_out/r2n_ncu.log 2>&1
echo rc $?

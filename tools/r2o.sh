set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
echo "bench rc $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'chain|point_proj' --launch-skip 12 --launch-count 6 -o gpurun_out/r2o_decoder_full -f python tools/diag_decoder.py 2146689 --once > gpurun_out/r2o_ncu_full.log 2>&1
echo "ncu full rc $?"

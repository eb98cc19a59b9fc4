set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
echo "bench rc $?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2k_launches.csv python bench.py --steps 1 --warmup 3 --shapes 1 --no-cpu-baseline --no-e2e --no-shard --profile-region > gpurun_out/r2k_launch_bench.log 2>&1
echo "launch list rc $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'chain|point_proj' --launch-skip 18 --launch-count 9 -o gpurun_out/r2k_decoder_full -f python tools/diag_decoder.py 2146689 --once > gpurun_out/r2k_ncu_full.log 2>&1
echo "ncu full rc $?"
ls -la gpurun_out/

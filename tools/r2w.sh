set -x
timeout 1400 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r2w_pytest.log
tail -3 gpurun_out/r2w_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err
echo "bench rc $?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2w_smoke.log 2>&1
echo "smoke rc $?"

# Full GPU verification: the -m gpu test suite, smoke(), the default bench line and the reference arm.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- "bash tools/gpu/verify.sh"   -> gpurun_out/verify_*
set -x
timeout 1400 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/verify_pytest.log
tail -3 gpurun_out/verify_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/verify_smoke.log 2>&1
echo "smoke rc $?"
python bench.py --steps 5 --warmup 3 > gpurun_out/verify_bench.json 2> gpurun_out/verify_bench.err
echo "bench rc $?"
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/verify_bench_reference.json 2> gpurun_out/verify_bench_reference.err
echo "reference arm rc $?"

set -x
timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_bench_config.py -x -q -m gpu -k "occ or bench_config or slabs" -s 2>&1 | grep -v "^$" | tail -30 > gpurun_out/r2t_pytest.log
timeout 300 python tools/diag_decoder.py > gpurun_out/r2t_diag.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2t_bench_n2.json 2> gpurun_out/r2t_bench_n2.err
echo "bench n2 rc $?"
tail -3 gpurun_out/r2t_pytest.log

timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "mha" -s 2>&1 | tail -15 > gpurun_out/r2q_pytest.log
ZS_CHAIN_DBG=1 timeout 300 python tools/diag_decoder.py > gpurun_out/r2q_diag_dbg1.log 2>&1
ZS_CHAIN_DBG=9 timeout 300 python tools/diag_decoder.py > gpurun_out/r2q_diag_dbg9.log 2>&1
ZS_CHAIN_DBG=9 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'chain_pmlp|chain_qkvattn2' --launch-skip 4 --launch-count 2 --csv --log-file gpurun_out/r2q_dbg9.csv python tools/diag_decoder.py 2146689 --once > /dev/null 2>&1
tail -3 gpurun_out/r2q_pytest.log

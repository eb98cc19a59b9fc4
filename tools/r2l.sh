set -x
timeout 900 python -m pytest tests/test_gpu_chain.py -x -q -m gpu -k "qkvattn or points_mode or attention_variants" -s 2>&1 | tail -30 > gpurun_out/r2l_pytest.log
timeout 300 python tools/diag_decoder.py > gpurun_out/r2l_diag.log 2>&1
timeout 300 python tools/trace_chain.py qkvattn > gpurun_out/r2l_trace.log 2>&1
tail -5 gpurun_out/r2l_pytest.log

set -x
timeout 900 python -m pytest tests/test_gpu_chain.py -x -q -m gpu -k "pmlp or qkvattn or points_mode or attention_variants" -s 2>&1 | tail -40 > gpurun_out/r2l_pytest.log
timeout 300 python tools/diag_decoder.py > gpurun_out/r2l_diag.log 2>&1
tail -5 gpurun_out/r2l_pytest.log

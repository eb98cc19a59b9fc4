timeout 600 python -m pytest tests/test_gpu_chain.py tests/test_gpu_bench_config.py -x -q -m gpu -k "occ or bench_config" -s 2>&1 | grep "parity\|passed\|failed\|Error" | tail -12
ZS_CHAIN_DBG=1 timeout 300 python tools/diag_decoder.py 2>&1 | grep -A4 "^==" | grep "occ\|pass" | head -4

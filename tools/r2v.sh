timeout 600 python -m pytest tests/test_gpu_chain.py tests/test_gpu_bench_config.py -x -q -m gpu -k "pmlp or points_mode or bench_config or attention_variants" -s 2>&1 | grep "parity\|passed\|failed\|Error\|pmlp" | tail -14
ZS_CHAIN_DBG=1 timeout 300 python tools/diag_decoder.py 2>&1 | grep -A5 "^==" | grep "pmlp\|pass" | head -8

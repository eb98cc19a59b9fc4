timeout 900 python -m pytest tests/test_gpu_preprocess.py -x -q -m gpu 2>&1 | tail -30 > gpurun_out/r2l_pytest.log
tail -5 gpurun_out/r2l_pytest.log

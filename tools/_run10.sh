python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_fullsize.py -x -q -m gpu -k "not cfg4" 2>&1 | tail -5
python tools/gpu_cfg_pass.py cfg1 1048576 3
python tools/gpu_cfg_pass.py cfg3 100000000 2

python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q -m gpu 2>&1 | tail -8
python tools/gpu_cfg_pass.py cfg1 1048576 4

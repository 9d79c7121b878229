python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q -m gpu 2>&1 | tail -4
python tools/gpu_cfg_pass.py cfg2 211957760 4

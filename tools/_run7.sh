python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q -m gpu 2>&1 | tail -5
python tools/gpu_cfg_pass.py cfg3 100000000 2
python tools/gpu_cfg_pass.py cfg3 100000000 1 --fixed

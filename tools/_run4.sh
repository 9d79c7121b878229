python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q -m gpu -k "ROLZ or rolz or stream or fuzz" 2>&1 | tail -5
python tools/gpu_cfg_pass.py cfg5 268435456 2
python tools/gpu_cfg_pass.py cfg5 2147483648 1

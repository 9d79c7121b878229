python -m pytest tests -m gpu -q > gpurun_out/r02_t40.log 2>&1; tail -3 gpurun_out/r02_t40.log
python bench.py > gpurun_out/r02_b40.json 2> gpurun_out/r02_b40.err; tail -c 300 gpurun_out/r02_b40.err
python bench.py --impl reference > gpurun_out/r02_b40_ref.json 2> gpurun_out/r02_b40_ref.err

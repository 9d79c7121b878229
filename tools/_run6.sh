python -m pytest tests -m gpu -q > gpurun_out/r02_t20.log 2>&1; tail -3 gpurun_out/r02_t20.log
python bench.py > gpurun_out/r02_b20.json 2> gpurun_out/r02_b20.err; tail -c 600 gpurun_out/r02_b20.err
python bench.py --impl reference > gpurun_out/r02_b20_ref.json 2> gpurun_out/r02_b20_ref.err
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv"
$NCU --log-file gpurun_out/r02_launches_cfg2_groups.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-others --no-strong > gpurun_out/r02_launches_cfg2_groups.log 2>&1
KZG_LZ_GROUPS=1 KZG_DEC_GROUPS=1 $NCU --log-file gpurun_out/r02_launches_cfg2_groups1.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-others --no-strong > gpurun_out/r02_launches_cfg2_groups1.log 2>&1

SEC="--section SourceCounters --section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section SpeedOfLight --section InstructionStats"
ncu $SEC --clock-control none --import-source on -k regex:'ans1_|sbrt_inv' -c 3 -o gpurun_out/r02_prof_cfg3 python tools/gpu_cfg_pass.py cfg3 16777216 1 > gpurun_out/r02_prof_cfg3.log 2>&1
ncu $SEC --clock-control none --import-source on -k regex:'fpaq_|srt_inverse' -c 3 -o gpurun_out/r02_prof_cfg4 python tools/gpu_cfg_pass.py cfg4 2097152 1 > gpurun_out/r02_prof_cfg4.log 2>&1
ncu $SEC --clock-control none --import-source on -k regex:'rolz_parse|rolz_replay' -c 2 -o gpurun_out/r02_prof_cfg5 python tools/gpu_cfg_pass.py cfg5 16777216 1 > gpurun_out/r02_prof_cfg5.log 2>&1
ncu $SEC --clock-control none --import-source on -k regex:'huff_' -c 3 -o gpurun_out/r02_prof_cfg1 python tools/gpu_cfg_pass.py cfg1 1048576 1 > gpurun_out/r02_prof_cfg1.log 2>&1
tail -2 gpurun_out/r02_prof_cfg*.log

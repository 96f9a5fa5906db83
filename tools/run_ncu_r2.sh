# round-2 ncu evidence (single GPU): launch list of the default bench command, full capture of the FP8 render kernel
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python tools/tc_trace.py fp8 > gpurun_out/plain_fp8.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fused_render_fp8 -s 1 -c 1 -o gpurun_out/r2_fp8_full python tools/tc_trace.py fp8 > gpurun_out/ncu_fp8.log 2>&1

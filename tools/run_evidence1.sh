# single-GPU evidence: DSMEM probe, hierarchical workload, standalone HBM kernels (plain + ncu)
tools/probe/dsmem_probe > gpurun_out/dsmem_probe.txt 2>&1
python bench.py --workload hierarchical --steps 8 --warmup 3 > gpurun_out/hier_n1.json 2> gpurun_out/hier_n1.err
python tools/bench_geometry.py > gpurun_out/geom.json 2> gpurun_out/geom.err &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/geom_ncu.csv python tools/bench_geometry.py > gpurun_out/geom_ncu.log 2>&1

N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus $N --workload hierarchical --steps 16 --warmup 4 > gpurun_out/hier_n$N.json 2> gpurun_out/hier_n$N.err
$TR --master-port 29522 bench.py --gpus $N --precision fp8 --no-extras --steps 20 --warmup 5 > gpurun_out/bench_fp8_n$N.json 2> gpurun_out/bench_fp8_n$N.err
$TR --master-port 29523 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err

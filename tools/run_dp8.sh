# 8-GPU evidence run (gpurun --gpus 8): exchange check with both transports, then the default bench (render + extras.train)
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 tools/dp_check.py p2p > gpurun_out/dp_p2p_n$N.json 2> gpurun_out/dp_p2p_n$N.err
$TR --master-port 29512 tools/dp_check.py multimem > gpurun_out/dp_mm_n$N.json 2> gpurun_out/dp_mm_n$N.err
$TR --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
$TR --master-port 29514 bench.py --gpus $N --workload train --unfused --steps 40 --warmup 8 > gpurun_out/train_unfused_n$N.json 2> gpurun_out/train_unfused_n$N.err

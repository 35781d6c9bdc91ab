#!/bin/bash
mkdir -p gpurun_out/r2z
O=gpurun_out/r2z
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 100 --warmup 10 --no-cpu-baseline"
timeout 240 $TR > $O/n4.json 2> $O/n4.err; echo "n4 exit $?"; cut -c1-250 $O/n4.json
timeout 240 $TR --global-batch 8 > $O/n4_strong8.json 2> $O/n4_strong8.err; echo "n4 strong exit $?"; cut -c1-250 $O/n4_strong8.json
echo done

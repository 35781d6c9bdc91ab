#!/bin/bash
mkdir -p gpurun_out/r2z
O=gpurun_out/r2z
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 100 --warmup 10 --no-cpu-baseline > $O/n8.json 2> $O/n8.err
echo "n8 exit $?"; grep -i "warn\|error" $O/n8.err | head -3; cut -c1-250 $O/n8.json
echo done

#!/bin/bash
# 2 GPUs: data-parallel parity + scaling probes
mkdir -p gpurun_out/r2i
O=gpurun_out/r2i
nvidia-smi -L > $O/gpus.txt
timeout 900 python -m pytest tests/test_dp_gpu.py -m gpu -q --tb=short -p no:cacheprovider > $O/pytest_dp.log 2>&1
echo "exit $?" >> $O/pytest_dp.log; tail -15 $O/pytest_dp.log
run() { # name, env..., args
  name=$1; shift
  timeout 300 env "$@" > $O/$name.json 2> $O/$name.err
  echo "$name exit $?"; tail -c 1500 $O/$name.json | head -c 1500; echo
}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 100 --warmup 10"
run n2_default $TR
run n2_fp32grad GCT2_DP_GRAD=fp32 $TR
run n2_ctas8 NCCL_MAX_CTAS=8 $TR
run n2_ctas32 NCCL_MAX_CTAS=32 $TR
run n2_replicated GCT2_DP_SHARD=0 $TR
run n2_b8 $TR --batch-per-gpu 8
run n2_strong8 $TR --global-batch 8
run n2_ref python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 2 --warmup 1
echo done

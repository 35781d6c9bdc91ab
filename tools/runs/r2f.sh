#!/bin/bash
mkdir -p gpurun_out/r2f
O=gpurun_out/r2f
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -p no:cacheprovider > $O/pytest_kernels.log 2>&1
echo "exit $?" >> $O/pytest_kernels.log; tail -25 $O/pytest_kernels.log
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q --tb=short -p no:cacheprovider > $O/pytest_engine.log 2>&1
echo "exit $?" >> $O/pytest_engine.log; tail -25 $O/pytest_engine.log
timeout 300 python tools/sweep_step.py --batch 1 --set "" --set key25=0 --set key24=4 > $O/sweep_b1.jsonl 2> $O/sweep_b1.err
cat $O/sweep_b1.jsonl
timeout 200 python tools/sweep_step.py --batch 8 --steps 50 --warmup 5 --set "" > $O/sweep_b8.jsonl 2> $O/sweep_b8.err
cat $O/sweep_b8.jsonl
echo done

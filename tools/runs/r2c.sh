#!/bin/bash
# round 2, call C: cluster split-K (DSMEM), FFMA2 down0 kernels, dense grid sweep
mkdir -p gpurun_out/r2c
O=gpurun_out/r2c
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -p no:cacheprovider > $O/pytest_kernels.log 2>&1
echo "pytest kernels exit $?" >> $O/pytest_kernels.log
tail -15 $O/pytest_kernels.log
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q --tb=short -p no:cacheprovider > $O/pytest_engine.log 2>&1
echo "pytest engine exit $?" >> $O/pytest_engine.log
tail -8 $O/pytest_engine.log
timeout 600 python tools/sweep_step.py --batch 1 \
  --set "" --set key25=1 --set key24=4 --set key24=8 --set key15=148 --set key15=128 --set GCT2_WEIGHTS_EARLY=0 \
  > $O/sweep_b1.jsonl 2> $O/sweep_b1.err
cat $O/sweep_b1.jsonl
timeout 300 python tools/sweep_step.py --batch 8 --steps 50 --warmup 5 --set "" --set key25=1 > $O/sweep_b8.jsonl 2> $O/sweep_b8.err
cat $O/sweep_b8.jsonl
timeout 200 python tools/step_trace.py --csv $O/step_trace_b1.csv > $O/step_trace_b1.txt 2>&1
timeout 200 python tools/timeline.py --batch 1 --weights-stable > $O/timeline_b1.jsonl 2> $O/timeline_b1.err
timeout 300 python bench.py --steps 200 --warmup 20 > $O/bench_b1.json 2> $O/bench_b1.err
echo done

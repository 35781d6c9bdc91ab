#!/bin/bash
# round 2, call A: parity of the restructured conv kernels + first measurements
mkdir -p gpurun_out/r2a
O=gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > $O/pytest.log 2>&1
echo "pytest exit $?" >> $O/pytest.log
tail -5 $O/pytest.log
timeout 300 python bench.py --steps 200 --warmup 20 > $O/bench_b1.json 2> $O/bench_b1.err
echo "bench exit $?"
timeout 120 python tools/adam_probe.py > $O/adam_probe.jsonl 2>&1
timeout 200 python tools/timeline.py --batch 1 > $O/timeline_b1.jsonl 2> $O/timeline_b1.err
timeout 200 python tools/timeline.py --batch 1 --weights-stable > $O/timeline_b1_early.jsonl 2> $O/timeline_b1_early.err
timeout 200 python bench.py --steps 50 --warmup 5 --batch-per-gpu 8 --no-cpu-baseline > $O/bench_b8.json 2> $O/bench_b8.err
timeout 200 python bench.py --steps 30 --warmup 5 --batch-per-gpu 32 --no-cpu-baseline > $O/bench_b32.json 2> $O/bench_b32.err
GCT2_WEIGHTS_EARLY=0 timeout 200 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > $O/bench_b1_noearly.json 2> $O/bench_b1_noearly.err
echo done

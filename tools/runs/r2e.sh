#!/bin/bash
mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "split or pair or cluster or width" > $O/pytest_kernels.log 2>&1
tail -3 $O/pytest_kernels.log
timeout 300 python tools/timeline.py --batch 1 --weights-stable > $O/tl_auto.jsonl 2> $O/tl_auto.err
timeout 400 python tools/timeline.py --batch 1 --weights-stable --passes fprop,dgrad \
  --cfg "3=64 4=8 25=2" --cfg "3=64 4=8 25=1" --cfg "3=64 4=4 25=2" --cfg "3=64 4=4 25=1" --cfg "3=64 4=2 25=2" --cfg "3=64 4=2 25=1" \
  --cfg "3=64 4=1" --cfg "3=128 4=4 25=2" --cfg "3=128 4=4 25=1" --cfg "3=128 4=2 25=2" --cfg "3=128 4=2 25=1" --cfg "3=128 4=8 25=2" --cfg "3=128 4=1" \
  --cfg "3=64 4=16 25=1" --cfg "3=256 4=2 25=2" --cfg "3=256 4=1" > $O/tl_forced.jsonl 2> $O/tl_forced.err
timeout 300 python tools/sweep_step.py --batch 1 --set "" --set key25=1 > $O/sweep_b1.jsonl 2> $O/sweep_b1.err
cat $O/sweep_b1.jsonl
echo done

#!/bin/bash
mkdir -p gpurun_out/r2r
O=gpurun_out/r2r
timeout 200 python tools/timeline.py --batch 1 --weights-stable --cold-weights > $O/timeline_b1_cold.jsonl 2> $O/timeline_b1_cold.err
timeout 200 python tools/timeline.py --batch 1 --weights-stable > $O/timeline_b1_warm.jsonl 2> $O/timeline_b1_warm.err
wc -l $O/timeline_b1_*.jsonl
timeout 560 python tools/tune_plans.py --batch 1 --passes 3 --reach 2 --budget-s 420 --write > $O/tune_b1.jsonl 2> $O/tune_b1.err
tail -4 $O/tune_b1.jsonl | cut -c1-700; tail -3 $O/tune_b1.err
timeout 200 python tools/sweep_step.py --batch 1 --set "" --set GCT2_TUNED=0 --set GCT2_BUCKET_MB=4 --set GCT2_BUCKET_MB=16 > $O/sweep_b1.jsonl 2> $O/sweep_b1.err
cat $O/sweep_b1.jsonl
echo done

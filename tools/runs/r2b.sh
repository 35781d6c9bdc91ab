#!/bin/bash
# round 2, call B: optimiser beside backward on disjoint SMs (sweep), refreshed cost model, step trace
mkdir -p gpurun_out/r2b
O=gpurun_out/r2b
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > $O/pytest.log 2>&1
echo "pytest exit $?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 600 python tools/sweep_step.py --batch 1 \
  --set GCT2_ADAM_SMS=0 \
  --set GCT2_ADAM_SMS=32,GCT2_ADAM_WIDE=5 --set GCT2_ADAM_SMS=48,GCT2_ADAM_WIDE=5 --set GCT2_ADAM_SMS=64,GCT2_ADAM_WIDE=5 \
  --set GCT2_ADAM_SMS=48,GCT2_ADAM_WIDE=3 --set GCT2_ADAM_SMS=48,GCT2_ADAM_WIDE=7 --set GCT2_ADAM_SMS=64,GCT2_ADAM_WIDE=7 \
  --set GCT2_ADAM_SMS=64,GCT2_ADAM_WIDE=9 --set GCT2_ADAM_SMS=40,GCT2_ADAM_WIDE=6 --set GCT2_ADAM_SMS=56,GCT2_ADAM_WIDE=6 \
  --set GCT2_ADAM_SMS=0,GCT2_BUCKET_MB=24 --set GCT2_ADAM_SMS=0,GCT2_WEIGHTS_EARLY=0 \
  > $O/sweep_b1.jsonl 2> $O/sweep_b1.err
cat $O/sweep_b1.jsonl
timeout 300 python tools/sweep_step.py --batch 8 --steps 50 --warmup 5 --set GCT2_ADAM_SMS=0 --set GCT2_ADAM_SMS=48,GCT2_ADAM_WIDE=5 --set GCT2_ADAM_SMS=32,GCT2_ADAM_WIDE=7 > $O/sweep_b8.jsonl 2> $O/sweep_b8.err
cat $O/sweep_b8.jsonl
timeout 200 python tools/step_trace.py --csv $O/step_trace_b1.csv > $O/step_trace_b1.txt 2>&1
GCT2_ADAM_SMS=0 timeout 200 python tools/step_trace.py --csv $O/step_trace_b1_adam0.csv > $O/step_trace_b1_adam0.txt 2>&1
timeout 300 python bench.py --steps 200 --warmup 20 > $O/bench_b1.json 2> $O/bench_b1.err
echo done

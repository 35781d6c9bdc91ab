#!/bin/bash
mkdir -p gpurun_out/r2h
O=gpurun_out/r2h
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > $O/pytest.log 2>&1
echo "exit $?" >> $O/pytest.log; tail -30 $O/pytest.log
timeout 300 python tools/sweep_step.py --batch 1 --set "" --set key21=1 --set GCT2_BUCKET_MB=24 --set key24=4 > $O/sweep_b1.jsonl 2> $O/sweep_b1.err
cat $O/sweep_b1.jsonl
timeout 200 python tools/sweep_step.py --batch 8 --steps 50 --warmup 5 --set "" > $O/sweep_b8.jsonl 2> $O/sweep_b8.err
cat $O/sweep_b8.jsonl
timeout 200 python tools/sweep_step.py --batch 32 --steps 30 --warmup 5 --set "" > $O/sweep_b32.jsonl 2> $O/sweep_b32.err
cat $O/sweep_b32.jsonl
for c in "default 1" "default 8" "tiny 2" "wide 1"; do set -- $c; timeout 200 python tools/parity_table.py --config $1 --batch $2 >> $O/parity.jsonl 2>> $O/parity.err; done
timeout 200 python tools/parity_table.py --config default --batch 1 --mixed-precision >> $O/parity.jsonl 2>> $O/parity.err
timeout 200 python tools/step_trace.py --csv $O/step_trace_b1.csv > $O/step_trace_b1.txt 2>&1
echo done

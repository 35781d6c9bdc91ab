#!/bin/bash
mkdir -p gpurun_out/r2u
O=gpurun_out/r2u
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > $O/pytest.log 2>&1
echo "exit $?" >> $O/pytest.log; tail -8 $O/pytest.log
timeout 460 python tools/tune_plans.py --batch 1 --passes 2 --reach 1 --budget-s 330 --write > $O/tune_b1.jsonl 2> $O/tune_b1.err
tail -3 $O/tune_b1.jsonl | cut -c1-700; tail -3 $O/tune_b1.err
timeout 200 python tools/sweep_step.py --batch 1 --set "" --set GCT2_TUNED=0 > $O/sweep_b1.jsonl 2> $O/sweep_b1.err
cat $O/sweep_b1.jsonl
echo done

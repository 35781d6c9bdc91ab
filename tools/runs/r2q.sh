#!/bin/bash
mkdir -p gpurun_out/r2q
O=gpurun_out/r2q
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x > $O/pytest_kernels.log 2>&1
echo "exit $?" >> $O/pytest_kernels.log; tail -5 $O/pytest_kernels.log
timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x -k "default_config_batch1 or tiny_batch2 or ragged or loss_curve_tiny or forced" > $O/pytest_engine.log 2>&1
echo "exit $?" >> $O/pytest_engine.log; tail -5 $O/pytest_engine.log
timeout 200 python tools/sweep_step.py --batch 1 --set "" --set GCT2_TUNED=0 > $O/sweep_b1.jsonl 2> $O/sweep_b1.err
cat $O/sweep_b1.jsonl
timeout 200 python tools/step_trace.py --csv $O/step_trace_b1.csv > $O/step_trace_b1.txt 2>&1
tail -3 $O/step_trace_b1.txt
echo done

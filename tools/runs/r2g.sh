#!/bin/bash
mkdir -p gpurun_out/r2g
O=gpurun_out/r2g
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -p no:cacheprovider > $O/pytest_kernels.log 2>&1
echo "exit $?" >> $O/pytest_kernels.log; tail -15 $O/pytest_kernels.log
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q --tb=short -p no:cacheprovider > $O/pytest_engine.log 2>&1
echo "exit $?" >> $O/pytest_engine.log; tail -30 $O/pytest_engine.log
for v in "" _nof16 _nof16inl; do
  GCT2_LIB=$PWD/gan_class_transfer2_b200/libgct2_b200$v.so timeout 200 python tools/sweep_step.py --batch 1 --set "" > $O/sweep_b1$v.jsonl 2> $O/sweep_b1$v.err
  echo "variant '$v'"; cat $O/sweep_b1$v.jsonl
done
timeout 200 python tools/step_trace.py --csv $O/step_trace_b1.csv > $O/step_trace_b1.txt 2>&1
echo done

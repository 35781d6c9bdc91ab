#!/bin/bash
mkdir -p gpurun_out/r2ab
O=gpurun_out/r2ab
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x > $O/pytest_kernels.log 2>&1
echo "exit $?" >> $O/pytest_kernels.log; tail -6 $O/pytest_kernels.log | cut -c1-600
timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x -k "default_config or tiny_batch2 or ragged or forced or sampling_loops_default or golden" > $O/pytest_engine.log 2>&1
echo "exit $?" >> $O/pytest_engine.log; tail -6 $O/pytest_engine.log | cut -c1-600
timeout 200 python tools/sweep_step.py --batch 32 --steps 30 --warmup 5 --set "" --set key27=1 > $O/sweep_b32.jsonl 2> $O/sweep_b32.err; cat $O/sweep_b32.jsonl
timeout 200 python tools/sweep_step.py --batch 8 --steps 60 --warmup 8 --set "" --set key27=1 > $O/sweep_b8.jsonl 2> $O/sweep_b8.err; cat $O/sweep_b8.jsonl
timeout 200 python tools/sweep_step.py --batch 1 --steps 200 --set "" --set key27=1 --set "" --set key27=1 > $O/sweep_b1.jsonl 2> $O/sweep_b1.err; cat $O/sweep_b1.jsonl
echo done

#!/bin/bash
mkdir -p gpurun_out/r2ak
timeout 200 python -m pytest tests/test_engine_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "forced and (depth2 or residual)" > gpurun_out/r2ak/pytest.log 2>&1
echo "exit $?" >> gpurun_out/r2ak/pytest.log; tail -12 gpurun_out/r2ak/pytest.log | cut -c1-600

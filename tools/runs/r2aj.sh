#!/bin/bash
mkdir -p gpurun_out/r2aj
O=gpurun_out/r2aj
timeout 400 python -m pytest tests/test_engine_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "dormant or layer_list or block or forced or residual or golden_fixtures or fp16_policy" > $O/pytest.log 2>&1
echo "exit $?" >> $O/pytest.log; tail -5 $O/pytest.log | cut -c1-400

#!/bin/bash
mkdir -p gpurun_out/r2o
O=gpurun_out/r2o
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "dormant or layer_list or block or surface or forced" > $O/pytest.log 2>&1
echo "exit $?" >> $O/pytest.log; tail -120 $O/pytest.log | cut -c1-600
echo done

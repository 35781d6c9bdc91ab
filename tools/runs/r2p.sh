#!/bin/bash
mkdir -p gpurun_out/r2p
O=gpurun_out/r2p
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_kernels_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "dormant or layer_list or block or surface or forced or residual or stride1" > $O/pytest.log 2>&1
echo "exit $?" >> $O/pytest.log; tail -150 $O/pytest.log | cut -c1-1800
echo done

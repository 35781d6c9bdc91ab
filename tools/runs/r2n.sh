#!/bin/bash
mkdir -p gpurun_out/r2n
O=gpurun_out/r2n
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "stride1 or block_cuda" > $O/pytest.log 2>&1
echo "exit $?" >> $O/pytest.log; tail -60 $O/pytest.log | cut -c1-400
echo done

#!/bin/bash
mkdir -p gpurun_out/r2ad
O=gpurun_out/r2ad
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "block_cuda or hbm_bound_kernels_match" > $O/plain.log 2>&1 &&
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "block_cuda or hbm_bound_kernels_match" > $O/memcheck.log 2>&1
echo "memcheck exit $?"; tail -15 $O/memcheck.log | cut -c1-300
echo done

#!/bin/bash
mkdir -p gpurun_out/r2aa
O=gpurun_out/r2aa
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > $O/pytest.log 2>&1
echo "exit $?" >> $O/pytest.log; tail -40 $O/pytest.log | cut -c1-1500
echo done

#!/bin/bash
mkdir -p gpurun_out/r2ah
O=gpurun_out/r2ah
timeout 400 python -m pytest tests/test_dp_gpu.py -m gpu -q --tb=short -p no:cacheprovider > $O/pytest_dp.log 2>&1
echo "exit $?" >> $O/pytest_dp.log; tail -3 $O/pytest_dp.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 100 --warmup 10 --no-cpu-baseline > $O/n2.json 2> $O/n2.err; echo "n2 exit $?"
python -c "
import json; d=json.loads(open('$O/n2.json').read().strip().split(chr(10))[-1]); c=d['comm']; print('n2 value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1),'exposed',round(c['comm_exposed_ms'],3),c['transport'],c['head_and_loss'],'equal',c['replicas_bit_equal'])"
echo done

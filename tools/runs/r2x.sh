#!/bin/bash
mkdir -p gpurun_out/r2x
O=gpurun_out/r2x
timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -q --tb=short -p no:cacheprovider -s > $O/pytest_dp.log 2>&1
echo "exit $?" >> $O/pytest_dp.log; grep -E "transport used|passed|failed|Error|error" $O/pytest_dp.log | head -12
run() { name=$1; shift; timeout 300 env "$@" > $O/$name.json 2> $O/$name.err; echo "$name exit $?"; grep -i "warn\|error" $O/$name.err | head -3; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 100 --warmup 10 --no-cpu-baseline"
run n2_head_p2p $TR
run n2_head_nccl GCT2_DP_HEAD=nccl $TR
run n2_head_p2p_again $TR
run n2_head_nccl_again GCT2_DP_HEAD=nccl $TR
run n2_strong8 $TR --global-batch 8
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2x/n2_*.json')):
    try: d=json.loads(open(f).read().strip().split('\n')[-1])
    except Exception as e: print(f,'ERR',e); continue
    c=d.get('comm') or {}
    print(f.split('/')[-1].ljust(28),'value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'compute_only',round(c.get('compute_only_ms',0),3),'comm_alone',round(c.get('comm_alone_ms',0),3),'exposed',round(c.get('comm_exposed_ms',0),3),c.get('transport'),c.get('head_and_loss'),'equal',c.get('replicas_bit_equal'),'launches',d.get('launches_per_step'),'loss',d.get('final_loss'))
PY
echo done

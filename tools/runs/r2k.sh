#!/bin/bash
mkdir -p gpurun_out/r2k
O=gpurun_out/r2k
timeout 900 python -m pytest tests/test_dp_gpu.py -m gpu -q --tb=short -p no:cacheprovider -s > $O/pytest_dp.log 2>&1
echo "exit $?" >> $O/pytest_dp.log; grep -E "transport used|passed|failed|Error|error|warn" $O/pytest_dp.log | head -30
run() { name=$1; shift; timeout 300 env "$@" > $O/$name.json 2> $O/$name.err; echo "$name exit $?"; grep -i "warn\|error" $O/$name.err | head -3; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 100 --warmup 10 --no-cpu-baseline"
run n2_p2p $TR
run n2_p2p_nomc GCT2_DP_MULTICAST=0 $TR
run n2_p2p_bucket96 GCT2_DP_BUCKET_MB=96 $TR
run n2_p2p_bucket16 GCT2_DP_BUCKET_MB=16 $TR
run n2_nccl GCT2_DP_TRANSPORT=nccl GCT2_DP_BUCKET_MB=96 $TR
run n2_p2p_b8 $TR --batch-per-gpu 8
run n2_p2p_strong8 $TR --global-batch 8
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2k/n2_*.json')):
    try: d=json.loads(open(f).read().strip().split('\n')[-1])
    except Exception as e: print(f,'ERR',e); continue
    c=d.get('comm') or {}
    print(f.split('/')[-1].ljust(24),'value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'compute_only',round(c.get('compute_only_ms',0),3),'comm_alone',round(c.get('comm_alone_ms',0),3),'exposed',round(c.get('comm_exposed_ms',0),3),c.get('transport'))
PY
echo done

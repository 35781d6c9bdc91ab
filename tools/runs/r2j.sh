#!/bin/bash
mkdir -p gpurun_out/r2j
O=gpurun_out/r2j
run() { name=$1; shift; timeout 300 env "$@" > $O/$name.json 2> $O/$name.err; echo "$name exit $?"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 100 --warmup 10 --no-cpu-baseline"
run n2_default $TR
run n2_shared GCT2_DP_SM=shared $TR
run n2_noprio GCT2_NCCL_PRIORITY=0 $TR
run n2_shared_noprio GCT2_DP_SM=shared GCT2_NCCL_PRIORITY=0 $TR
run n2_ctas32 NCCL_MAX_CTAS=32 $TR
run n2_bucket96 GCT2_DP_BUCKET_MB=96 $TR
run n2_bucket24 GCT2_DP_BUCKET_MB=24 $TR
run n2_strong8 $TR --global-batch 8
timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "bf16 and cuda_graph" > $O/pytest_dp.log 2>&1
tail -3 $O/pytest_dp.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2j/n2_*.json')):
    try: d=json.loads(open(f).read().strip().split('\n')[-1])
    except Exception as e: print(f,'ERR',e); continue
    c=d.get('comm') or {}
    print(f.split('/')[-1].ljust(24),'value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'compute_only',round(c.get('compute_only_ms',0),3),'comm_alone',round(c.get('comm_alone_ms',0),3),'exposed',round(c.get('comm_exposed_ms',0),3))
PY
echo done

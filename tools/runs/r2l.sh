#!/bin/bash
mkdir -p gpurun_out/r2l
O=gpurun_out/r2l
nvidia-smi -L | wc -l > $O/ngpus.txt
run() { name=$1; n=$2; shift 2; timeout 240 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $n --steps 100 --warmup 10 --no-cpu-baseline > $O/$name.json 2> $O/$name.err; echo "$name exit $?"; grep -i "warn\|error" $O/$name.err | head -2; }
run n8_p2p_mc 8 GCT2_X=1
run n8_p2p_nomc 8 GCT2_DP_MULTICAST=0
run n8_nccl 8 GCT2_DP_TRANSPORT=nccl
run n4_p2p_mc 4 GCT2_X=1
run n4_p2p_nomc 4 GCT2_DP_MULTICAST=0
run n2_p2p_mc 2 GCT2_X=1
run n2_p2p_nomc 2 GCT2_DP_MULTICAST=0
run n8_p2p_mc_b96 8 GCT2_DP_BUCKET_MB=96
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 100 --warmup 10 --no-cpu-baseline --global-batch 8 > $O/n8_strong8.json 2> $O/n8_strong8.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2l/n*.json')):
    try: d=json.loads(open(f).read().strip().split('\n')[-1])
    except Exception as e: print(f,'ERR',e); continue
    c=d.get('comm') or {}
    print(f.split('/')[-1].ljust(24),'value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'compute_only',round(c.get('compute_only_ms',0),3),'comm_alone',round(c.get('comm_alone_ms',0),3),'exposed',round(c.get('comm_exposed_ms',0),3),c.get('transport'))
PY
echo done

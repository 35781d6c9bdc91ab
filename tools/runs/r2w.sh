#!/bin/bash
mkdir -p gpurun_out/r2w
O=gpurun_out/r2w
nvidia-smi -L | wc -l > $O/ngpus.txt
timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -q --tb=short -p no:cacheprovider -s > $O/pytest_dp.log 2>&1
echo "exit $?" >> $O/pytest_dp.log; grep -E "transport used|passed|failed|Error|error" $O/pytest_dp.log | head -12
GCT2_DP_L2_FINISH=1 timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -q --tb=short -p no:cacheprovider -s > $O/pytest_dp_l2.log 2>&1
echo "exit $?" >> $O/pytest_dp_l2.log; grep -E "transport used|passed|failed|Error|error" $O/pytest_dp_l2.log | head -12
run() { name=$1; shift; timeout 300 env "$@" > $O/$name.json 2> $O/$name.err; echo "$name exit $?"; grep -i "warn\|error" $O/$name.err | head -3; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 100 --warmup 10 --no-cpu-baseline"
run n2_default $TR
run n2_l2finish GCT2_DP_L2_FINISH=1 $TR
run n2_default_again $TR
run n2_l2finish_again GCT2_DP_L2_FINISH=1 $TR
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2w/n2_*.json')):
    try: d=json.loads(open(f).read().strip().split('\n')[-1])
    except Exception as e: print(f,'ERR',e); continue
    c=d.get('comm') or {}
    print(f.split('/')[-1].ljust(28),'value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'compute_only',round(c.get('compute_only_ms',0),3),'comm_alone',round(c.get('comm_alone_ms',0),3),'exposed',round(c.get('comm_exposed_ms',0),3),c.get('transport'),'equal',c.get('replicas_bit_equal'),'launches',d.get('launches_per_step'))
PY
for extra in "--residual" "--residual --block-depth 1"; do
  timeout 200 python tools/parity_table.py --config tiny --batch 2 $extra --forced >> $O/parity_forced_residual.jsonl 2>> $O/parity.err
  timeout 200 python tools/parity_table.py --config tiny --batch 2 $extra >> $O/parity_residual.jsonl 2>> $O/parity.err
done
python - <<'PY'
import json
mx={}
for l in open('gpurun_out/r2w/parity_forced_residual.jsonl'):
    d=json.loads(l)
    if d['quantity']!='loss': mx[(d['block_depth'],d['residual'])]=max(mx.get((d['block_depth'],d['residual']),0),d['err'])
print('forced residual max err', mx)
PY
echo done

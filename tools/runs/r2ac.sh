#!/bin/bash
mkdir -p gpurun_out/r2ac
O=gpurun_out/r2ac
M=gpu__time_duration.sum,launch__grid_size,launch__registers_per_thread,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active
python tools/profile_step.py > $O/profile_plain_b1.log 2>&1 &&
ncu --profile-from-start off --clock-control none -k regex:conv_umma --metrics $M --csv --log-file $O/conv_family_b1_raw.csv python tools/profile_step.py > $O/ncu_family_b1.log 2>&1
echo "ncu family b1 exit $?"
python tools/profile_step.py --batch 32 > $O/profile_plain_b32.log 2>&1 &&
ncu --profile-from-start off --clock-control none -k regex:conv_umma --metrics $M --csv --log-file $O/conv_family_b32_raw.csv python tools/profile_step.py --batch 32 > $O/ncu_family_b32.log 2>&1
echo "ncu family b32 exit $?"
python tools/ncu_summary.py --raw $O/conv_family_b1_raw.csv --out profiles/r2_conv_family_ncu_b1.csv --traffic-key b1
python tools/ncu_summary.py --raw $O/conv_family_b32_raw.csv --out profiles/r2_conv_family_ncu_b32.csv --traffic-key b32
cp profiles/roofline_traffic.json profiles/r2_conv_family_ncu_b1.csv profiles/r2_conv_family_ncu_b32.csv $O/
timeout 600 python bench.py > $O/bench_b1.json 2> $O/bench_b1.err; echo "bench b1 exit $?"
timeout 300 python bench.py --batch-per-gpu 8 --steps 60 --warmup 8 --no-cpu-baseline > $O/bench_b8.json 2> $O/bench_b8.err; echo "b8 exit $?"
timeout 300 python bench.py --batch-per-gpu 32 --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_b32.json 2> $O/bench_b32.err; echo "b32 exit $?"
python - <<'PY'
import json
for n in ('bench_b1','bench_b8','bench_b32'):
    d=json.loads(open(f'gpurun_out/r2ac/{n}.json').read().strip().split('\n')[-1]); r=d['roofline']
    print(n,'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',round(r['frac'],4),'traffic',r['traffic'],'clk',d['clocks'])
PY
echo done

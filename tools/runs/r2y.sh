#!/bin/bash
mkdir -p gpurun_out/r2y
O=gpurun_out/r2y
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
timeout 600 python bench.py > $O/bench_b1.json 2> $O/bench_b1.err; echo "bench b1 exit $?"; cut -c1-300 $O/bench_b1.json
python -c "
import json; d=json.loads(open('$O/bench_b1.json').read().strip().split(chr(10))[-1]); r=d['roofline']; print('value',d['value'],'e2e',d['e2e']['value'],'frac',r['frac'],'traffic',r['traffic'],r['traffic_source'])"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?"; tail -2 $O/smoke.log
echo done

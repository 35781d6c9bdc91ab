#!/bin/bash
# Round-2 final measurements on one B200: bench lines, per-layer microbench, sampling loops, config 4, parity tables,
# step trace, ncu launch list + per-launch metrics of the conv family + one full-set capture of the top kernel.
mkdir -p gpurun_out/r2v
O=gpurun_out/r2v
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $O/smi.txt
timeout 600 python bench.py > $O/bench_b1.json 2> $O/bench_b1.err; echo "bench b1 exit $?"; cut -c1-400 $O/bench_b1.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err; echo "ref exit $?"; cut -c1-300 $O/bench_reference_arm.json
timeout 300 python bench.py --batch-per-gpu 8 --steps 60 --warmup 8 --no-cpu-baseline > $O/bench_b8.json 2> $O/bench_b8.err; echo "b8 exit $?"; cut -c1-200 $O/bench_b8.json
timeout 300 python bench.py --batch-per-gpu 32 --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_b32.json 2> $O/bench_b32.err; echo "b32 exit $?"; cut -c1-200 $O/bench_b32.json
for b in 1 8 32; do timeout 200 python tools/bench_layers.py --batch $b > $O/layers_b$b.jsonl 2> $O/layers_b$b.err; done
timeout 300 python tools/bench_sample.py > $O/sampling_loops.jsonl 2> $O/sampling_loops.err
timeout 300 python tools/bench_config4.py > $O/config4_wide_n1.json 2> $O/config4_wide_n1.err
for c in "default 1" "default 8" "tiny 2"; do set -- $c; timeout 200 python tools/parity_table.py --config $1 --batch $2 >> $O/parity.jsonl 2>> $O/parity.err; done
timeout 200 python tools/parity_table.py --config default --batch 1 --mixed-precision >> $O/parity.jsonl 2>> $O/parity.err
for extra in "--block-depth 1" "--block-depth 2" "--no-concat" "--block-depth 1 --no-concat"; do
  timeout 200 python tools/parity_table.py --config tiny --batch 2 $extra >> $O/parity_dormant.jsonl 2>> $O/parity.err
  timeout 200 python tools/parity_table.py --config tiny --batch 2 $extra --forced >> $O/parity_forced.jsonl 2>> $O/parity.err
done
timeout 200 python tools/parity_table.py --config default --batch 1 --forced >> $O/parity_forced.jsonl 2>> $O/parity.err
timeout 200 python tools/parity_table.py --config default --batch 1 --block-depth 1 --forced >> $O/parity_forced.jsonl 2>> $O/parity.err
timeout 200 python tools/parity_table.py --config default --batch 1 --block-depth 1 >> $O/parity_dormant.jsonl 2>> $O/parity.err
timeout 200 python tools/step_trace.py --csv $O/step_trace_b1.csv > $O/step_trace_b1.txt 2>&1
M=gpu__time_duration.sum,launch__grid_size,launch__registers_per_thread,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active
python tools/profile_step.py > $O/profile_plain_b1.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_one_step_b1_raw.csv python tools/profile_step.py > $O/ncu_launches.log 2>&1
echo "ncu launches exit $?"
ncu --profile-from-start off --clock-control none -k regex:conv_umma --metrics $M --csv --log-file $O/conv_family_b1_raw.csv python tools/profile_step.py > $O/ncu_family_b1.log 2>&1
echo "ncu family b1 exit $?"
python tools/profile_step.py --batch 32 > $O/profile_plain_b32.log 2>&1 &&
ncu --profile-from-start off --clock-control none -k regex:conv_umma --metrics $M --csv --log-file $O/conv_family_b32_raw.csv python tools/profile_step.py --batch 32 > $O/ncu_family_b32.log 2>&1
echo "ncu family b32 exit $?"
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_umma -s 16 -c 3 -o $O/conv_top3_full python tools/profile_step.py > $O/ncu_full.log 2>&1
echo "ncu full exit $?"; ls -la $O | head -50
echo done

#!/bin/bash
mkdir -p gpurun_out/r2ai
O=gpurun_out/r2ai
for b in 1 8 32; do timeout 100 python tools/bench_layers.py --batch $b > $O/layers_b$b.jsonl 2> $O/layers_b$b.err; done
timeout 120 python tools/bench_config4.py > $O/config4_wide_n1.json 2> $O/config4_wide_n1.err
cat $O/config4_wide_n1.json; grep '"up0"' $O/layers_b32.jsonl | cut -c1-200
echo done

#!/bin/bash
mkdir -p gpurun_out/r2s
O=gpurun_out/r2s
P=gan_class_transfer2_b200
run() { lib=$1; shift; GCT2_LIB=$PWD/$P/libgct2_b200$lib.so timeout 200 python tools/sweep_step.py --batch 1 --steps 150 "$@" > $O/sweep$lib.jsonl 2> $O/sweep$lib.err; echo "== lib '$lib'"; cat $O/sweep$lib.jsonl; tail -2 $O/sweep$lib.err; }
run "" --set "" --set key13=296 --set key13=592
run _v1 --set "" --set key13=296 --set key13=444 --set key13=592
run _v2 --set "" --set key13=148 --set key13=296
run _v3 --set "" --set key13=296 --set key13=444
run _v4 --set "" --set key13=148 --set key13=296
GCT2_LIB=$PWD/$P/libgct2_b200_v1.so timeout 100 python tools/step_trace.py --csv $O/step_trace_v1.csv > $O/step_trace_v1.txt 2>&1
echo done

#!/bin/bash
mkdir -p gpurun_out/r2t
O=gpurun_out/r2t
P=gan_class_transfer2_b200
run() { lib=$1; b=$2; shift 2; GCT2_LIB=$PWD/$P/libgct2_b200$lib.so timeout 200 python tools/sweep_step.py --batch $b "$@" > $O/sweep${lib}_b$b.jsonl 2> $O/sweep${lib}_b$b.err; echo "== lib '$lib' batch $b"; cat $O/sweep${lib}_b$b.jsonl; tail -2 $O/sweep${lib}_b$b.err; }
run "" 8 --steps 60 --warmup 8 --set ""
run _v4 8 --steps 60 --warmup 8 --set ""
run "" 32 --steps 30 --warmup 5 --set ""
run _v4 32 --steps 30 --warmup 5 --set ""
run "" 1 --steps 200 --set ""
run _v4 1 --steps 200 --set ""
echo done

#!/bin/bash
mkdir -p gpurun_out/r2ag
O=gpurun_out/r2ag
P=gan_class_transfer2_b200
run() { lib=$1; b=$2; shift 2; GCT2_LIB=$PWD/$P/libgct2_b200$lib.so timeout 200 python tools/sweep_step.py --batch $b "$@" > $O/sweep${lib}_b$b.jsonl 2> $O/sweep${lib}_b$b.err; echo "== lib '$lib' batch $b"; cat $O/sweep${lib}_b$b.jsonl; tail -2 $O/sweep${lib}_b$b.err; }
run "" 1 --steps 300 --set "" --set ""
run _c3ty2 1 --steps 300 --set "" --set ""
run _both 1 --steps 300 --set "key24=3" --set "key24=3" --set "key24=2"
run _both 32 --steps 30 --warmup 5 --set "key24=3"
run "" 32 --steps 30 --warmup 5 --set ""
echo done

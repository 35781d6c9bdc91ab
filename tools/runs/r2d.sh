#!/bin/bash
mkdir -p gpurun_out/r2d
O=gpurun_out/r2d
timeout 200 python tools/timeline.py --batch 1 --weights-stable > $O/tl_auto.jsonl 2> $O/tl_auto.err
timeout 200 python tools/timeline.py --batch 1 --weights-stable --debug 25=1 > $O/tl_nocluster.jsonl 2> $O/tl_nocluster.err
for L in down3 down4 up4 up3 down2 up2; do
 for cfg in "3=64 4=8 25=2" "3=64 4=8 25=1" "3=64 4=4 25=2" "3=64 4=2 25=2" "3=64 4=1" "3=128 4=4 25=2" "3=128 4=4 25=1" "3=128 4=2 25=2" "3=128 4=8 25=2" "3=64 4=16 25=1"; do
   args=""; for kv in $cfg; do args="$args --debug $kv"; done
   echo "## $L $cfg" >> $O/tl_forced.jsonl
   timeout 60 python tools/timeline.py --batch 1 --weights-stable --only $L $args 2>> $O/tl_forced.err | grep -v wgrad >> $O/tl_forced.jsonl
 done
done
echo done

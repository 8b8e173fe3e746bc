#!/bin/bash
# one full ncu capture of a kernel of the micro-benchmark: tools/ncu_dw.sh <kernel-regex> <out-name> [rays]
K=$1; OUT=$2; RAYS=${3:-2048}
ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -s 2 -o gpurun_out/$OUT -f \
    python tools/bench_mlp_tc.py --rays $RAYS --samples 128 --save --iters 1 > gpurun_out/$OUT.log 2>&1

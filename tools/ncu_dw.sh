#!/bin/bash
# one full ncu capture of a kernel of the micro-benchmark:
#   tools/ncu_dw.sh <kernel-regex> <out-name> [rays] [skip]
# (bench_mlp_tc.py --iters 1 launches each MLP kernel 3 warm-up + 1 timed times; `skip` selects which)
K=$1; OUT=$2; RAYS=${3:-2048}; SKIP=${4:-2}
ncu --set full --clock-control none --import-source on -k "regex:$K" -c 1 -s $SKIP -o gpurun_out/$OUT -f \
    python tools/bench_mlp_tc.py --rays $RAYS --samples 128 --save --iters 1 > gpurun_out/$OUT.log 2>&1

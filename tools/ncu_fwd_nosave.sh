#!/bin/bash
# full ncu capture of the inference forward (no activation saves): tools/ncu_fwd_nosave.sh <out-name> [rays]
OUT=$1; RAYS=${2:-4096}
ncu --set full --clock-control none --import-source on -k "regex:mlp_tc_chain" -c 1 -s 3 -o gpurun_out/$OUT -f \
    python tools/bench_mlp_tc.py --rays $RAYS --samples 128 --iters 1 > gpurun_out/$OUT.log 2>&1

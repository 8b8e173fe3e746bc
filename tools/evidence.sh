#!/bin/bash
# Round evidence on ONE GPU: GPU test suite, the bench line, the ncu launch list of the bench command and
# `ncu --set full` summaries of the MLP kernels.  .ncu-rep files stay in /tmp; summaries land in gpurun_out/<tag>_*.
#   bash tools/evidence.sh <tag>
TAG=${1:-rXX}
mkdir -p gpurun_out /tmp/ncu_$TAG
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -1 gpurun_out/${TAG}_pytest_gpu.log
python bench.py > gpurun_out/${TAG}_bench_bf16.json 2> gpurun_out/${TAG}_bench_bf16.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2>> gpurun_out/${TAG}_bench_bf16.err
# launch list of the same command (eager steps so that every kernel is a separate launch)
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-graph --no-render --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches_summary.md
# the DDNeRF step (cfg1: config_blender.yml, 1024 rays)
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches_cfg1.csv \
    python bench.py --workload cfg1 --steps 2 --warmup 3 --no-graph --no-render --no-cfg4 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches_cfg1.log 2>&1
python tools/summarize_launches.py gpurun_out/${TAG}_launches_cfg1.csv > gpurun_out/${TAG}_launches_cfg1_summary.md
# full captures of the MLP kernels (micro-benchmark, 4096 rays x 128 samples)
cap() {  # name regex skip extra-args
  ncu --set full --clock-control none --import-source on -k "regex:$2" -c 1 -s $3 -f -o /tmp/ncu_$TAG/$1 \
      python tools/bench_mlp_tc.py --rays 4096 --samples 128 --iters 1 $4 > gpurun_out/${TAG}_ncu_$1.log 2>&1
  python tools/ncu_summary.py /tmp/ncu_$TAG/$1.ncu-rep > gpurun_out/${TAG}_ncu_$1.md 2>> gpurun_out/${TAG}_ncu_$1.log && echo "$1 ok"
}
# (ncu matches the function name without template arguments: the pair kernel's launches in bench_mlp_tc.py are, in order,
#  image-fed forward x4 [3 warm-up + 1], ray-fed forward x4, then -- with --save -- dX chain x3)
cap mlp_tc_pair_fwd_nosave "mlp_tc_pair_kernel" 3 ""
cap mlp_tc_pair_fwd "mlp_tc_pair_kernel" 3 "--save"
cap mlp_tc_pair_fwd_rays "mlp_tc_pair_kernel" 7 "--save"
cap mlp_tc_pair_dx "mlp_tc_pair_kernel" 10 "--save"
cap mlp_tc_dw "mlp_tc_dw_kernel" 2 "--save"

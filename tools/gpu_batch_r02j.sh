cd $GRAFT_REPO_ROOT
TAG=r02j
mkdir -p /tmp/ncu_$TAG
# launch list of the bench command (eager steps so that every kernel is a separate launch), cfg2 and cfg1 (DDNeRF)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-graph --no-render --no-cfg4 --no-cpu-baseline > gpurun_out/${TAG}_ncu_cfg2.log 2>&1; echo "ncu cfg2 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches_cfg1.csv python bench.py --steps 2 --warmup 3 --no-graph --no-render --no-cfg4 --no-cpu-baseline --workload cfg1 > gpurun_out/${TAG}_ncu_cfg1.log 2>&1; echo "ncu cfg1 rc=$?"
python tools/summarize_launches.py gpurun_out/${TAG}_launches_cfg2.csv > gpurun_out/${TAG}_launches_cfg2_summary.md
python tools/summarize_launches.py gpurun_out/${TAG}_launches_cfg1.csv > gpurun_out/${TAG}_launches_cfg1_summary.md
cap() {  # name regex skip extra-args
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$2" -c 1 -s $3 -f -o /tmp/ncu_$TAG/$1 \
      python tools/bench_mlp_tc.py --rays 4096 --samples 128 --iters 1 --pair 1 $4 > gpurun_out/${TAG}_ncu_$1.log 2>&1
  python tools/ncu_summary.py /tmp/ncu_$TAG/$1.ncu-rep > gpurun_out/${TAG}_ncu_$1.md 2>> gpurun_out/${TAG}_ncu_$1.log && echo "$1 ok"
}
# launches of the pair kernel in bench_mlp_tc: image-fed forward x4, ray-fed forward x4, then (--save) dX chain x3
cap mlp_tc_pair_fwd_nosave "mlp_tc_pair_kernel" 3 ""
cap mlp_tc_pair_fwd "mlp_tc_pair_kernel" 3 "--save"
cap mlp_tc_pair_fwd_rays "mlp_tc_pair_kernel" 7 "--save"
cap mlp_tc_pair_dx "mlp_tc_pair_kernel" 10 "--save"
cap mlp_tc_dw "mlp_tc_dw_kernel" 2 "--save"

cd $GRAFT_REPO_ROOT
timeout 200 python tools/check_pair.py > gpurun_out/pair_check.log 2>&1; echo check rc=$?; tail -2 gpurun_out/pair_check.log
rm -f gpurun_out/pair_micro.log
for p in 0 1; do
timeout 200 python tools/bench_mlp_tc.py --rays 4096 --samples 128 --iters 20 --pair $p --prof >> gpurun_out/pair_micro.log 2>&1
timeout 200 python tools/bench_mlp_tc.py --rays 4096 --samples 128 --iters 20 --pair $p --save --prof >> gpurun_out/pair_micro.log 2>&1
done
grep -v "^dw op" gpurun_out/pair_micro.log
mkdir -p /tmp/ncu
cap() {
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$2" -c 1 -s $3 -f -o /tmp/ncu/$1 \
      python tools/bench_mlp_tc.py --rays 4096 --samples 128 --iters 1 --pair 1 $4 > gpurun_out/pairncu_$1.log 2>&1
  python tools/ncu_summary.py /tmp/ncu/$1.ncu-rep > gpurun_out/pairncu_$1.md 2>> gpurun_out/pairncu_$1.log && echo "$1 ok"
  python tools/ncu_roles.py /tmp/ncu/$1.ncu-rep > gpurun_out/pairncu_$1_roles.txt 2>&1
}
cap fwd_nosave "mlp_tc_pair_kernel" 3 ""
cap dx "mlp_tc_pair_kernel" 8 "--save"

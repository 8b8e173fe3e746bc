cd $GRAFT_REPO_ROOT
timeout 200 python tools/check_pair.py > gpurun_out/pair_check.log 2>&1; echo check rc=$?; tail -3 gpurun_out/pair_check.log
rm -f gpurun_out/pair_micro.log
for p in ${PAIRS:-1}; do
  timeout 200 python tools/bench_mlp_tc.py --rays 4096 --samples 128 --iters 20 --pair $p --prof >> gpurun_out/pair_micro.log 2>&1
  timeout 200 python tools/bench_mlp_tc.py --rays 4096 --samples 128 --iters 20 --pair $p --save --prof >> gpurun_out/pair_micro.log 2>&1
done
grep -v "^dw op" gpurun_out/pair_micro.log

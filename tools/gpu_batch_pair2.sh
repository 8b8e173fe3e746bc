cd $GRAFT_REPO_ROOT
rm -f gpurun_out/pair_micro.log
timeout 200 python tools/bench_mlp_tc.py --rays 4096 --samples 128 --iters 20 --pair 1 --prof >> gpurun_out/pair_micro.log 2>&1
timeout 200 python tools/bench_mlp_tc.py --rays 4096 --samples 128 --iters 20 --pair 1 --save --prof >> gpurun_out/pair_micro.log 2>&1
grep -v "^dw op" gpurun_out/pair_micro.log

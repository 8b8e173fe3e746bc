cd $GRAFT_REPO_ROOT
timeout 200 python tools/check_pair.py > gpurun_out/pair_check.log 2>&1; echo check rc=$?; tail -1 gpurun_out/pair_check.log
for sh in 0 1; do
echo "== SHARE=$sh"
DDNERF_TC_PAIR_SHARE=$sh timeout 200 python tools/bench_mlp_tc.py --rays 4096 --samples 128 --iters 20 --pair 1 --prof 2>&1 | grep -v "^dw op" | cut -c1-330
DDNERF_TC_PAIR_SHARE=$sh timeout 200 python tools/bench_mlp_tc.py --rays 4096 --samples 128 --iters 20 --pair 1 --save --prof 2>&1 | grep -v "^dw op" | cut -c1-450
done

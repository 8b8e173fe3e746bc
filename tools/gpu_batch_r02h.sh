set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r02h_rc.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; echo "bench rc=$?" >> gpurun_out/r02h_rc.txt
timeout 300 python tools/bench_mlp_tc.py --rays 4096 --samples 128 --iters 20 --save > gpurun_out/r02h_mlp_micro.log 2>&1; echo "micro rc=$?" >> gpurun_out/r02h_rc.txt
cat gpurun_out/r02h_rc.txt; tail -5 gpurun_out/r02h_pytest.log; tail -20 gpurun_out/r02h_mlp_micro.log

cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r02i_rc.txt
timeout 200 python tools/check_pair.py > gpurun_out/pair_check.log 2>&1; echo "check rc=$?" >> gpurun_out/r02i_rc.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; echo "bench rc=$?" >> gpurun_out/r02i_rc.txt
DDNERF_TC_PAIR=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02i_bench_single.json 2> gpurun_out/r02i_bench_single.err; echo "bench single rc=$?" >> gpurun_out/r02i_rc.txt
cat gpurun_out/r02i_rc.txt; tail -5 gpurun_out/r02i_pytest.log

set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -k "accumulation or train_tail" > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r02g_rc.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; echo "bench rc=$?" >> gpurun_out/r02g_rc.txt
# launch lists (ncu, cold-cache serialised): cfg2 (mip-NeRF) and cfg1 (DDNeRF) training steps, eager launches of 2 steps
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02g_launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-graph --no-render --no-cfg4 --no-cpu-baseline > gpurun_out/r02g_ncu_cfg2.log 2>&1; echo "ncu cfg2 rc=$?" >> gpurun_out/r02g_rc.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02g_launches_cfg1.csv python bench.py --steps 2 --warmup 3 --no-graph --no-render --no-cfg4 --no-cpu-baseline --workload cfg1 > gpurun_out/r02g_ncu_cfg1.log 2>&1; echo "ncu cfg1 rc=$?" >> gpurun_out/r02g_rc.txt
cat gpurun_out/r02g_rc.txt; tail -5 gpurun_out/r02g_pytest.log

cd $GRAFT_REPO_ROOT
TAG=${1:-r02m}
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench_bf16.json 2> gpurun_out/${TAG}_bench_bf16.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2>> gpurun_out/${TAG}_bench_bf16.err; echo "ref rc=$?"
timeout 300 python bench.py --workload cfg1 --steps 20 --warmup 5 --no-render --no-cfg4 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg1_n1.json 2>>gpurun_out/${TAG}_bench_bf16.err; echo "cfg1 rc=$?"

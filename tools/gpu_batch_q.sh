cd $GRAFT_REPO_ROOT
TAG=${1:-r02q}
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches_cfg2.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-render --no-cfg4 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches_cfg2.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/${TAG}_launches_cfg2.csv > gpurun_out/${TAG}_launches_cfg2_summary.md
head -5 gpurun_out/${TAG}_launches_cfg2_summary.md

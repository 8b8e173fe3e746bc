cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_tc.py -q -x 2>&1 | tail -2
for i in 1 2; do for l in base new; do if [ $l = base ]; then export DDNERF_B200_LIB=$PWD/ddnerf_b200/libddnerf_b200_base.so; else unset DDNERF_B200_LIB; fi; timeout 200 python tools/bench_mlp_tc.py --rays 4096 --samples 128 --iters 20 --pair 1 --save 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$l', {k:round(d[k],4) for k in ('fwd_ms','fwd_rays_ms','dx_ms','dw_ms')})"; done; done

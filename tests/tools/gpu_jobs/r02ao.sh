set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02ao; mkdir -p $O
timeout 95 python tests/tools/wide_sweep.py 14000 14900 100 > $O/wide_sweep_gpu_strict.log 2>&1
tail -n 3 $O/wide_sweep_gpu_strict.log

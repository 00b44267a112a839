set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02i; mkdir -p $O
timeout 200 python -m tests.tools.ab --workload cfg2 --set strip_decay_early=320,64 --set strips_async=0,1 --steps 20 --rounds 1 > $O/ab_strip_geom.jsonl 2>&1
timeout 200 python -m tests.tools.ab --workload cfg1 --set strips_async=0,1 --steps 20 --rounds 1 > $O/ab_cfg1.jsonl 2>&1
timeout 200 python -m tests.tools.ab --workload cfg2 --mode exact --set decim_threads=0 --steps 5 --rounds 1 > $O/ab_exact.jsonl 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $O/pytest_parity.log 2>&1; echo "pytest rc=$?" >> $O/pytest_parity.log
timeout 500 python tests/tools/wide_sweep.py 5000 5600 400 --more-segments > $O/wide_sweep_gpu_strict.log 2>&1
ls -la $O

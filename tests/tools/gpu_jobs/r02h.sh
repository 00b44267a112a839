set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02h; mkdir -p $O
timeout 500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $O/pytest_parity.log 2>&1; echo "pytest rc=$?" >> $O/pytest_parity.log
timeout 200 python -m tests.tools.ab --workload cfg2 --set strip_decay_early=320,64 --set strips_async=0,1,2 --steps 20 --rounds 1 > $O/ab_strip_geom.jsonl 2>&1
timeout 200 python -m tests.tools.ab --workload cfg2 --set late_mix=1,0 --steps 20 --rounds 1 > $O/ab_fold.jsonl 2>&1
timeout 200 python -m tests.tools.ab --workload cfg1 --set strips_async=0,1,2 --steps 20 --rounds 1 > $O/ab_cfg1.jsonl 2>&1
timeout 200 python -m tests.tools.ab --workload cfg2 --mode exact --set decim_threads=0 --steps 5 --rounds 1 > $O/ab_exact.jsonl 2>&1
ls -la $O

set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02f; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "taper or standin or fast_golden or batch" > $O/pytest_subset.log 2>&1; echo "pytest rc=$?" >> $O/pytest_subset.log
timeout 200 python -m tests.tools.ab --workload cfg2 --set fir_threads=256,128 --set strips_async=0,1 --steps 20 --rounds 2 > $O/ab_fir_threads.jsonl 2>&1
timeout 200 python -m tests.tools.ab --workload cfg1 --set fir_threads=256,128 --steps 20 --rounds 2 > $O/ab_fir_threads_cfg1.jsonl 2>&1
timeout 200 python -m tests.tools.ab --workload cfg2 --set strip_decay=320,256,192 --set strips_async=0,1 --steps 20 --rounds 1 > $O/ab_strip_decay.jsonl 2>&1
ls -la $O

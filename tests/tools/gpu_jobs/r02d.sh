set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02d; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fast or decimated or standin or tile" > $O/pytest_subset.log 2>&1; echo "pytest rc=$?" >> $O/pytest_subset.log
timeout 200 python -m tests.tools.ab --workload cfg2 --set strip_split=0,1 --set strips_async=0,1 --steps 20 --rounds 2 > $O/ab_strip_split.jsonl 2>&1
timeout 200 python -m tests.tools.ab --workload cfg1 --set strip_split=0,1 --set strips_async=0,1 --steps 20 --rounds 1 > $O/ab_strip_split_cfg1.jsonl 2>&1
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err
ls -la $O

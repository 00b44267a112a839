set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02k; mkdir -p $O
timeout 200 python -m tests.tools.ab --workload cfg2 --set strips_async=0,1,2 --set iir_l2_keep=70,100 --steps 20 --rounds 2 > $O/ab_cfg2.jsonl 2>&1
timeout 200 python -m tests.tools.ab --workload cfg1 --set strips_async=0,1,2 --steps 20 --rounds 1 > $O/ab_cfg1.jsonl 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "channel or cfg4 or fast or golden" > $O/pytest_subset.log 2>&1; echo "pytest rc=$?" >> $O/pytest_subset.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_cfg2.json 2> $O/bench_cfg2.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload cfg1 > $O/bench_cfg1.json 2> $O/bench_cfg1.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload cfg3 > $O/bench_cfg3.json 2> $O/bench_cfg3.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload cfg4 > $O/bench_cfg4.json 2> $O/bench_cfg4.err
ls -la $O

set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02m; mkdir -p $O
# Welch: packed complex adds + uniform fast-path loads (all variants), pruned accumulators (welch_prune 1),
# pruned + 3 CTAs/SM (welch_prune 2) against the unpruned kernel
timeout 200 python -m tests.tools.ab --workload cfg2 --set welch_prune=0,1,2 --steps 20 --rounds 2 > $O/ab_welch_cfg2.jsonl 2>&1
timeout 200 python -m tests.tools.ab --workload cfg1 --set welch_prune=0,1,2 --steps 20 --rounds 2 > $O/ab_welch_cfg1.jsonl 2>&1
timeout 200 python -m tests.tools.ab --workload cfg4 --set welch_prune=0,1,2 --steps 10 --rounds 2 > $O/ab_welch_cfg4.jsonl 2>&1
timeout 200 python tests/tools/sweep.py --no-cpu --ratios 1,4 --sizes 1024,2048,4096,8192 > $O/sweep_r1_r4.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $O/pytest_parity.log 2>&1; echo "pytest rc=$?" >> $O/pytest_parity.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_cfg2.json 2> $O/bench_cfg2.err
timeout 200 python tests/tools/wide_sweep.py 9000 9600 100 > $O/wide_sweep_gpu_strict.log 2>&1
ls -la $O

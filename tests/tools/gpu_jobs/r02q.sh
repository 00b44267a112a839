set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02q; mkdir -p $O
# who is on the critical path: strips (side stream) or FIR + last stage?  priority of the side stream,
# where the strips are submitted (1: first, 2: after the FIR chain, 3: after it AND waiting for it), tile depth
timeout 300 python -m tests.tools.ab --workload cfg2 --set strips_priority=0,1 --set strips_async=1,2,3 --set iir_depth=0,1 --steps 20 --rounds 2 > $O/ab_strips_sched_cfg2.jsonl 2>&1
timeout 300 python -m tests.tools.ab --workload cfg1 --set strips_priority=0,1 --set strips_async=1,3 --set iir_depth=0,1 --steps 20 --rounds 2 > $O/ab_strips_sched_cfg1.jsonl 2>&1
timeout 300 python -m tests.tools.ab --workload cfg4 --set strips_priority=0,1 --set strips_async=1,3 --set iir_depth=0,1 --steps 10 --rounds 1 > $O/ab_strips_sched_cfg4.jsonl 2>&1
# cfg3 without the mean pre-pass
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload cfg3 > $O/bench_cfg3.json 2> $O/bench_cfg3.err
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cfg3 or corner or sweep or big or 65536 or golden" > $O/pytest_big.log 2>&1; echo "pytest rc=$?" >> $O/pytest_big.log
timeout 300 python tests/tools/cfg3_job.py > $O/cfg3_full_job.json 2> $O/cfg3_full_job.err
timeout 300 python tests/tools/sweep.py --no-cpu --ratios 1,8 --sizes 65536,131072 > $O/sweep_big.log 2>&1
ls -la $O

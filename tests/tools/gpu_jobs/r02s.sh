set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02s; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cfg3 or corner or big or golden or dc" > $O/pytest_big.log 2>&1; echo "pytest rc=$?" >> $O/pytest_big.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload cfg3 > $O/bench_cfg3.json 2> $O/bench_cfg3.err
timeout 300 python tests/tools/cfg3_job.py > $O/cfg3_full_job.json 2> $O/cfg3_full_job.err
timeout 200 python -m tests.tools.ab --workload cfg1 --set welch_prune=0,1,2 --steps 20 --rounds 2 > $O/ab_welch_prune_cfg1.jsonl 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload cfg1 > $O/bench_cfg1.json 2> $O/bench_cfg1.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload cfg1 --group 512 > $O/bench_cfg1_group512.json 2> $O/bench_cfg1_group512.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_cfg2.json 2> $O/bench_cfg2.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --group 1024 --frames 1024 > $O/bench_cfg2_group1024.json 2> $O/bench_cfg2_group1024.err
# cfg5 sweep with the fp64 few-segment path switched off (round-1 behaviour) beside the default of r02r
timeout 600 python tests/tools/sweep.py --no-cpu --set precise=0 --out $O/sweep_cfg5_precise0.jsonl > $O/sweep_cfg5_precise0.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"big_|welch_kernel" -c 10 -o $O/prof_cfg3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --sustain-s 0 --workload cfg3 --e2e-steps 1 > $O/ncu_cfg3.log 2>&1
ls -la $O

set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02r; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cfg3 or corner or big or golden or dc" > $O/pytest_big.log 2>&1; echo "pytest rc=$?" >> $O/pytest_big.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload cfg3 > $O/bench_cfg3.json 2> $O/bench_cfg3.err
timeout 300 python tests/tools/cfg3_job.py > $O/cfg3_full_job.json 2> $O/cfg3_full_job.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"big_|welch_kernel" -c 10 -o $O/prof_cfg3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --sustain-s 0 --workload cfg3 --e2e-steps 1 > $O/ncu_cfg3.log 2>&1
# BASELINE configs[4]: the decimation / FFT-size sweep with this round's kernels (CPU oracle beside it)
timeout 900 python tests/tools/sweep.py --out $O/sweep_cfg5.jsonl > $O/sweep_cfg5.log 2>&1
timeout 300 python tests/tools/wide_sweep.py 11000 11600 120 > $O/wide_sweep_gpu_strict.log 2>&1
ls -la $O

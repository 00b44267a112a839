set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02v; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cs16 or golden or fast or error" > $O/pytest_cs16.log 2>&1; echo "pytest rc=$?" >> $O/pytest_cs16.log
timeout 300 python -m tests.tools.ab --workload cfg2cs16 --set cs16_fused=0,1 --steps 20 --rounds 2 > $O/ab_cs16_fused.jsonl 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload cfg2cs16 > $O/bench_cfg2cs16.json 2> $O/bench_cfg2cs16.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_cfg2.json 2> $O/bench_cfg2.err
timeout 300 python tests/tools/wide_sweep.py 13000 13900 150 > $O/wide_sweep_gpu_strict.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"fir_run" -s 2 -c 1 -o /tmp/prof_cs16 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sustain-s 0 --workload cfg2cs16 --e2e-steps 1 > $O/ncu_cs16.log 2>&1
python tests/tools/ncu_summary.py /tmp/prof_cs16.ncu-rep "r02v: ncu --set full, fir_run_kernel<KIND_CS16_RAW> (int16 IQ converted on load), bench.py --workload cfg2cs16" > $O/ncu_full_fir_run_cs16.txt 2>&1
ls -la $O

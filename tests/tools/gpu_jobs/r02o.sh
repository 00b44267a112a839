set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02o; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "waterfall or level or display or autolevel or image" > $O/pytest_display.log 2>&1; echo "pytest rc=$?" >> $O/pytest_display.log
timeout 300 python tests/tools/image_bench.py --widths 1024,32768 --no-cpu > $O/image_bench.jsonl 2>&1
timeout 200 python -m tests.tools.ab --workload cfg2 --set welch_splits=1,2,4,8 --steps 20 --rounds 2 > $O/ab_welch_splits_cfg2.jsonl 2>&1
timeout 200 python -m tests.tools.ab --workload cfg1 --set welch_splits=1,2,4,8 --steps 20 --rounds 2 > $O/ab_welch_splits_cfg1.jsonl 2>&1
# launch list of the device-resident leg (no skip: warm-up + timed steps), same command as the bench
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sustain-s 0 > $O/bench_plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 54 --csv --log-file $O/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sustain-s 0 > $O/ncu_launches.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload cfg1 > $O/bench_cfg1.json 2> $O/bench_cfg1.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload cfg3 > $O/bench_cfg3.json 2> $O/bench_cfg3.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload cfg4 > $O/bench_cfg4.json 2> $O/bench_cfg4.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"wf_select|big_r16|welch_kernel|big_gather|big_halfsum" -c 8 -o $O/prof_cfg3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --sustain-s 0 --workload cfg3 --e2e-steps 1 > $O/ncu_cfg3.log 2>&1
ls -la $O

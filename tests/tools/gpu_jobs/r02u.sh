set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02u; mkdir -p $O
# the whole GPU suite on the candidate final build
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench_cfg2.json 2> $O/bench_cfg2.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err
for w in cfg1 cfg3 cfg4 cfg2cs16; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload $w > $O/bench_$w.json 2> $O/bench_$w.err
done
timeout 300 python tests/tools/cfg3_job.py > $O/cfg3_full_job.json 2> $O/cfg3_full_job.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 54 --csv --log-file $O/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sustain-s 0 > $O/ncu_launches.log 2>&1
# (.ncu-rep files are summarised here and left on the box: two of them exceed the 64 MiB that come back)
timeout 900 ncu --set full --clock-control none -k regex:"fir_run|strip_cascade|iir_stream|welch_kernel|ema_rows" -s 16 -c 7 -o /tmp/prof_cfg2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sustain-s 0 > $O/ncu_cfg2.log 2>&1
python tests/tools/ncu_summary.py /tmp/prof_cfg2.ncu-rep "r02u: ncu --set full --clock-control none, every kernel of one cfg2 step (launches 16..22 of bench.py --steps 3 --warmup 3), final build of round 2" > $O/ncu_full_cfg2_fast.txt 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"big_|welch_kernel" -c 6 -o /tmp/prof_cfg3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --sustain-s 0 --workload cfg3 --e2e-steps 1 > $O/ncu_cfg3.log 2>&1
python tests/tools/ncu_summary.py /tmp/prof_cfg3.ncu-rep "r02u: ncu --set full --clock-control none, bench.py --workload cfg3 (N = 65536), final build of round 2" > $O/ncu_full_cfg3.txt 2>&1
# BASELINE configs[4] sweep with the batched fp64 few-segment path (default policy)
timeout 900 python tests/tools/sweep.py --no-cpu --out $O/sweep_cfg5.jsonl > $O/sweep_cfg5.log 2>&1
timeout 200 python tests/tools/wide_sweep.py 12000 12500 100 > $O/wide_sweep_gpu_strict.log 2>&1
ls -la $O

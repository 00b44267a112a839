set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02l; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
# A/B of the Welch staging variant (cp.async into shared memory) against the plain loop
timeout 200 python -m tests.tools.ab --workload cfg2 --set welch_stage=0,1 --steps 20 --rounds 2 > $O/ab_welch_cfg2.jsonl 2>&1
timeout 200 python -m tests.tools.ab --workload cfg1 --set welch_stage=0,1 --steps 20 --rounds 2 > $O/ab_welch_cfg1.jsonl 2>&1
timeout 200 python tests/tools/sweep.py --no-cpu --ratios 1 --sizes 2048,4096,8192 --set welch_stage=0 > $O/sweep_r1_stage0.log 2>&1
timeout 200 python tests/tools/sweep.py --no-cpu --ratios 1 --sizes 2048,4096,8192 --set welch_stage=1 > $O/sweep_r1_stage1.log 2>&1
# the whole GPU suite on this build
timeout 1200 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
# bench lines (driver contract), reference arm
timeout 600 python bench.py > $O/bench_cfg2.json 2> $O/bench_cfg2.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload cfg1 > $O/bench_cfg1.json 2> $O/bench_cfg1.err
# launch list + full counters of every kernel of a cfg2 step
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sustain-s 0 > $O/bench_plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file $O/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sustain-s 0 > $O/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fir_run|strip_cascade|iir_stream|welch_kernel|ema_rows" -s 16 -c 7 -o $O/prof_all python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sustain-s 0 > $O/ncu_all.log 2>&1
# strict-floor wide sweep on the GPU (verdict item 6: >= 1500 configurations)
timeout 400 python tests/tools/wide_sweep.py 7000 8700 330 > $O/wide_sweep_gpu_strict.log 2>&1
ls -la $O

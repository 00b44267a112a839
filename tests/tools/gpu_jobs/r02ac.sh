set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02ac; mkdir -p $O
# pipelined batches (zfb_join): rows bit-identical?  what does it buy at 1 GPU?
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "slabs or pipeline or ema_batch or device_path or ema_rows" > $O/pytest_pipeline.log 2>&1
echo "pytest rc=$?" >> $O/pytest_pipeline.log
B="--steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 4 --sustain-s 2"
for p in 0 1; do
  timeout 200 python bench.py $B --pipeline $p > $O/bench_cfg2_pipe$p.json 2>> $O/bench.err
  timeout 200 python bench.py $B --pipeline $p --workload cfg1 > $O/bench_cfg1_pipe$p.json 2>> $O/bench.err
  timeout 200 python bench.py $B --pipeline $p --workload cfg2cs16 > $O/bench_cfg2cs16_pipe$p.json 2>> $O/bench.err
  timeout 200 python bench.py $B --pipeline $p --frames 128 > $O/bench_cfg2_f128_pipe$p.json 2>> $O/bench.err
  timeout 200 python bench.py $B --pipeline $p --frames 32 > $O/bench_cfg2_f32_pipe$p.json 2>> $O/bench.err
done
tail -5 $O/pytest_pipeline.log
ls -la $O

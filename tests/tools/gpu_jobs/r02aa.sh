set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02aa; mkdir -p $O
# slab pipelining inside zfb_process_device: rows bit-identical to one lane?  what does it buy?
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "slabs or ema_batch or device_path or ema_rows" > $O/pytest_slabs.log 2>&1
echo "pytest rc=$?" >> $O/pytest_slabs.log
B="--steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 4"
for s in 1 2; do
  timeout 200 python bench.py $B --sustain-s 2 --slabs $s > $O/bench_cfg2_slabs$s.json 2>> $O/bench.err
  timeout 200 python bench.py $B --sustain-s 2 --slabs $s --workload cfg1 > $O/bench_cfg1_slabs$s.json 2>> $O/bench.err
  timeout 200 python bench.py $B --sustain-s 2 --slabs $s --workload cfg1 --set fir_smem_pad=24576 > $O/bench_cfg1_pad_slabs$s.json 2>> $O/bench.err
  timeout 200 python bench.py $B --sustain-s 0 --slabs $s --frames 128 > $O/bench_cfg2_f128_slabs$s.json 2>> $O/bench.err
  timeout 200 python bench.py $B --sustain-s 0 --slabs $s --frames 64 > $O/bench_cfg2_f64_slabs$s.json 2>> $O/bench.err
  timeout 200 python bench.py $B --sustain-s 0 --slabs $s --workload cfg2cs16 > $O/bench_cfg2cs16_slabs$s.json 2>> $O/bench.err
done
timeout 200 python bench.py $B --sustain-s 0 --slabs 2 --frames 1024 > $O/bench_cfg2_f1024_slabs2.json 2>> $O/bench.err
tail -5 $O/pytest_slabs.log
ls -la $O

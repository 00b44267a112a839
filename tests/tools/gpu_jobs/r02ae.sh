set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02ae; mkdir -p $O
# virtual receivers (cfg4) through pipelined batches
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipeline or multi_channel or cfg4" > $O/pytest_pipeline_channels.log 2>&1
echo "pytest rc=$?" >> $O/pytest_pipeline_channels.log
B="--steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 4 --sustain-s 2 --workload cfg4"
for p in 0 1; do
  timeout 200 python bench.py $B --pipeline $p > $O/bench_cfg4_pipe$p.json 2>> $O/bench.err
  timeout 200 python bench.py $B --pipeline $p --frames 2 > $O/bench_cfg4_f2_pipe$p.json 2>> $O/bench.err
done
tail -n 5 $O/pytest_pipeline_channels.log

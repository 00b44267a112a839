set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02al; mkdir -p $O
# host batches ending in ever smaller sub-groups: what is left of e2e after the last copy
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "device_path or batch_equals or ema_rows or ema_batch" > $O/pytest_host_taper.log 2>&1
echo "pytest rc=$?" >> $O/pytest_host_taper.log
tail -n 4 $O/pytest_host_taper.log
B="--steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 40 --sustain-s 0"
for t in 0 1 0 1; do
  echo "{\"host_taper\": $t}" >> $O/ab_host_taper.jsonl
  timeout 200 python bench.py $B --set host_taper=$t >> $O/ab_host_taper.jsonl 2>> $O/bench.err
done
tail -n 3 $O/bench.err

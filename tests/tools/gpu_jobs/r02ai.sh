set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02ai; mkdir -p $O
# N = 65536 in one pass over a 16-CTA cluster (DSMEM) against the two-kernel path through the scratch buffer
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "big_cluster or big_fft or 65536 or cfg3" > $O/pytest_big_cluster.log 2>&1
echo "pytest rc=$?" >> $O/pytest_big_cluster.log
tail -n 15 $O/pytest_big_cluster.log
B="--steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 2 --sustain-s 2 --workload cfg3"
for c in 0 1; do
  timeout 200 python bench.py $B --set big_cluster=$c > $O/bench_cfg3_cluster$c.json 2>> $O/bench.err
done
timeout 200 python bench.py $B --set big_cluster=1 --welch-splits 2 > $O/bench_cfg3_cluster1_splits2.json 2>> $O/bench.err
timeout 200 python bench.py $B --set big_cluster=1 --welch-splits 1 > $O/bench_cfg3_cluster1_splits1.json 2>> $O/bench.err
tail -n 5 $O/bench.err

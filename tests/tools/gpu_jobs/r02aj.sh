set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02aj; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "big_cluster" > $O/pytest_big_cluster.log 2>&1
echo "pytest rc=$?" >> $O/pytest_big_cluster.log
tail -n 5 $O/pytest_big_cluster.log
B="--steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 2 --sustain-s 1.5 --workload cfg3"
for c in 2 1 0; do
  timeout 200 python bench.py $B --set big_cluster=$c > $O/bench_cfg3_cluster$c.json 2>> $O/bench.err
done
tail -n 5 $O/bench.err

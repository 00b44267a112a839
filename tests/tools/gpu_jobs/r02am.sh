set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02am; mkdir -p $O
# the whole GPU suite + the driver's commands on the final build of the round (pipelined batches on)
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 600 python bench.py > $O/bench_cfg2.json 2> $O/bench_cfg2.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err
for w in cfg1 cfg3 cfg4 cfg2cs16; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload $w > $O/bench_$w.json 2> $O/bench_$w.err
done
ls -la $O

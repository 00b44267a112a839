set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02ab; mkdir -p $O
# slabs small enough that a lane's intermediates stay in L2 (group 128: 41 MB of FIR output per slab),
# the other lane filling the tails of the short launches
B="--steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 2 --sustain-s 1.5"
for sg in "1 128" "2 128" "2 64" "2 256" "1 256"; do
  set -- $sg
  timeout 200 python bench.py $B --slabs $1 --group $2 > $O/bench_cfg2_slabs$1_g$2.json 2>> $O/bench.err
done
timeout 200 python bench.py $B --slabs 2 --group 128 --frames 2048 > $O/bench_cfg2_f2048_slabs2_g128.json 2>> $O/bench.err
timeout 200 python bench.py $B --slabs 1 --frames 2048 > $O/bench_cfg2_f2048_slabs1.json 2>> $O/bench.err
timeout 200 python bench.py $B --slabs 2 --group 128 --workload cfg1 > $O/bench_cfg1_slabs2_g128.json 2>> $O/bench.err
ls -la $O

set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02an; mkdir -p $O
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline --sustain-s 1.5 --e2e-steps 10 > $O/bench_n4.json 2> $O/bench_n4.err; echo "rc=$?" >> $O/bench_n4.err
tail -n 3 $O/bench_n4.err

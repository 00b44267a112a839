set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02w; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
# the driver's scaling launch at N = 8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err; echo "rc=$?" >> $O/bench_n8.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 --workload cfg4 > $O/bench_cfg4_n8.json 2> $O/bench_cfg4_n8.err; echo "rc=$?" >> $O/bench_cfg4_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 20 --warmup 5 > $O/bench_n4.json 2> $O/bench_n4.err; echo "rc=$?" >> $O/bench_n4.err
ls -la $O

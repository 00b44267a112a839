set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02n; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
# the driver's scaling launch at N = 2 (sustained leg included): must end, must verify bit-exact rows
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo "rc=$?" >> $O/bench_n2.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --workload cfg4 > $O/bench_cfg4_n2.json 2> $O/bench_cfg4_n2.err; echo "rc=$?" >> $O/bench_cfg4_n2.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 1 --impl reference > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err; echo "rc=$?" >> $O/bench_ref_n2.err
timeout 900 python -m pytest tests/test_multigpu.py -x -q -m gpu > $O/pytest_multigpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_multigpu.log
ls -la $O

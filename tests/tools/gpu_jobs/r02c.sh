set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02c; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
python bench.py --steps 20 --warmup 5 --workload cfg4 --no-cpu-baseline --sustain-s 1 > $O/bench_cfg4_n1.json 2> $O/bench_cfg4_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --workload cfg4 --sustain-s 1 > $O/bench_cfg4_n2.json 2> $O/bench_cfg4_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --workload cfg4 --sustain-s 1 --cfg4-feed replicate > $O/bench_cfg4_n2_replicate.json 2> $O/bench_cfg4_n2_replicate.err
timeout 900 python -m pytest tests/test_multigpu.py -x -q -m gpu > $O/pytest_multigpu.log 2>&1
tail -c 600 $O/*.err
ls -la $O

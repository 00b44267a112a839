set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02y; mkdir -p $O
# cfg4 on 8 GPUs: every rank uploads 1/8 of the stream + all-gather over NVLink, against rank 0 uploading all of it + broadcast
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 --workload cfg4 --sustain-s 2 > $O/bench_cfg4_n8_allgather.json 2> $O/bench_cfg4_n8_allgather.err; echo "rc=$?" >> $O/bench_cfg4_n8_allgather.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 10 --warmup 3 --workload cfg4 --sustain-s 2 --cfg4-feed broadcast > $O/bench_cfg4_n8_broadcast.json 2> $O/bench_cfg4_n8_broadcast.err; echo "rc=$?" >> $O/bench_cfg4_n8_broadcast.err
ls -la $O

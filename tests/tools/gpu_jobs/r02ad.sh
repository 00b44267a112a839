set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02ad; mkdir -p $O
# 2 GPUs, the driver's launch line: pipelined batches with the NCCL gather ordered by zfb_join(comm stream)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --sustain-s 2 > $O/bench_n2.json 2> $O/bench_n2.err; echo "rc=$?" >> $O/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --sustain-s 2 --pipeline 0 > $O/bench_n2_plain.json 2> $O/bench_n2_plain.err; echo "rc=$?" >> $O/bench_n2_plain.err
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q > $O/pytest_multigpu.log 2>&1; echo "rc=$?" >> $O/pytest_multigpu.log
tail -3 $O/*.err $O/pytest_multigpu.log

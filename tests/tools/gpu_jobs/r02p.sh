set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02p; mkdir -p $O
# deeper tile pipelines in the streaming last stage (3+3 / 4+4 / 5+4 tiles in flight per warp)
timeout 200 python -m tests.tools.ab --workload cfg2 --set iir_depth=0,1,2 --steps 20 --rounds 2 > $O/ab_iir_depth_cfg2.jsonl 2>&1
timeout 200 python -m tests.tools.ab --workload cfg2 --set iir_depth=0,1,2 --set strips_async=0 --steps 20 --rounds 1 > $O/ab_iir_depth_cfg2_serial.jsonl 2>&1
timeout 200 python -m tests.tools.ab --workload cfg1 --set iir_depth=0,1,2 --steps 20 --rounds 2 > $O/ab_iir_depth_cfg1.jsonl 2>&1
# slab concurrency probe for cfg3: front pass (DRAM-bound) of one slab beside the FFT pass of another?
timeout 300 python -m tests.tools.concurrency_probe --workload cfg3 --frames 64 --steps 20 --engines 1,2,4 > $O/probe_cfg3.jsonl 2>&1
timeout 300 python -m tests.tools.concurrency_probe --workload cfg3 --frames 64 --steps 20 --engines 2,4 --priority 1 > $O/probe_cfg3_prio.jsonl 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fast or golden or cfg4 or channel" > $O/pytest_subset.log 2>&1; echo "pytest rc=$?" >> $O/pytest_subset.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_cfg2.json 2> $O/bench_cfg2.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload cfg1 > $O/bench_cfg1.json 2> $O/bench_cfg1.err
ls -la $O

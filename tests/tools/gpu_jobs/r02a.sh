set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a/smi.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or fast or cfg4 or batch" > gpurun_out/r02a/pytest_subset.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a/pytest_subset.log
timeout 300 python -m tests.tools.ab --workload cfg2 --set iir_stream=0,1 --steps 20 --rounds 2 > gpurun_out/r02a/ab_stream_onoff.jsonl 2>&1
timeout 300 python -m tests.tools.ab --workload cfg2 --set iir_stream_len=256,544,1088,2048 --steps 20 --rounds 2 > gpurun_out/r02a/ab_stream_len.jsonl 2>&1
timeout 300 python -m tests.tools.ab --workload cfg2 --set iir_stream_warm=128,192,256 --set strips_async=1,0 --steps 20 --rounds 1 > gpurun_out/r02a/ab_stream_warm.jsonl 2>&1
timeout 300 python -m tests.tools.ab --workload cfg1 --set iir_stream=0,1 --steps 20 --rounds 2 > gpurun_out/r02a/ab_cfg1.jsonl 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02a/bench_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:iir_stream -s 4 -c 2 -o gpurun_out/r02a/prof_iir python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02a/ncu_iir.log 2>&1
ls -la gpurun_out/r02a

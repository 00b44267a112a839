set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02b; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or fast or cfg4 or batch or standin or survives or ring_wrap" > $O/pytest_subset.log 2>&1; echo "pytest rc=$?" >> $O/pytest_subset.log
timeout 300 python -m tests.tools.ab --workload cfg2 --set iir_stream=0,1 --set strips_async=1,0 --steps 20 --rounds 2 > $O/ab_stream_onoff.jsonl 2>&1
timeout 300 python -m tests.tools.ab --workload cfg2 --set iir_stream_len=320,416,640,1248 --set strips_async=0 --steps 20 --rounds 1 > $O/ab_stream_len.jsonl 2>&1
timeout 300 python -m tests.tools.ab --workload cfg2 --set iir_stream_warm=128,192,256 --set strips_async=0 --steps 20 --rounds 1 > $O/ab_stream_warm.jsonl 2>&1
timeout 300 python -m tests.tools.ab --workload cfg1 --set iir_stream=0,1 --set strips_async=1,0 --steps 20 --rounds 1 > $O/ab_cfg1.jsonl 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:iir_stream -s 4 -c 2 -o $O/prof_iir python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_iir.log 2>&1
ls -la $O

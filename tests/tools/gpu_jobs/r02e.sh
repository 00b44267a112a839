set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02e; mkdir -p $O
timeout 200 python -m tests.tools.concurrency_probe --engines 1,2,4 > $O/probe.jsonl 2>&1
timeout 200 python -m tests.tools.concurrency_probe --engines 2,4 --priority 1 > $O/probe_prio.jsonl 2>&1
timeout 200 python -m tests.tools.ab --workload cfg2 --set iir_l2_keep=0,30,50,70 --set strips_async=0 --steps 20 --rounds 2 > $O/ab_l2keep.jsonl 2>&1
timeout 200 python -m tests.tools.ab --workload cfg2 --set iir_stream_warm=192,256 --set iir_stream_len=640,1248 --steps 20 --rounds 1 > $O/ab_warm.jsonl 2>&1
ls -la $O

set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02z; mkdir -p $O
# does the step gain from slabs in flight when the FIR run kernel leaves room on the SM (3 or 2 CTAs/SM
# through extra dynamic shared memory) for the latency-bound kernels of the other slab?
for pad in 0 24576 43008; do
  timeout 200 python -m tests.tools.concurrency_probe --workload cfg2 --frames 512 --steps 30 --engines 1,2,4 --set fir_smem_pad=$pad >> $O/probe_cfg2_pad.jsonl 2>&1
done
timeout 200 python -m tests.tools.concurrency_probe --workload cfg2 --frames 1024 --steps 20 --engines 2,4 --set fir_smem_pad=24576 >> $O/probe_cfg2_pad_1024.jsonl 2>&1
timeout 200 python -m tests.tools.concurrency_probe --workload cfg2 --frames 1024 --steps 20 --engines 1,2 --set fir_smem_pad=0 >> $O/probe_cfg2_pad_1024.jsonl 2>&1
timeout 200 python -m tests.tools.concurrency_probe --workload cfg1 --frames 512 --steps 20 --engines 1,2 --set fir_smem_pad=0 >> $O/probe_cfg1_pad.jsonl 2>&1
timeout 200 python -m tests.tools.concurrency_probe --workload cfg1 --frames 512 --steps 20 --engines 1,2 --set fir_smem_pad=24576 >> $O/probe_cfg1_pad.jsonl 2>&1
ls -la $O

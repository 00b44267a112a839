set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02ak; mkdir -p $O
for c in 1 2; do
timeout 400 ncu --set full --clock-control none -k regex:"big_cluster" -s 2 -c 1 -o /tmp/prof_cluster$c python bench.py --steps 1 --warmup 3 --no-cpu-baseline --sustain-s 0 --workload cfg3 --e2e-steps 1 --set big_cluster=$c > $O/ncu_cluster$c.log 2>&1
python tests/tools/ncu_summary.py /tmp/prof_cluster$c.ncu-rep "r02ak: ncu --set full --clock-control none, big_cluster_kernel (big_cluster=$c), bench.py --workload cfg3" > $O/ncu_full_big_cluster$c.txt 2>&1
done
ls -la $O /tmp/*.ncu-rep

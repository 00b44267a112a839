set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02j; mkdir -p $O
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sustain-s 0 > $O/bench_plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file $O/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sustain-s 0 > $O/ncu_launches.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sustain-s 0 > $O/bench_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fir_run|strip_cascade|iir_stream|welch_kernel" -s 12 -c 8 -o $O/prof_all python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sustain-s 0 > $O/ncu_all.log 2>&1
ls -la $O

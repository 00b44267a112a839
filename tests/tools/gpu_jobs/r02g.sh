set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02g; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fast or decimated or golden or random or sweep or cfg4" > $O/pytest_subset.log 2>&1; echo "pytest rc=$?" >> $O/pytest_subset.log
timeout 200 python -m tests.tools.ab --workload cfg2 --set strip_decay=320,192 --set strip_decay_early=320,64 --set strips_async=0,1 --steps 20 --rounds 1 > $O/ab_strip_geom.jsonl 2>&1
timeout 200 python -m tests.tools.ab --workload cfg1 --set strip_decay_early=320,64 --set strips_async=0,1 --steps 20 --rounds 1 > $O/ab_strip_geom_cfg1.jsonl 2>&1
ls -la $O

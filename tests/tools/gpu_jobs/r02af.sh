set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02af; mkdir -p $O
# measurement knobs again, now with pipelined batches (the other lane fills what the knobs used to fight for)
B="--steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 2 --sustain-s 2 --pipeline 1"
i=0
for opts in "" "--strips-async 0" "--strips-async 2" "--strips-async 3" "--set iir_depth=1" "--set iir_depth=2" \
            "--set fir_smem_pad=24576" "--set iir_depth=1 --strips-async 3" "--set strips_priority=1" "--welch-splits 2" "--welch-splits 8"; do
  echo "{\"variant\": \"$opts\"}" >> $O/ab_pipeline_knobs.jsonl
  timeout 200 python bench.py $B $opts >> $O/ab_pipeline_knobs.jsonl 2>> $O/bench.err
done
ls -la $O

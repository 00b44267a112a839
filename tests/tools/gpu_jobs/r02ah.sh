set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02ah; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
cat $O/smoke.log

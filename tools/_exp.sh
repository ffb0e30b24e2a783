cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r5b.log
: > $L
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^$" | tail -12 >> $L
timeout 300 python tools/stage_times.py C3 C4 C5 >> $L 2>&1
cat $L

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r5h.log
: > $L
timeout 300 python tools/stage_times.py C3 C4 C5 >> $L 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 >> $L
cat $L

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r5a.log
: > $L
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q 2>&1 | grep -v "^$" | tail -40 >> $L
timeout 300 python tools/stage_times.py C3 C4 C5 >> $L 2>&1
cat $L

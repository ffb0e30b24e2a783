cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r5k.log
: > $L
timeout 300 python tools/stage_times.py C3 C4 C5 >> $L 2>&1
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_densify.py -x -q -m gpu 2>&1 | tail -3 >> $L
timeout 300 python tools/knn_bench.py >> $L 2>&1
cat $L

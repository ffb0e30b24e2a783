cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/config_sweep.py > gpurun_out/r4n_config_sweep.jsonl 2> gpurun_out/r4n_config_sweep.err
for c in C4 C2 C3; do timeout 300 python tools/loop_bench.py --config $c >> gpurun_out/r4n_loop_skeleton.jsonl 2>> gpurun_out/r4n_loop.err; done
cat gpurun_out/r4n_config_sweep.jsonl gpurun_out/r4n_loop_skeleton.jsonl | cut -c1-400

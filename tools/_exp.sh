# scratch script for one-off gpurun calls:  gpurun -- 'bash tools/_exp.sh'
# NOTE: keep what is written under gpurun_out/ below 64 MiB or nothing is copied back (a --set full capture of a whole
# iteration with --import-source on is ~70 MiB: capture a kernel subset with -k, as tools/one_step.py's docstring shows)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py > gpurun_out/bench_new.json 2> gpurun_out/bench_new.err
cat gpurun_out/bench_reference.json gpurun_out/bench_new.json

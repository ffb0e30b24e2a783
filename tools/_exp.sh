cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --impl reference > gpurun_out/r5l_bench_reference.json 2> gpurun_out/r5l_bench_reference.err
python bench.py > gpurun_out/r5l_bench_new.json 2> gpurun_out/r5l_bench_new.err
timeout 300 python tools/stage_times.py C3 C4 C5 > gpurun_out/r5l_stage_times.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r5l_launches.csv python bench.py --views-per-gpu 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r5l_ncu_bench.log 2>&1
cut -c1-300 gpurun_out/r5l_bench_new.json

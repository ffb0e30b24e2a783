set -x
run() { # N config steps
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500+$1)) bench.py --gpus $1 --config $2 --steps $3 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench_$2_n$1.json 2> gpurun_out/r2d_bench_$2_n$1.err || tail -5 gpurun_out/r2d_bench_$2_n$1.err
}
run 8 C5 5
run 8 C4 10
run 4 C4 10
run 2 C4 10
python bench.py --config C4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench_C4_n1.json 2> gpurun_out/r2d_bench_C4_n1.err
python bench.py --config C5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench_C5_n1.json 2> gpurun_out/r2d_bench_C5_n1.err
for f in gpurun_out/r2d_bench_*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print('$f', d['n_gpus'], round(d['value'],1), d['unit'], round(d['ms_per_iter'],3), 'e2e', round(d['e2e']['value'],1))
"; done

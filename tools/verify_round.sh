# One-box verification sequence at the end of a session: GPU tests, smoke(), both bench arms, the loop configs, the config
# sweep, then (only after everything above exited 0 without a profiler) the launch list and the ncu --set full capture of the
# render kernels.  Usage on the GPU box:  bash tools/verify_round.sh r8z
T=${1:-rXX}
O=gpurun_out
set -x
python -m pytest tests -x -q -m gpu > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${T}_pytest.log; tail -2 $O/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/${T}_smoke.log; tail -2 $O/${T}_smoke.log
python bench.py --impl reference > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err; echo "ref rc=$?"
python bench.py > $O/${T}_bench_new.json 2> $O/${T}_bench_new.err; echo "new rc=$?"
python bench.py --config C4-loop > $O/${T}_c4loop.json 2> $O/${T}_c4loop.err; echo "c4loop rc=$?"
python bench.py --config C5-stress --steps 3 > $O/${T}_c5.json 2> $O/${T}_c5.err; echo "c5 rc=$?"
python tools/config_sweep.py > $O/${T}_config_sweep.jsonl 2> $O/${T}_config_sweep.err; echo "sweep rc=$?"
python tools/stage_times.py C1 C2 C3 C4 C5 > $O/${T}_stage_times.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv python tools/profile_iter.py C3 3 > $O/${T}_ncu_launches.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:render_ -s 2 -c 2 -f -o $O/${T}_render python tools/one_step.py --iters 2 > $O/${T}_ncu_full.log 2>&1; echo "ncu full rc=$?"
for f in $O/${T}_bench_reference.json $O/${T}_bench_new.json $O/${T}_c4loop.json $O/${T}_c5.json; do python -c "
import json
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1])
    print('$f', round(d['value'],1), d['unit'], 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'roofline', (d.get('roofline') or {}).get('frac'), 'traffic', (d.get('roofline') or {}).get('traffic'))
except Exception as e: print('$f', 'NO LINE', e)
"; done

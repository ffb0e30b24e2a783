cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r5m_variants.log
: > $L
VD=$PWD/langscene-x_b200/csrc/build/variants
for v in 0 1 2 4 5 6; do
  if [ $v = 0 ]; then unset LSX_B200_LIB; else export LSX_B200_LIB=$VD/liblsx_v$v.so; fi
  echo "VARIANT $v" >> $L
  timeout 120 python tools/stage_times.py C3 >> $L 2>&1
done
unset LSX_B200_LIB
# candidates ordered by render_bwd time; the first one whose parity tests pass is the winner
order=$(python - <<'PY'
import json
best = []
v = None
for line in open("gpurun_out/r5m_variants.log"):
    if line.startswith("VARIANT"):
        v = line.split()[1]
    elif line.startswith("{") and v is not None:
        d = json.loads(line)
        best.append((d["stages_ms"]["render_bwd"] + d["stages_ms"]["preprocess_bwd"], v))
        v = None
print(" ".join(x[1] for x in sorted(best)))
PY
)
echo "ORDER $order" >> $L
win=0
for v in $order; do
  if [ $v = 0 ]; then unset LSX_B200_LIB; else export LSX_B200_LIB=$VD/liblsx_v$v.so; fi
  echo "PARITY VARIANT $v" >> $L
  if timeout 200 python -m pytest tests/test_parity_gpu.py -x -q -m gpu > gpurun_out/r5m_parity_v$v.log 2>&1; then
    tail -2 gpurun_out/r5m_parity_v$v.log >> $L
    win=$v
    break
  fi
  tail -15 gpurun_out/r5m_parity_v$v.log >> $L
done
echo "WINNER $win" >> $L
if [ $win = 0 ]; then unset LSX_B200_LIB; else export LSX_B200_LIB=$VD/liblsx_v$win.so; fi
timeout 200 python bench.py > gpurun_out/r5m_bench_new.json 2> gpurun_out/r5m_bench_new.err
timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r5m_tests.log
cat gpurun_out/r5m_tests.log >> $L
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 > gpurun_out/r5m_smoke.log
timeout 120 python tools/stage_times.py C3 C4 C5 > gpurun_out/r5m_stage_times.jsonl 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:render_ -s 2 -c 2 -f -o gpurun_out/r5m_render python tools/one_step.py --iters 2 > gpurun_out/r5m_ncu.log 2>&1
cat $L
cut -c1-400 gpurun_out/r5m_bench_new.json

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r4j.log
: > $L
for t in 128 64 256; do
  lib=$GRAFT_REPO_ROOT/langscene-x_b200/lsx_b200/liblsx_b200_t$t.so
  [ $t = 128 ] && lib=$GRAFT_REPO_ROOT/langscene-x_b200/lsx_b200/liblsx_b200.so
  echo "PRE_THREADS $t" >> $L
  LSX_B200_LIB=$lib timeout 300 python tools/stage_times.py C3 C5 >> $L 2>&1
done
echo "PARITY t64" >> $L
LSX_B200_LIB=$GRAFT_REPO_ROOT/langscene-x_b200/lsx_b200/liblsx_b200_t64.so timeout 900 python -m pytest tests/test_parity_gpu.py -x -q 2>&1 | tail -3 >> $L
cat $L

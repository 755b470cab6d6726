mkdir -p gpurun_out
T=${TAG:-r02r}
for v in build_variants/lib_chol_staged.so ""; do
  n=$( [ -z "$v" ] && echo new || echo staged )
  PYPERIOD_B200_LIB=$v timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:qo_solve --csv --log-file gpurun_out/${T}_launches_$n.csv python tools/prof_ram_solve.py 2048 > gpurun_out/${T}_$n.log 2>&1
  echo "== $n" >> gpurun_out/${T}_qo.log
  PYPERIOD_B200_LIB=$v timeout 300 python tools/perf_qo.py 8192 >> gpurun_out/${T}_qo.log 2>&1
done

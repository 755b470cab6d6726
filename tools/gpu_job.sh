mkdir -p gpurun_out
T=${TAG:-r04k}
: > gpurun_out/${T}_soak.log
for i in 1 2 3 4; do
  timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider > /tmp/run_$i.log 2>&1
  tail -1 /tmp/run_$i.log >> gpurun_out/${T}_soak.log
  if grep -q "failed" /tmp/run_$i.log; then cp /tmp/run_$i.log gpurun_out/${T}_fail_$i.log; fi
done

mkdir -p gpurun_out
T=${TAG:-r03y}
: > gpurun_out/${T}_soak.log
for i in $(seq 1 16); do
  timeout 600 python -m pytest tests/test_gpu_periods.py -q -x -p no:cacheprovider > /tmp/run_$i.log 2>&1
  tail -1 /tmp/run_$i.log >> gpurun_out/${T}_soak.log
  if grep -q "failed" /tmp/run_$i.log; then cp /tmp/run_$i.log gpurun_out/${T}_fail_$i.log; fi
done

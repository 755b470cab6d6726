mkdir -p gpurun_out
T=${TAG:-r04b}
: > gpurun_out/${T}_soak.log
for i in $(seq 1 40); do
  timeout 300 python -m pytest tests/test_gpu_periods.py -q -x -p no:cacheprovider -k "ragged or scaled_and_offset or large_window or best_frequency_vs or differential or pipelined or device_tensor_io" > /tmp/run_$i.log 2>&1
  tail -1 /tmp/run_$i.log >> gpurun_out/${T}_soak.log
  if grep -q "failed" /tmp/run_$i.log; then cp /tmp/run_$i.log gpurun_out/${T}_fail_$i.log; fi
done

mkdir -p gpurun_out
T=${TAG:-r04f}
: > gpurun_out/${T}_soak_det.log
for i in $(seq 1 24); do
  timeout 300 python tools/soak_determinism.py 2>&1 | grep -v "Warn\|warn(" | grep -v " ok$" >> gpurun_out/${T}_soak_det.log
done

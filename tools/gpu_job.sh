mkdir -p gpurun_out
T=${TAG:-r04d}
timeout 1200 python tools/soak_determinism.py > gpurun_out/${T}_soak_det.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_soak_det.log

mkdir -p gpurun_out
T=${TAG:-r03a}
timeout 600 python -m pytest tests/test_gpu_periods.py -q -k "f32 or fold_modes or differential" > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
timeout 300 python tools/perf_mbest.py 16384 f32,hier,f32 > gpurun_out/${T}_perf.log 2>&1
timeout 300 python tools/check_f32.py > gpurun_out/${T}_check_f32.log 2>&1

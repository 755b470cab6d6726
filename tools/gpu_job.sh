mkdir -p gpurun_out
T=${TAG:-r02l}
timeout 900 python -m pytest tests/test_gpu_periods.py tests/test_gpu_determinism.py -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
timeout 300 python tools/probe_s2l.py > gpurun_out/${T}_probe_s2l.log 2>&1
timeout 300 python tools/perf_mbest.py > gpurun_out/${T}_perf_mbest.log 2>&1

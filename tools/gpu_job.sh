mkdir -p gpurun_out
T=${TAG:-r03n}
timeout 1200 python -m pytest tests/test_gpu_periods.py -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
PP_TRUNC=1 timeout 300 python tools/perf_mbest.py 16384 hier,norider > gpurun_out/${T}_perf_trunc.log 2>&1
timeout 300 python tools/perf_mbest.py 16384 hier > gpurun_out/${T}_perf_plain.log 2>&1

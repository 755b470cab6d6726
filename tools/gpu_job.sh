mkdir -p gpurun_out
T=${TAG:-r03t}
timeout 300 python tools/perf_qo.py 8192 > gpurun_out/${T}_perf.log 2>&1
timeout 300 python tools/perf_qo.py 8192 >> gpurun_out/${T}_perf.log 2>&1
timeout 300 python tools/perf_qo_trunc.py 8192 >> gpurun_out/${T}_perf.log 2>&1

set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r02_pytest1.log
timeout 300 python tools/perf_qo.py 8192 > gpurun_out/r02_perf_qo.log 2>&1
timeout 600 python tools/perf_ram_weights.py 1024 > gpurun_out/r02_perf_ramw.log 2>&1
tail -5 gpurun_out/r02_pytest1.log; cat gpurun_out/r02_perf_qo.log gpurun_out/r02_perf_ramw.log

set -x
mkdir -p gpurun_out
timeout 300 python tools/perf_qo.py 8192 > gpurun_out/r02_perf_qo4.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_qo.py tests/test_gpu_ramanujan.py -m gpu -q 2>&1 | tail -30 > gpurun_out/r02_pytest4.log
timeout 600 python tools/perf_ram_weights.py 4096 > gpurun_out/r02_perf_ramw4.log 2>&1
cat gpurun_out/r02_perf_qo4.log gpurun_out/r02_perf_ramw4.log; tail -12 gpurun_out/r02_pytest4.log

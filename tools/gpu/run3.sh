set -x
mkdir -p gpurun_out
timeout 900 python tools/diag_cfg5.py > gpurun_out/r02_diag_cfg5.log 2>&1
timeout 300 python tools/perf_qo.py 8192 > gpurun_out/r02_perf_qo3.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_qo.py tests/test_gpu_ramanujan.py -m gpu -q 2>&1 | tail -30 > gpurun_out/r02_pytest3.log
cat gpurun_out/r02_diag_cfg5.log gpurun_out/r02_perf_qo3.log; tail -12 gpurun_out/r02_pytest3.log

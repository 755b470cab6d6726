set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r02_pytest2.log
timeout 300 python tools/perf_qo.py 8192 > gpurun_out/r02_perf_qo2.log 2>&1
timeout 600 python tools/perf_ram_weights.py 1024 > gpurun_out/r02_perf_ramw2.log 2>&1
timeout 300 python tools/perf_mbest.py 32768 hier > gpurun_out/r02_perf_mbest2.log 2>&1
tail -25 gpurun_out/r02_pytest2.log; cat gpurun_out/r02_perf_qo2.log gpurun_out/r02_perf_ramw2.log gpurun_out/r02_perf_mbest2.log

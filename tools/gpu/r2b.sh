set -x
mkdir -p gpurun_out
timeout 300 python tools/perf_qo.py 8192 > gpurun_out/r02b_perf_qo.log 2>&1
timeout 300 python tools/probe_qo_timing.py 8192 > gpurun_out/r02b_probe_qo.log 2>&1
cat gpurun_out/r02b_perf_qo.log gpurun_out/r02b_probe_qo.log

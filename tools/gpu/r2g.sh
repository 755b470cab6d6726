set -x
mkdir -p gpurun_out
timeout 600 python tools/probe_qo_big.py 32768 > gpurun_out/r02g_probe_qo_big.log 2>&1
cat gpurun_out/r02g_probe_qo_big.log

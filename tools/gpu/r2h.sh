set -x
mkdir -p gpurun_out
timeout 120 python tools/probe_tf32.py 70 2048 400 > gpurun_out/r02h_tf32.log 2>&1
timeout 120 python tools/probe_tf32.py 300 4096 >> gpurun_out/r02h_tf32.log 2>&1
timeout 200 python tools/probe_tf32.py 2048 4096 >> gpurun_out/r02h_tf32.log 2>&1
cat gpurun_out/r02h_tf32.log

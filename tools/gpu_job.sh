mkdir -p gpurun_out
T=${TAG:-r02j}
timeout 600 python -m pytest tests/test_gpu_ramanujan.py -q -x > gpurun_out/${T}_pytest_ram.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_ram.log
timeout 600 python tools/perf_ram.py 4096 fp64 > gpurun_out/${T}_perf_ram.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ram_fused -c 1 -o gpurun_out/${T}_ram_fused python tools/prof_ram.py 1024 > gpurun_out/${T}_ncu_ram.log 2>&1

mkdir -p gpurun_out
T=${TAG:-r04i}
timeout 300 python -m pytest tests/test_gpu_ramanujan.py -q -x -k "tf32" > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
timeout 300 python tools/perf_ram.py 4096 fp64,tf32 > gpurun_out/${T}_perf.log 2>&1

mkdir -p gpurun_out
T=${TAG:-r02t}
timeout 900 python -m pytest tests/test_gpu_qo.py tests/test_gpu_ramanujan.py tests/test_gpu_determinism.py -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:qo_solve --csv --log-file gpurun_out/${T}_launches.csv python tools/prof_ram_solve.py 2048 > gpurun_out/${T}_prof.log 2>&1
timeout 300 python tools/perf_qo.py 8192 > gpurun_out/${T}_qo.log 2>&1
timeout 300 python tools/perf_ram_weights.py 4096 > gpurun_out/${T}_ramw.log 2>&1

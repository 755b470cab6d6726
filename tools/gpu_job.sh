mkdir -p gpurun_out
T=${TAG:-r03m}
timeout 900 python -m pytest tests/test_gpu_qo.py tests/test_gpu_ramanujan.py -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
timeout 600 python tools/probe_qo_e2e.py > gpurun_out/${T}_e2e.log 2>&1
timeout 900 python bench.py --secondary 5qo,5ram > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "rc=$?" >> gpurun_out/${T}_bench.err

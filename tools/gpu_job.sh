mkdir -p gpurun_out
T=${TAG:-r03u}
timeout 900 python -m pytest tests/test_gpu_qo.py -q > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log

set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_qo.py tests/test_gpu_ramanujan.py -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r02d_pytest.log
cat gpurun_out/r02d_pytest.log

set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r02e_pytest.log
cat gpurun_out/r02e_pytest.log

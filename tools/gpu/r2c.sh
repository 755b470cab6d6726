set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_periods.py tests/test_gpu_qo.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02c_pytest.log
timeout 1500 python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
tail -5 gpurun_out/r02c_pytest.log; tail -30 gpurun_out/r02c_bench.err; cat gpurun_out/r02c_bench.json | head -c 20000

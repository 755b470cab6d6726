set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r02a_pytest.log
timeout 600 python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
timeout 900 python tools/bench_configs.py --no-cpu > gpurun_out/r02a_configs.jsonl 2> gpurun_out/r02a_configs.err
tail -12 gpurun_out/r02a_pytest.log; cat gpurun_out/r02a_bench.json; cat gpurun_out/r02a_configs.jsonl; tail -5 gpurun_out/r02a_configs.err

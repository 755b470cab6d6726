mkdir -p gpurun_out
T=${TAG:-r03o}
timeout 900 python -m pytest tests/test_gpu_ramanujan.py tests/test_gpu_qo.py -q > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
timeout 600 python tools/prof_ram_full.py 65536 > gpurun_out/${T}_full.log 2>&1
for args in "14 592" "18 592" "6 592"; do timeout 300 python tools/probe_solve_big.py $args 2>&1 | tail -1 >> gpurun_out/${T}_big.log; done

mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ramanujan.py -x -q -k "tf32" > gpurun_out/r02h_tf32.log 2>&1; echo "tf32 rc=$?" >> gpurun_out/r02h_tf32.log
timeout 300 python -m pytest tests/test_gpu_ramanujan.py -x -q -k "f32_compat" > gpurun_out/r02h_compat.log 2>&1; echo "compat rc=$?" >> gpurun_out/r02h_compat.log
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_ramanujan.py::test_tf32_option_tracks_fp64 > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02h_pytest.log
timeout 900 python bench.py > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; echo "bench rc=$?" >> gpurun_out/r02h_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02h_launches.csv python bench.py --steps 1 --warmup 3 --windows 8192 --e2e-steps 1 --no-cpu-baseline --secondary none > gpurun_out/r02h_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mbest_kernel -s 1 -c 1 -o gpurun_out/r02h_mbest python tools/prof_one.py 2072 > gpurun_out/r02h_ncu_full.log 2>&1
ls -la gpurun_out

mkdir -p gpurun_out
T=${TAG:-r02m}
for v in "" build_variants/lib_ch16.so build_variants/lib_ch24.so build_variants/lib_ch64.so ""; do
  echo "== variant '$v'" >> gpurun_out/${T}_s2l_variants.log
  PYPERIOD_B200_LIB=$v timeout 300 python tools/probe_s2l.py 2>&1 | tail -2 >> gpurun_out/${T}_s2l_variants.log
done
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "rc=$?" >> gpurun_out/${T}_bench.err

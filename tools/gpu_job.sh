mkdir -p gpurun_out
T=${TAG:-r02v}
for args in "14 148" "14 592" "18 592" "6 592"; do
  timeout 300 python tools/probe_solve_big.py $args 2>&1 | tail -2 >> gpurun_out/${T}_big.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qo_solve_kernel -s 1 -c 1 -o gpurun_out/${T}_deep python tools/probe_solve_big.py 14 148 > gpurun_out/${T}_ncu.log 2>&1

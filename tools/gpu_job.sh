mkdir -p gpurun_out
T=${TAG:-r03r}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ram_umma_tf32 -s 1 -c 1 -o gpurun_out/${T}_tf32 python tools/prof_ram_tf32.py 1024 > gpurun_out/${T}_ncu_tf32.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qo_find_kernel -s 1 -c 1 -o gpurun_out/${T}_qo_find python tools/prof_qo.py 2368 > gpurun_out/${T}_ncu_qo.log 2>&1

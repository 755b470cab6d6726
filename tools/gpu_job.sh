mkdir -p gpurun_out
T=${TAG:-r02y}
timeout 600 python tools/prof_ram_full.py 65536 > gpurun_out/${T}_full.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv python tools/prof_ram_full.py 65536 > gpurun_out/${T}_full_ncu.log 2>&1

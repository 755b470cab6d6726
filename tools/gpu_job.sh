mkdir -p gpurun_out
T=${TAG:-r03h}
timeout 900 python bench.py --secondary 5qo > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "rc=$?" >> gpurun_out/${T}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:qo_find --csv --log-file gpurun_out/${T}_launches.csv python tools/probe_qo_full.py 65536 > gpurun_out/${T}_qo_ncu.log 2>&1

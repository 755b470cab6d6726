#!/bin/bash
# run the perf probe against each prebuilt library variant in tools/libs
for f in tools/libs/lib_*.so; do
  cp $f pyperiod_b200/libpyperiod_b200.so
  echo "=== $f"
  python tools/perf_mbest.py ${1:-32768} ${2:-hier} 2>&1 | grep win/s
done

#!/bin/bash
# usage: tools/perf_variants.sh B variant.so [variant.so ...]   (default lib first)
B=$1; shift
echo "== default"; python tools/perf_mbest.py $B hier 2>&1 | grep -v Warn
for v in "$@"; do echo "== $v"; PYPERIOD_B200_LIB=$PWD/$v python tools/perf_mbest.py $B hier 2>&1 | grep -v Warn; done

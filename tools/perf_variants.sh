#!/bin/bash
# usage: tools/perf_variants.sh B modes variant.so [variant.so ...]   (default lib first, then again last: noise check)
B=$1; shift; M=$1; shift
echo "== default"; python tools/perf_mbest.py $B $M 2>&1 | grep -v Warn
for v in "$@"; do echo "== $v"; PYPERIOD_B200_LIB=$PWD/$v python tools/perf_mbest.py $B $M 2>&1 | grep -v Warn; done
echo "== default (again)"; python tools/perf_mbest.py $B $M 2>&1 | grep -v Warn

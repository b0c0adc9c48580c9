#!/bin/bash
# ncu --set full capture of the launches whose DEMANGLED name matches $1 (a regex), into gpurun_out/$2.ncu-rep
#   scripts/ncu_kernel.sh '<regex>' <out-name> <count> <command...>
re="$1"; out="$2"; cnt="$3"; shift 3
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$re" -c "$cnt" \
    -f -o "gpurun_out/$out" "$@" 2>&1 | tail -4
ls -la gpurun_out/$out.ncu-rep

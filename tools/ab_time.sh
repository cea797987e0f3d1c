#!/bin/bash
# A/B timing of two builds of libfdx_b200.so (sequential processes, interleaved rounds):
#   tools/ab_time.sh <libA.so> <libB.so> [rounds] -- op acc dtype n [op acc dtype n ...]
A=$1; B=$2; R=${3:-3}; shift 3; [ "$1" == "--" ] && shift
for r in $(seq 1 $R); do
  for lib in $A $B; do
    args=("$@")
    for ((i=0; i<${#args[@]}; i+=4)); do
      echo -n "$(basename $lib) ${args[i]} a=${args[i+1]} ${args[i+2]} n=${args[i+3]}: "
      FDX_B200_LIB=$lib ROUNDS=3 python tools/time_variants.py ${args[i]} ${args[i+1]} ${args[i+2]} ${args[i+3]} 2>&1 | tail -1 | cut -c1-120
    done
  done
done

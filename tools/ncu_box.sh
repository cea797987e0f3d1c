#!/bin/bash
# usage (on the GPU box): tools/ncu_box.sh <tag> <kernel regex> <skip> -- <command...>
# One `ncu --set full` capture of one launch; the report stays on the box (tens of MB), its summary and the
# per-instruction source page come back as text under gpurun_out/.
tag=$1; kre=$2; skip=$3; shift 4
ncu --set full --clock-control none --import-source on -k regex:$kre -s $skip -c 1 -o /tmp/ncu_$tag -f "$@" > gpurun_out/ncu_$tag.log 2>&1
python tools/ncu_summary.py /tmp/ncu_$tag.ncu-rep "$tag: $*" > gpurun_out/ncu_$tag.txt 2>&1
ncu -i /tmp/ncu_$tag.ncu-rep --page source --csv > gpurun_out/ncu_${tag}_source.csv 2>/dev/null
ncu -i /tmp/ncu_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_${tag}_raw.csv 2>/dev/null
rm -f /tmp/ncu_$tag.ncu-rep

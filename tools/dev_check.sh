#!/bin/bash
# usage (on the GPU box): tools/dev_check.sh <op> <n> <dtype> <acc> <tag> [cfgs] [chunks]
op=${1:-laplacian}; n=${2:-512}; dt=${3:-f32}; acc=${4:-4}; tag=${5:-dev}; cfgs=${6:-0}; chunks=${7:-0}
SWEEP_CFGS=$cfgs SWEEP_CHUNKS=$chunks python tools/sweep_fused.py $op $n $dt $acc > gpurun_out/sweep_$tag.log 2>&1
grep -v BEST gpurun_out/sweep_$tag.log | python -c "
import sys,json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: print(l.strip()); continue
    print(r['op'],r['n'],'cfg',r['cfg'],'chunk',r['chunk'],'ms',r['ms'],'gpts',r['gpts'],'gbs',r['gbs'],'err %.2e'%r['err'])
"
python tools/prof_one.py $op $n $dt $acc 3 > gpurun_out/plain_$tag.log 2>&1 && ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread -k regex:fused3d -s 1 -c 1 --csv --log-file gpurun_out/ncu_$tag.csv python tools/prof_one.py $op $n $dt $acc 3 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/ncu_$tag.csv')) if len(r)>10]
for r in rows[1:]:
    print(r[-3], r[-1])
PY

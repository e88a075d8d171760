#!/bin/bash
# usage: scripts/ncu_longest.sh <precision> <kernel-regex> <out-name>
# pass 1: duration of every matching launch of scripts/profile_target.py; pass 2: --set full on the longest one.
P=$1; RX=$2; OUT=$3
mkdir -p gpurun_out
CP_PRECISION=$P timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:$RX --csv --log-file gpurun_out/${OUT}_list.csv python scripts/profile_target.py > /dev/null 2>&1
IDX=$(python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/${OUT}_list.csv', errors='ignore')) if len(r)>10]
hdr=rows[0]; vi=hdr.index('Metric Value')
vals=[float(r[vi].replace(',','')) for r in rows[1:]]
half=len(vals)//2                      # two forwards: take the longest launch of the second one
best=max(range(half,len(vals)), key=lambda i: vals[i])
print(best)
PY
)
echo "$P $RX longest launch index $IDX" >> gpurun_out/${OUT}_idx.txt
CP_PRECISION=$P timeout 500 ncu --set full --clock-control none --import-source on -k regex:$RX -s $IDX -c 1 -f -o gpurun_out/$OUT python scripts/profile_target.py > gpurun_out/${OUT}_ncu.log 2>&1

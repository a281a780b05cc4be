#!/bin/bash
# GPU box: per-kernel device times of the Monte-Carlo step (ncu launch list; shares, not absolutes).
# usage: bash tools/mc_launches.sh <tag> <chains>
TAG=${1:-r2}; M=${2:-131072}
mkdir -p gpurun_out
CMD="python tools/mc_throughput.py $M 3 0"
$CMD > gpurun_out/mc_plain_${TAG}_$M.log 2>&1 || exit 2
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 24 --csv --log-file gpurun_out/mc_launches_${TAG}_$M.csv $CMD > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/mc_launches_${TAG}_$M.csv")) if len(r)>5 and r[0].isdigit()]
for r in rows: print(r[4][:60].ljust(62), r[-1])
PY

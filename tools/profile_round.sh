#!/bin/bash
# GPU box: the round's evidence.  usage: bash tools/profile_round.sh <tag> <stage>
#   stage bench    : plain bench (the numbers) + reference arm
#   stage launches : ncu launch list of the short command (after it exited 0 without ncu)
#   stage phase1   : ncu --set full of the three root-search launches
#   stage other    : ncu --set full of prep + phase 2
#   stage mcbench  : bench lines of the Monte-Carlo / grid workloads (configs 1, 3, 4, 5)
# Nothing printed under ncu is a bench value.
TAG=${1:-r2}; STAGE=${2:-bench}
mkdir -p gpurun_out
CMD="python bench.py --models 524288 --steps 1 --warmup 3 --no-cpu --no-e2e"
case $STAGE in
bench)
  python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || exit 1
  python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
  cut -c 1-600 gpurun_out/bench_$TAG.json; cut -c 1-300 gpurun_out/bench_ref_$TAG.json ;;
launches)
  $CMD > gpurun_out/plain_$TAG.log 2>&1 || exit 2
  ncu --metrics gpu__time_duration.sum --clock-control none -s 16 -c 8 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > /dev/null 2>&1
  tail -9 gpurun_out/launches_$TAG.csv | cut -c 1-200 ;;
phase1)
  $CMD > gpurun_out/plain_$TAG.log 2>&1 || exit 2
  ncu --set full --clock-control none --import-source on -k regex:phase1 -s 9 -c 3 -o gpurun_out/prof_phase1_$TAG -f $CMD > gpurun_out/ncu_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_$TAG.log ;;
other)
  $CMD > gpurun_out/plain_$TAG.log 2>&1 || exit 2
  ncu --set full --clock-control none --import-source on -k regex:"prep|phase2" -s 6 -c 2 -o gpurun_out/prof_other_$TAG -f $CMD > gpurun_out/ncu_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_$TAG.log ;;
mcbench)
  rm -f gpurun_out/mc_bench_$TAG.jsonl
  for w in "--workload mc --chains 256" "--workload mc --chains 256 --love" "--workload mc --chains 131072 --mc-steps 10" "--workload grid --points 2000 --chains 16 --mc-steps 20" "--workload mc --thermal --chains 256" "--workload mc --thermal --chains 32768 --mc-steps 10"; do
    python bench.py $w --cpu-seconds 4 >> gpurun_out/mc_bench_$TAG.jsonl 2>> gpurun_out/mc_bench_$TAG.err
  done
  python - <<PY
import json
for l in open("gpurun_out/mc_bench_$TAG.jsonl"):
    d = json.loads(l); print(d["config"]["workload"][:48], "%.3g evals/s" % d["value"], "%.3f ms/MC step" % d["config"]["ms_per_mc_step"], "e2e %.3g" % d["e2e"]["value"], "cpu %.3g" % d["cpu_baseline"]["value"])
PY
  ;;
esac

#!/bin/bash
# GPU box: ncu --set full of every kernel of one Monte-Carlo step (131072 chains), after the plain run exited 0.
TAG=${1:-r2}
mkdir -p gpurun_out
CMD="python tools/mc_throughput.py 131072 2 0"
$CMD > gpurun_out/mc_plain_$TAG.log 2>&1 || exit 2
ncu --set full --clock-control none --import-source on -k regex:"mc_propose_build|mc_finish|prep_kernel|phase1" -s 24 -c 6 -o gpurun_out/prof_mc_$TAG -f $CMD > gpurun_out/ncu_mc_$TAG.log 2>&1
tail -2 gpurun_out/ncu_mc_$TAG.log

#!/bin/bash
# Interleaved A/B of two builds on one B200 (box clocks drift by a few percent under the power cap: alternate, repeat).
#   gpurun --timeout 1200 -- 'bash tools/gpu_exp_ab.sh <tag> <libA.so> <libB.so> [cases] [rounds]'
cd ${GRAFT_REPO_ROOT:-.}
tag=$1; A=$2; B=$3
CASES=${4:-ntom:mixed,2perstage:mixed,linear:mixed,v0:mixed,oneonen:mixed,multiproduct:mixed}
R=${5:-3}
out=gpurun_out/ab_$tag.log
mkdir -p gpurun_out
{
for i in $(seq 1 $R); do
  echo "== round $i: A = $A"
  SCG_LIB=$A timeout 600 python tools/quick_rollout.py --cases $CASES
  echo "== round $i: B = $B"
  SCG_LIB=$B timeout 600 python tools/quick_rollout.py --cases $CASES
done
} > $out 2>&1
cat $out

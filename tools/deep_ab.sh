#!/bin/bash
# C4 (4096 spheres) and the 20 000-sphere scene with two and three levels of sphere bounds.
# usage (under gpurun): bash tools/deep_ab.sh <tag>
TAG=${1:-deep}
OUT=gpurun_out/${TAG}_deep.txt
: > $OUT
for setting in "RL_DEEP_CLUSTERS=0" "RL_DEEP_CLUSTERS=1" "RL_DEEP_CLUSTERS=1 RL_CLUSTER_LEAF=16" "RL_DEEP_CLUSTERS=1 RL_CLUSTER_LEAF=12"; do
  echo "== $setting" | tee -a $OUT
  env $setting RL_RATES_ONLY=C4 timeout 300 python tools/config_rates.py 2>&1 | tail -1 | tee -a $OUT
done

#!/bin/bash
# Experiment: the warp-ring form of the trace kernel with a block barrier every N iterations, against
# the block-ring form (one barrier per iteration), on the built-in scene.
# usage (under gpurun): bash tools/sync_every_probe.sh <tag> variants/S_sync_every.so
TAG=${1:-sync}; V=$2
LIB=robigo-luculenta_b200/librl_b200.so
OUT=gpurun_out/${TAG}_sync_every.txt
mkdir -p gpurun_out; : > $OUT
cp $LIB /tmp/librl_b200.keep
echo "== library in place (block ring)" | tee -a $OUT
RL_RATES_ONLY=C2 timeout 200 python tools/config_rates.py 2>&1 | tail -1 | tee -a $OUT
cp $V $LIB; touch $LIB
echo "== $V, block ring" | tee -a $OUT
RL_RATES_ONLY=C2 timeout 200 python tools/config_rates.py 2>&1 | tail -1 | tee -a $OUT
for N in 0 1 2 3 4 8; do
  echo "== $V, warp rings, barrier every $N" | tee -a $OUT
  RL_TRACE_LOCKSTEP=0 RL_TRACE_SYNC_EVERY=$N RL_RATES_ONLY=C2 timeout 200 python tools/config_rates.py 2>&1 | tail -1 | tee -a $OUT
done
timeout 120 python -c "import os; os.environ['RL_TRACE_LOCKSTEP']='0'; os.environ['RL_TRACE_SYNC_EVERY']='2'; import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee -a $OUT
cp /tmp/librl_b200.keep $LIB; touch $LIB

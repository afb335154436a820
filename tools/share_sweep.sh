#!/bin/bash
# Strict-mode replay, one launch per batch, over the share of the GPU's block slots a small launch
# may take (RL_TRACE_SHARE_MAX), each setting several times (the replay's end effects make single
# runs noisy); gpurun_out/<tag>_share.txt.  usage: bash tools/share_sweep.sh <tag> [batches] [threads]
TAG=${1:-s}; B=${2:-4096}; T=${3:-16}
R=robigo-luculenta_b200/rl_replay
OUT=gpurun_out/${TAG}_share.txt
mkdir -p gpurun_out; : > $OUT
one() { # share_max groups threads batches
  line=$(RL_TRACE_SHARE_MAX=$1 RL_TRACE_GROUPS=$2 timeout 120 $R --width 1024 --height 1024 --threads $3 --batches $4 --batch 524288 --seed 24301 --scene 2 --out /tmp/ssweep --mode strict 2>>gpurun_out/${TAG}_share.err | tail -1)
  echo "share_max=$1 groups=$2 threads=$3 batches=$4: $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read() or "{}"); print(d.get("mrays_per_s"), d.get("seconds"), d.get("dispatch"), d.get("worker_seconds",{}).get("sleep"))' 2>/dev/null)" | tee -a $OUT
}
for rep in 1 2 3; do for s in 3 6 12 24; do one $s 0 $T $B; done; done
one 12 0 4 $B
one 12 0 4 $B
one 12 0 $T 10240

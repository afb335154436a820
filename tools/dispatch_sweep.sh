#!/bin/bash
# Strict-mode replay over the dispatcher's two knobs: launches in flight (RL_TRACE_SLOTS) and
# batches per launch (RL_TRACE_GROUP_MAX); gpurun_out/<tag>_dispatch.txt.  usage: bash tools/dispatch_sweep.sh <tag> [batches] [threads]
TAG=${1:-d}; B=${2:-2048}; T=${3:-16}
R=robigo-luculenta_b200/rl_replay
OUT=gpurun_out/${TAG}_dispatch.txt
mkdir -p gpurun_out; : > $OUT
one() { # slots group
  line=$(RL_TRACE_SLOTS=$1 RL_TRACE_GROUP_MAX=$2 RL_TRACE_GROUPS=${3:-1} timeout 120 $R --width 1024 --height 1024 --threads $T --batches $B --batch 524288 --seed 24301 --scene 2 --out /tmp/dsweep --mode strict 2>>gpurun_out/${TAG}_dispatch.err | tail -1)
  echo "slots=$1 group_max=$2 groups=${3:-1} threads=$T conn=${CUDA_DEVICE_MAX_CONNECTIONS:-default}: $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read() or "{}"); print(d.get("mrays_per_s"), d.get("seconds"), d.get("dispatch"), d.get("worker_seconds",{}).get("sleep"))' 2>/dev/null)" | tee -a $OUT
}
one 2 32 0
one 2 32 0
CUDA_DEVICE_MAX_CONNECTIONS=32 one 2 32 0
CUDA_DEVICE_MAX_CONNECTIONS=32 one 2 32 0
for s in 2 3 4; do for g in 4 16 32; do one $s $g; done; done
CUDA_DEVICE_MAX_CONNECTIONS=32 one 3 16
CUDA_DEVICE_MAX_CONNECTIONS=32 one 2 32
T=4
one 2 32 0
one 3 16
one 2 32

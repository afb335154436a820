#!/bin/bash
# The strict-mode replay several times over (run-to-run spread), then under a few knob settings.
# usage (under gpurun): bash tools/replay_repeat.sh <tag> [batches]
TAG=${1:-rep}; B=${2:-6144}
R=robigo-luculenta_b200/rl_replay
OUT=gpurun_out/${TAG}_replay_repeat.txt
mkdir -p gpurun_out; : > $OUT
one() { # label env...
  label=$1; shift
  line=$(env "$@" timeout 120 $R --width 1024 --height 1024 --threads ${RL_REPLAY_THREADS:-16} --batches $B --batch 524288 --seed 24301 --scene 2 --out /tmp/rep --mode strict 2>>gpurun_out/${TAG}_replay_repeat.err | tail -1)
  echo "$label: $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read() or "{}"); print(d.get("mrays_per_s"), d.get("seconds"))' 2>/dev/null)" | tee -a $OUT
}
for i in 1 2 3 4; do one "default #$i" RL_NOOP=1; done
for i in 1 2; do one "share 24 #$i" RL_TRACE_SHARE_MAX=24; done
for i in 1 2; do one "connections 32 #$i" CUDA_DEVICE_MAX_CONNECTIONS=32; done
for i in 1 2; do one "threads 4 #$i" RL_REPLAY_THREADS=4; done
for i in 1 2; do one "threads 8 #$i" RL_REPLAY_THREADS=8; done

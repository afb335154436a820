#!/bin/bash
# The strict-mode replay N times, full JSON lines kept (run-to-run spread and where the time goes).
# usage (under gpurun): bash tools/replay_spread.sh <tag> [runs] [batches]
TAG=${1:-spread}; N=${2:-8}; B=${3:-6144}
R=robigo-luculenta_b200/rl_replay
OUT=gpurun_out/${TAG}_replay_spread.jsonl
mkdir -p gpurun_out; : > $OUT
for i in $(seq 1 $N); do
  timeout 120 $R --width 1024 --height 1024 --threads 16 --batches $B --batch 524288 --seed 24301 --scene 2 --out /tmp/spread --mode strict 2>>gpurun_out/${TAG}_replay_spread.err | tail -1 | tee -a $OUT | python -c 'import sys,json; d=json.loads(sys.stdin.read() or "{}"); print(d.get("mrays_per_s"), d.get("seconds"), d.get("worker_seconds"))'
done

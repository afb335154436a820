#!/bin/bash
# Strict-mode replay (host/rl_replay.cpp) in its variants, one JSON line each (gpurun_out/e2e_sweep.jsonl).
# usage: tools/e2e_sweep.sh [batches] [threads]
B=${1:-2048}; T=${2:-16}
R=robigo-luculenta_b200/rl_replay
OUT=gpurun_out/e2e_sweep.jsonl
mkdir -p gpurun_out; : > $OUT
run() { # label, env..., -- args
  label=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  line=$(env "${envs[@]}" timeout 300 $R --width 1024 --height 1024 --threads $T --batches $B --batch 524288 --seed 24301 --scene 2 --out /tmp/e2e_sweep "$@" 2>>gpurun_out/e2e_sweep.err | tail -1)
  echo "{\"variant\": \"$label\", \"result\": ${line:-null}}" >> $OUT
  echo "$label: $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read() or "{}"); print(d.get("mrays_per_s"), d.get("seconds"), d.get("dispatch"), d.get("worker_seconds"))' 2>/dev/null)"
}
run "strict, group launches (default)" X=1 -- --mode strict
run "strict, group launches, sleep 1 ms" X=1 -- --mode strict --sleep-ms 1
run "strict, groups of at most 8" RL_TRACE_GROUP_MAX=8 -- --mode strict
run "strict, one launch per batch" RL_TRACE_GROUPS=0 -- --mode strict
run "strict, one launch per batch, host re-upload (round 1)" RL_TRACE_GROUPS=0 -- --mode strict --consume host --sleep-ms 1
run "strict, deferred records" X=1 -- --mode strict --records deferred
run "device mode" X=1 -- --mode device
run "strict, 8 threads" X=1 -- --mode strict --threads 8
run "strict, 4 threads" X=1 -- --mode strict --threads 4
run "strict, synchronous render" X=1 -- --mode strict --async-render 0

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
  echo "$label: $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read() or "{}"); print(d.get("mrays_per_s"), d.get("seconds"), d.get("worker_seconds"))' 2>/dev/null)"
}
run "strict service R=2 (default)" -- --mode strict
run "strict service R=0" RL_SERVICE_RESERVED_SMS=0 -- --mode strict
run "strict service R=1" RL_SERVICE_RESERVED_SMS=1 -- --mode strict
run "strict service R=4" RL_SERVICE_RESERVED_SMS=4 -- --mode strict
run "strict launches (no service)" RL_TRACE_SERVICE=0 -- --mode strict
run "strict round-1 (no service, host re-upload)" RL_TRACE_SERVICE=0 -- --mode strict --consume host
run "strict service deferred records" -- --mode strict --records deferred
run "device mode service" -- --mode device
run "strict service wait-kernel" RL_SERVICE_WAIT_KERNEL=1 -- --mode strict
run "strict service 8 threads" -- --mode strict --threads 8
run "strict service 4 threads" -- --mode strict --threads 4

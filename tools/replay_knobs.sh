#!/bin/bash
# Strict-mode replay under the launch-policy knobs of launch_trace; gpurun_out/<tag>_knobs.txt.
# usage: bash tools/replay_knobs.sh <tag> [batches]
TAG=${1:-k}; B=${2:-6144}
R=robigo-luculenta_b200/rl_replay
OUT=gpurun_out/${TAG}_knobs.txt
mkdir -p gpurun_out; : > $OUT
one() { # label env...
  label=$1; shift
  line=$(env "$@" timeout 120 $R --width 1024 --height 1024 --threads ${RL_REPLAY_THREADS:-16} --batches $B --batch 524288 --seed 24301 --scene 2 --out /tmp/knobs --mode strict 2>>gpurun_out/${TAG}_knobs.err | tail -1)
  echo "$label: $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read() or "{}"); print(d.get("mrays_per_s"), d.get("seconds"), d.get("worker_seconds",{}).get("trace"), d.get("worker_seconds",{}).get("plot"), d.get("worker_seconds",{}).get("gather"))' 2>/dev/null)" | tee -a $OUT
}
one "cta 384 share 12" RL_TRACE_SMALL_CTA=384 RL_TRACE_SHARE_MAX=12
one "cta 384 share 16" RL_TRACE_SMALL_CTA=384 RL_TRACE_SHARE_MAX=16
one "cta 384 share 24" RL_TRACE_SMALL_CTA=384 RL_TRACE_SHARE_MAX=24
one "cta 384 share 32" RL_TRACE_SMALL_CTA=384 RL_TRACE_SHARE_MAX=32
one "cta 384 share 8" RL_TRACE_SMALL_CTA=384 RL_TRACE_SHARE_MAX=8
one "cta 384 share 24 again" RL_TRACE_SMALL_CTA=384 RL_TRACE_SHARE_MAX=24
one "cta 512 share 12" RL_TRACE_SMALL_CTA=512 RL_TRACE_SHARE_MAX=12
one "cta 384 share 24 4 threads" RL_TRACE_SMALL_CTA=384 RL_TRACE_SHARE_MAX=24 RL_REPLAY_THREADS=4

#!/bin/bash
# Replay and small-launch rate under the ring-size and splat carve-out knobs.  usage: bash tools/replay_knobs2.sh <tag>
TAG=${1:-k2}; B=${2:-6144}
R=robigo-luculenta_b200/rl_replay
OUT=gpurun_out/${TAG}_knobs2.txt
mkdir -p gpurun_out; : > $OUT
one() { label=$1; shift
  line=$(env "$@" timeout 120 $R --width 1024 --height 1024 --threads 16 --batches $B --batch 524288 --seed 24301 --scene 2 --out /tmp/k2 --mode strict 2>>gpurun_out/${TAG}_knobs2.err | tail -1)
  echo "replay $label: $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read() or "{}"); print(d.get("mrays_per_s"), d.get("seconds"))' 2>/dev/null)" | tee -a $OUT
}
for rep in 1 2; do
  one "default" RL_NOOP=1
  one "splat carve-out default (-1)" RL_SPLAT_CARVEOUT=-1
  one "ring half" RL_TRACE_RING_SHRINK=1
  one "ring quarter" RL_TRACE_RING_SHRINK=2
done
for S in 0 1 2; do
  echo "== small-launch probe, RL_TRACE_RING_SHRINK=$S" | tee -a $OUT
  RL_TRACE_RING_SHRINK=$S RL_PROBE_CTAS=384 RL_PROBE_SHARES=12,24 timeout 200 python tools/small_launch_probe.py 2>&1 | grep "^{" | tee -a $OUT
  echo "== big launch with 384-thread CTAs, RL_TRACE_RING_SHRINK=$S" | tee -a $OUT
  RL_TRACE_RING_SHRINK=$S RL_TRACE_THREADS_MAX=384 RL_RATES_ONLY=C2 timeout 200 python tools/config_rates.py 2>&1 | tail -1 | tee -a $OUT
done

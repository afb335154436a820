#!/bin/bash
# Small launches and the scheduler replay under CUDA_DEVICE_MAX_CONNECTIONS (hardware work queues:
# streams beyond their number share a queue, and a launch waiting for its own stream's previous
# launch then holds back the launches of other streams behind it).
# usage (under gpurun): bash tools/connections_probe.sh <tag>
TAG=${1:-conn}
OUT=gpurun_out/${TAG}_connections.txt
R=robigo-luculenta_b200/rl_replay
mkdir -p gpurun_out; : > $OUT
for C in 8 32; do
  echo "== CUDA_DEVICE_MAX_CONNECTIONS=$C" | tee -a $OUT
  CUDA_DEVICE_MAX_CONNECTIONS=$C RL_PROBE_CTAS=384 RL_PROBE_SHARES=12,24,48 timeout 300 python tools/small_launch_probe.py 2>&1 | grep "^{" | tee -a $OUT
  for S in 12 24 48; do
    line=$(CUDA_DEVICE_MAX_CONNECTIONS=$C RL_TRACE_SHARE_MAX=$S timeout 120 $R --width 1024 --height 1024 --threads 16 --batches 6144 --batch 524288 --seed 24301 --scene 2 --out /tmp/conn --mode strict 2>>gpurun_out/${TAG}_connections.err | tail -1)
    echo "replay share_max=$S: $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read() or "{}"); print(d.get("mrays_per_s"), d.get("seconds"))' 2>/dev/null)" | tee -a $OUT
  done
done

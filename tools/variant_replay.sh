#!/bin/bash
# A kernel variant against the library in place, on the built-in scene only: one fused launch, the
# small-launch probe, the strict-mode replay.  usage (under gpurun): bash tools/variant_replay.sh <tag> variants/X.so
TAG=${1:-vr}; V=$2
LIB=robigo-luculenta_b200/librl_b200.so
OUT=gpurun_out/${TAG}_variant_replay.txt
R=robigo-luculenta_b200/rl_replay
mkdir -p gpurun_out; : > $OUT
cp $LIB /tmp/librl_b200.keep
run_all() {
  RL_RATES_ONLY=C2 timeout 200 python tools/config_rates.py 2>&1 | tail -1 | tee -a $OUT
  RL_PROBE_CTAS=384 RL_PROBE_SHARES=12 timeout 200 python tools/small_launch_probe.py 2>&1 | grep "^{" | tee -a $OUT
  for i in 1 2 3; do
    line=$(timeout 120 $R --width 1024 --height 1024 --threads 16 --batches 6144 --batch 524288 --seed 24301 --scene 2 --out /tmp/vr --mode strict 2>>gpurun_out/${TAG}_variant_replay.err | tail -1)
    echo "replay #$i: $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read() or "{}"); print(d.get("mrays_per_s"), d.get("seconds"))' 2>/dev/null)" | tee -a $OUT
  done
}
echo "== library in place" | tee -a $OUT; run_all
echo "== $V" | tee -a $OUT; cp $V $LIB; touch $LIB; run_all
cp /tmp/librl_b200.keep $LIB; touch $LIB

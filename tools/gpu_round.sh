#!/bin/bash
# One GPU visit: GPU tests, replay sweep, one ncu capture of the trace kernel, the bench line.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [tests|sweep|dsweep|ssweep|ab|ncu|ncu4|traffic|launches|leaf|bench ...]   (default: tests sweep ncu bench)
TAG=${1:-round}; shift
WHAT=${@:-tests sweep ncu bench}
mkdir -p gpurun_out
for w in $WHAT; do
  case $w in
    tests) timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/${TAG}_gputest.log 2>&1; tail -4 gpurun_out/${TAG}_gputest.log ;;
    sweep) bash tools/e2e_sweep.sh 2048 16 > gpurun_out/${TAG}_sweep.txt 2>&1; mv gpurun_out/e2e_sweep.jsonl gpurun_out/${TAG}_sweep.jsonl; cat gpurun_out/${TAG}_sweep.txt ;;
    ncu) timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_trace python tools/profile_trace.py > gpurun_out/${TAG}_ncu.log 2>&1; tail -2 gpurun_out/${TAG}_ncu.log ;;
    ncu4) RL_PROFILE_SCENE=4 RL_PROFILE_CANVAS=2048 RL_PROFILE_TRACE_ONLY=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_trace_c4 python tools/profile_trace.py > gpurun_out/${TAG}_ncu4.log 2>&1; tail -2 gpurun_out/${TAG}_ncu4.log ;;
    dsweep) bash tools/dispatch_sweep.sh ${TAG} ;;
    ssweep) bash tools/share_sweep.sh ${TAG} ;;
    ab) bash tools/kernel_ab.sh ${TAG} ;;
    deep) bash tools/deep_ab.sh ${TAG} ;;
    traffic) timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_traffic.csv python tools/profile_trace.py > gpurun_out/${TAG}_traffic.log 2>&1; tail -1 gpurun_out/${TAG}_traffic.log ;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_launches.log 2>&1; tail -1 gpurun_out/${TAG}_launches.log | cut -c1-200 ;;
    leaf) for L in 12 16 20 24 29 40; do echo "RL_CLUSTER_LEAF=$L"; RL_CLUSTER_LEAF=$L RL_RATES_ONLY=C4 timeout 200 python tools/config_rates.py 2>&1 | tail -1; done | tee gpurun_out/${TAG}_leaf.txt ;;
    bench) timeout 1500 python bench.py ${BENCH_ARGS:-} > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err; python - <<PY
import json
try:
    l = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("value", l["value"], "ms/step", l["ms_per_step"], "launches", l["gpu_launches"])
    for k in ("e2e", "e2e_deferred_records", "e2e_device"):
        e = l.get(k) or {}
        print(k, e.get("value"), e.get("seconds"), e.get("note"))
    for o in l.get("other_configs") or []:
        print(o["config"], round(o["mrays_per_s"], 1))
    for a in l["roofline"].get("also", []):
        print(a["kernel"][:30], round(a["frac"], 3))
except Exception as e:
    print("bench line unreadable:", e)
PY
    ;;
  esac
done

#!/bin/bash
# One visit to an N-GPU box: the device-to-host ceiling of the host with N GPUs copying (three ways
# of page-locking the host buffer), then the bench line at N GPUs.
# usage (under gpurun --gpus N): bash tools/run_ngpu.sh <tag> <N> [steps] [probe|bench ...]
TAG=${1:-ngpu}; N=${2:-8}; STEPS=${3:-10}; shift; shift; shift
WHAT=${@:-probe bench}
mkdir -p gpurun_out
for w in $WHAT; do
  case $w in
    probe)
      : > gpurun_out/${TAG}_pcie.jsonl
      for a in pinned registered wc; do
        timeout 300 python tools/pcie_probe.py --only $N --directions d2h --alloc $a 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_pcie.jsonl
      done ;;
    bench)
      timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
          bench.py --gpus $N --steps $STEPS --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
      tail -3 gpurun_out/${TAG}_bench.err
      python - <<PY
import json
try:
    l = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("value", l["value"], "n_gpus", l["n_gpus"], "parity", l.get("parity_multi_gpu"))
    for k in ("e2e", "e2e_deferred_records", "e2e_device"):
        e = l.get(k) or {}
        print(k, e.get("value"), e.get("seconds"), e.get("d2h_gb_per_s"), e.get("d2h_bytes_per_step"), e.get("worker_threads_per_gpu"), e.get("note"))
    print("c5", l.get("c5"))
    for o in l.get("other_configs") or []:
        print(o["config"], round(o["mrays_per_s"], 1))
except Exception as e:
    print("bench line unreadable:", e)
PY
      ;;
  esac
done

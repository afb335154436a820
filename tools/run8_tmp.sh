timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2f_bench8.json 2> gpurun_out/r2f_bench8.err
tail -5 gpurun_out/r2f_bench8.err
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/r2f_bench8.json").read().strip().splitlines()[-1])
    print("value", l["value"], "n_gpus", l["n_gpus"], "parity", l.get("parity_multi_gpu"))
    for k in ("e2e", "e2e_deferred_records", "e2e_device"):
        e = l.get(k) or {}
        print(k, e.get("value"), e.get("seconds"), e.get("d2h_gb_per_s"), e.get("worker_threads_per_gpu"), e.get("note"))
    print("c5", l.get("c5"))
    for o in l.get("other_configs") or []:
        print(o["config"], round(o["mrays_per_s"], 1))
except Exception as e:
    print("bench line unreadable:", e)
PY

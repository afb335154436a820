# rl_replay over worker counts and record modes (strict call pattern, built-in scene, 1024^2, 2048 reference batches)
for thr in 4 16; do for rec in host deferred; do
./robigo-luculenta_b200/rl_replay --width 1024 --height 1024 --threads $thr --batches 2048 --batch 524288 --mode strict --scene 2 --out /tmp/rr --records $rec | python -c "
import sys, json; r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('thr', r['threads'], r['records'], round(r['mrays_per_s'],1), 'h2d', r['h2d_bytes']>>20, 'd2h', r['d2h_bytes']>>20, r['worker_seconds'])"
done; done
./robigo-luculenta_b200/rl_replay --width 1024 --height 1024 --threads 16 --batches 2048 --batch 524288 --mode device --scene 2 --out /tmp/rr | python -c "
import sys, json; r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('thr', r['threads'], 'device mode', round(r['mrays_per_s'],1), r['worker_seconds'])"

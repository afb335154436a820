for rep in 1 2; do for thr in 4 6 8 12 16; do for a in 1 0; do
./robigo-luculenta_b200/rl_replay --width 1024 --height 1024 --threads $thr --batches 2048 --batch 524288 --mode strict --scene 2 --out /tmp/rr --async-render $a | python -c "
import sys, json; r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('thr', r['threads'], 'async', r['async_render'], round(r['mrays_per_s'],1), r['worker_seconds'])"
done; done; done
